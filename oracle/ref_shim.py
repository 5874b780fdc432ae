"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference feature extractor.

Imports /root/reference/feature_extractor.py under the 3-item in-memory shim that
SURVEY.md section 8(c) describes (no reference file is modified or copied):

  1. a stub ``tensorflow`` module (reference utils.py:2,9,99 touch TF at import
     time; the extractor itself only uses utils.get_device, feature_extractor.py:11),
  2. ``torchaudio.functional.complex_norm`` (removed after torchaudio 0.10; the 0.8
     definition; call site feature_extractor.py:63),
  3. a lazy-``n_stft`` ``torchaudio.transforms.MelScale`` (0.8 inferred n_stft on
     first forward; call site feature_extractor.py:59-60).

/root/reference only exists in the authoring container, so this module is used
by (a) oracle/make_golden.py to generate tests/golden/*.npz and (b) CPU tests that
skip when the reference is absent.  Nothing under seld_b200/ may import it.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('SELD_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'feature_extractor.py'))


def _install_tf_stub():
    if 'tensorflow' in sys.modules:
        return
    tf = types.ModuleType('tensorflow')
    tf.__spec__ = importlib.machinery.ModuleSpec('tensorflow', None)
    keras = types.ModuleType('tensorflow.keras')
    optim = types.ModuleType('tensorflow.keras.optimizers')
    sched = types.ModuleType('tensorflow.keras.optimizers.schedules')
    sched.LearningRateSchedule = object
    optim.schedules = sched
    optim.Optimizer = object
    keras.optimizers = optim
    tf.keras = keras
    sys.modules['tensorflow'] = tf


def _install_torchaudio_compat():
    import torch
    import torchaudio

    if not hasattr(torchaudio.functional, 'complex_norm'):
        def complex_norm(x, power=1.0):
            if torch.is_complex(x):
                x = torch.view_as_real(x)
            return x.pow(2.).sum(-1).pow(0.5 * power)
        torchaudio.functional.complex_norm = complex_norm

    real_cls = torchaudio.transforms.MelScale
    if getattr(real_cls, '_seld_lazy', False):
        return

    class LazyMelScale(torch.nn.Module):
        _seld_lazy = True

        def __init__(self, n_mels=128, sample_rate=16000, f_min=0., f_max=None,
                     n_stft=None, **kw):
            super().__init__()
            self._args = dict(n_mels=n_mels, sample_rate=sample_rate,
                              f_min=f_min, f_max=f_max, **kw)
            self._n_stft = n_stft
            self._impl = None
            self._device = None

        def to(self, device):
            self._device = device
            return self

        def forward(self, x):
            if self._impl is None:
                n_stft = self._n_stft or x.size(-2)
                self._impl = real_cls(n_stft=n_stft, **self._args)
                if self._device is not None:
                    self._impl = self._impl.to(self._device)
            return self._impl(x)

    torchaudio.transforms.MelScale = LazyMelScale


_ref_module = None


def load_reference_extractor():
    """Return the reference ``feature_extractor`` module (imported once)."""
    global _ref_module
    if _ref_module is not None:
        return _ref_module
    if not reference_available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    _install_tf_stub()
    _install_torchaudio_compat()
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        saved = {k: sys.modules.pop(k) for k in ('feature_extractor', 'data_utils', 'utils')
                 if k in sys.modules}
        _ref_module = importlib.import_module('feature_extractor')
        # keep reference helper modules private to this handle
        for k in ('feature_extractor', 'data_utils', 'utils'):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return _ref_module
