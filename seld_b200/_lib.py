"""ctypes binding of libseld_b200.so (the C ABI declared in include/seld_b200.h).

There is deliberately no fallback: if the library is missing or no sm_100 device is
present, every compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SELD_B200_LIB') or os.path.join(_HERE, 'libseld_b200.so')      # (override: kernel experiments only)

MODE_FOA, MODE_MIC, MODE_FOA_TF = 0, 1, 2
LAYOUT_PLANAR_CL, LAYOUT_INTERLEAVED_LC = 0, 1
RNG_PHILOX_COUNTER, RNG_TF_EAGER_COMPAT, RNG_TF_EAGER_TWO_PASS = 0, 1, 2
DTYPE_CODES = {'float32': 0, 'float64': 1, 'float16': 2, 'bfloat16': 3, 'int32': 4, 'int64': 5, 'int16': 6, 'uint8': 7}

_c = ctypes
_vp, _i, _i64, _u64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint64, _c.c_float

# name -> (restype, argtypes); mirrors include/seld_b200.h one to one
SIGNATURES = {
    'seld_last_error': (_c.c_char_p, []),
    'seld_version': (_i, []),
    'seld_launch_count': (_i64, []),
    'seld_device_check': (_i, [_i]),
    'seld_plan_create': (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _c.POINTER(_vp)]),
    'seld_plan_destroy': (_i, [_vp]),
    'seld_plan_out_channels': (_i, [_vp]),
    'seld_plan_num_frames': (_i64, [_vp, _i64]),
    'seld_extract': (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _vp, _vp, _i64, _vp]),
    'seld_extract_chunks': (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _vp, _vp, _i64, _vp]),
    'seld_extract_tf': (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _vp, _vp]),
    'seld_extract_workspace_bytes': (_i64, [_vp, _i, _i64, _i]),
    'seld_extract_pcm16': (_i, [_vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _i64, _vp]),
    'seld_clip_max_decode': (_i, [_vp, _i, _vp, _vp]),
    'seld_finalize': (_i, [_i, _i, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _f, _vp, _vp]),
    'seld_stats_peer_buffer_bytes': (_i64, [_i]),
    'seld_stats_peer_allreduce': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'seld_stats_workspace_doubles': (_i64, [_i, _i]),
    'seld_stats': (_i, [_i, _i, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    'seld_stats_finish': (_i, [_i, _i, _vp, _vp, _vp, _vp]),
    'seld_mask': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _i64, _i, _i, _i, _i, _i, _u64, _u64, _i, _vp, _vp, _vp]),
    'seld_channel_remap': (_i, [_vp, _vp, _i64, _i64, _i, _i64, _vp, _vp]),
    'seld_channel_offset': (_i, [_vp, _vp, _i64, _i64, _i, _i, _vp, _vp]),
    'seld_augment_batch': (_i, [_vp, _vp, _i64, _i64, _i64, _i, _vp, _vp, _i64, _i, _i, _f, _i, _i, _i, _i, _i, _u64, _u64, _vp, _vp]),
    'seld_complex_spec': (_i, [_vp, _vp, _i, _i64, _f, _vp, _vp]),
    'seld_foa_iv': (_i, [_vp, _i64, _f, _vp, _vp]),
    'seld_gcc': (_i, [_vp, _i, _i64, _i, _i, _i, _vp, _vp]),
    'seld_gcc_gemm': (_i, [_vp, _vp, _i64, _f, _vp, _vp]),
}

_lib = None


class SeldError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeldError(f'{LIB_PATH} is missing: build it with `python -m seld_b200.build` '
                        '(seld_b200 has no CPU or PyTorch fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().seld_last_error()
        msg = msg.decode() if msg else ''
        if rc == -1:
            raise ValueError(msg or 'invalid argument')
        raise SeldError(f'seld_b200 error {rc}: {msg}')


def ptr(t):
    """Device (or pinned host) address of a torch tensor, or None."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_checked_devices = set()


def require_device():
    """Fail loudly unless the current CUDA device is an sm_100 part.  The driver query behind seld_device_check
    (cudaGetDeviceProperties) costs milliseconds and contends on driver locks, so its verdict is cached per device:
    calling it on every launch starved the GPU queue (sporadic 10x slow steps in per-launch timings)."""
    import torch
    if not torch.cuda.is_available():
        raise SeldError('seld_b200 needs a CUDA sm_100 (B200) device; none is visible and there is no CPU fallback')
    dev = torch.cuda.current_device()
    if dev not in _checked_devices:
        check(load().seld_device_check(-1))
        _checked_devices.add(dev)
