"""The kernel's per-lane code (seld_b200/csrc/extract_core.cuh) executed lane by lane on the CPU and compared with the
reference's outputs: checks frame indexing, reflection, tables, the packed FFT, the mel pieces and the pruned GCC
transform without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from cases import CASES, PROD, case_input, check_features, input_matches_golden, load_golden
from seld_b200 import tables

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, 'emu')
EMU_SO = os.path.join(EMU_DIR, 'libemu_extract.so')
DEPS = [os.path.join(EMU_DIR, 'emu_extract.cpp')] + [os.path.join(HERE, '..', 'seld_b200', 'csrc', f)
                                                       for f in ('extract_core.cuh', 'seld_common.cuh', 'mel_pieces.h')]


@pytest.fixture(scope='module')
def emu():
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in DEPS):
        subprocess.run(['g++', '-std=c++17', '-O2', '-shared', '-fPIC', '-o', EMU_SO, DEPS[0]], check=True)
    return ctypes.CDLL(EMU_SO)


def run_emu(lib, wav, sr, mode, n_mels=64, t_out=None, layout=0, **kw):
    n_fft, win, hop = tables.resolve_stft(**kw)
    window = tables.padded_window(n_fft, win)
    tw = tables.twiddles(n_fft)
    fb = np.ascontiguousarray(tables.melscale_fbanks_htk(n_fft // 2 + 1, sr, n_mels).numpy())
    x = np.ascontiguousarray(wav.numpy() if layout == 0 else wav.numpy().T)
    n_samples = wav.shape[1]
    t_raw = 1 + n_samples // hop
    t_out = t_raw if t_out is None else t_out
    c = 7 if mode == 'foa' else 10
    out = np.full((t_out, n_mels, c), np.nan, np.float32)
    cmax = np.zeros(1, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.emu_extract(p(x), layout, 1, ctypes.c_longlong(n_samples), n_fft, hop, n_mels, 0 if mode == 'foa' else 1,
                         p(window), p(tw), p(fb), t_out, p(out), p(cmax))
    assert rc == 0
    valid = min(t_raw, t_out)
    out[:valid, :, :4] = np.maximum(out[:valid, :, :4], cmax[0] - 80.0)      # what seld_finalize does
    return out, float(cmax[0])


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_emulated_kernel_vs_reference_golden(emu, name, mode):
    wav, sr, n_mels, kw = case_input(name)
    g = load_golden(name)
    if not input_matches_golden(wav, g):
        pytest.skip('torch RNG stream differs')
    got, cmax = run_emu(emu, wav, sr, mode, n_mels, **kw)
    check_features(got, g[mode], mode, f'{name}/{mode}')
    assert abs(cmax - (g[mode][..., :4].max())) <= 1e-4 or name == 'zeros'


def test_emulated_interleaved_layout_and_padding(emu):
    wav, sr, n_mels, kw = case_input('prod')
    g = load_golden('prod')
    if not input_matches_golden(wav, g):
        pytest.skip('torch RNG stream differs')
    t_raw = g['foa'].shape[0]
    got, _ = run_emu(emu, wav, sr, 'foa', n_mels, layout=1, **kw)
    check_features(got, g['foa'], 'foa', 'interleaved')
    got, _ = run_emu(emu, wav, sr, 'mic', n_mels, t_out=t_raw + 3, **kw)
    check_features(got[:t_raw], g['mic'], 'mic', 'padded')
    assert np.all(got[t_raw:] == 0.0)
    got, _ = run_emu(emu, wav, sr, 'foa', n_mels, t_out=t_raw - 5, **kw)
    check_features(got, g['foa'][:t_raw - 5], 'foa', 'truncated')       # floor still from ALL frames


def test_ordered_key_roundtrip(emu):
    emu.emu_key_roundtrip.restype = ctypes.c_float
    emu.emu_key_roundtrip.argtypes = [ctypes.c_float]
    for v in (-100.0, -0.0, 0.0, 1e-30, 25.9, -1e30, float('inf'), float('-inf')):
        assert emu.emu_key_roundtrip(v) == np.float32(v)
