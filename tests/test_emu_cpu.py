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


@pytest.mark.parametrize('n_fft,sr,n_mels', [(1024, 24000, 64), (1024, 48000, 64), (512, 16000, 64), (256, 8000, 32),
                                               (2048, 48000, 64), (1024, 24000, 40), (1024, 24000, 128)])
def test_mel_piece_layout_invariants(emu, n_fft, sr, n_mels):
    """The record layout the bin phase writes and the gather reads (csrc/mel_pieces.h): every piece gets its own slot
    below n_slots, the zero slot is never written, and in the segment-major layout the j-th piece of segment s is exactly
    where the gather looks for it -- slot s, slot s + 65, then the two overflow slots ov[s] names."""
    n_bins = n_fft // 2 + 1
    fb = np.ascontiguousarray(tables.melscale_fbanks_htk(n_bins, sr, n_mels).numpy())
    info = np.zeros(8, np.int32)
    slot0, slot1, ov = np.zeros(64, np.int32), np.zeros(64, np.int32), np.zeros(64, np.int32)
    endmask = np.zeros(64, np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert emu.emu_mel_layout(p(fb), n_bins, n_mels, p(info), p(slot0), p(slot1), p(ov), p(endmask)) == 0
    bpt, n_pieces, max_per_seg, seg_major, n_slots, zero_slot, pitch = (int(v) for v in info[:7])
    assert bpt == -(-n_bins // 64) and 64 * bpt <= n_fft
    seg_of_bin = [int(np.flatnonzero(fb[k])[0]) if fb[k].any() else -1 for k in range(n_bins)]
    # walk every team lane the way bin_phase does
    pieces = []                                     # (slot, seg) in write order
    for u in range(64):
        slot, nxt = int(slot0[u]), int(slot1[u])
        for i in range(bpt):
            if (int(endmask[u]) >> i) & 1:
                k = u * bpt + i
                assert k < n_bins and seg_of_bin[k] >= 0
                pieces.append((slot, seg_of_bin[k]))
                slot, nxt = nxt, nxt + 1
    assert len(pieces) == n_pieces
    slots = [s for s, _ in pieces]
    assert len(set(slots)) == n_pieces and max(slots) < n_slots and min(slots) >= 0
    assert not seg_major or (n_mels <= 64 and max_per_seg <= 4)
    assert seg_major or n_mels > 64 or max_per_seg > 4 or n_fft != 1024      # the production bank must take the fast layout
    if seg_major:
        assert zero_slot not in slots and zero_slot < n_slots
        by_seg = {}
        for s, seg in pieces:
            by_seg.setdefault(seg, []).append(s)
        for seg in range(64):
            want = by_seg.get(seg, [])
            reads = [seg, seg + pitch, int(ov[seg]) & 0xffff, int(ov[seg]) >> 16]
            assert reads[:len(want)] == want, (seg, want, reads)                    # pieces in rank order where the gather reads
            assert all(r == zero_slot or r not in slots for r in reads[len(want):])  # the rest are never-written slots
    else:
        assert slots == list(range(n_pieces))                                       # compact layout: piece index


@pytest.mark.parametrize('n_fft,sr,n_mels,bank', [(1024, 24000, 64, 'htk'), (1024, 48000, 64, 'htk'), (1024, 24000, 64, 'tf'),
                                                    (1024, 24000, 40, 'htk'), (512, 16000, 64, 'htk')])
def test_mel_lane_form_invariants(emu, n_fft, sr, n_mels, bank):
    """The flush-free lane form of the bank (csrc/mel_pieces.h build_lane_form): the records the 64 lanes write and the
    table-driven gather reproduce the dense projection, every lane's reads stay inside the spectrum, the zero record is never
    a lane's own, and -- for the production banks -- the layout is conflict-free under the shared-memory model it is tuned for."""
    n_bins = n_fft // 2 + 1
    if bank == 'htk':
        fb = np.ascontiguousarray(tables.melscale_fbanks_htk(n_bins, sr, n_mels).numpy())
    else:
        fb = np.ascontiguousarray(np.asarray(tables.tf_mel_weight_matrix(n_mels, n_bins, sr), dtype=np.float32))
    info = np.zeros(16, np.int32)
    bpt = -(-n_bins // 64)
    lane_beg, gtab, w4 = np.zeros(64, np.int32), np.zeros(64 * 8, np.int32), np.zeros(64 * bpt * 4, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert emu.emu_mel_lanes(p(fb), n_bins, n_mels, p(info), p(lane_beg), p(gtab), p(w4)) == 0
    ok, bpt2, gn0, gn1, rec_words, gmax, zero_rec, spread, modelled = (int(v) for v in info[:9])
    assert bpt2 == bpt and gmax == 8
    if n_fft == 1024 and n_mels == 64:
        assert ok and spread                       # the production banks take the lane form, conflict-free bin reads
    if not ok:
        return
    assert lane_beg.min() >= 0 and lane_beg.max() + bpt <= n_fft         # S[k] and S[N - k] stay inside the n_fft-long spectrum
    rng = np.random.default_rng(5)
    val = rng.uniform(0.5, 2.0, n_fft).astype(np.float64)
    rec = np.zeros((zero_rec + 1) * rec_words)
    w4 = w4.reshape(64, bpt, 4).astype(np.float64)
    for u in range(64):
        for i in range(bpt):
            k = lane_beg[u] + i
            if w4[u, i].any():
                assert k < n_bins
            rec[u * rec_words:u * rec_words + 4] += w4[u, i] * val[k]
    gtab = gtab.reshape(64, gmax)
    assert (gtab < (zero_rec + 1) * rec_words).all()
    for m in range(n_mels):
        n = gn0 if m < 32 else gn1
        assert (gtab[m, n:] == zero_rec * rec_words).all()              # the gather stops at gather_n: nothing may sit past it
        got = rec[gtab[m, :n]].sum()
        want = 0.25 * (fb[:, m].astype(np.float64) * val[:n_bins]).sum()
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (m, got, want)
    # the shared-memory model: 64-bit reads 16 lanes at a time (bank pair = word pair index mod 16) ...
    if spread:
        for g in range(0, 64, 16):
            assert len({int(b) % 16 for b in lane_beg[g:g + 16]}) == 16
    # ... 32-bit gather reads 32 lanes at a time, one wavefront per distinct address sharing a bank
    cost = 0
    for w, n in ((0, gn0), (1, gn1)):
        for t in range(n):
            banks = {}
            for m in range(32 * w, min(32 * w + 32, n_mels)):
                a = int(gtab[m, t])
                if a != zero_rec * rec_words:
                    banks.setdefault(a % 32, set()).add(a)
            cost += max([len(v) for v in banks.values()] + [1])
    assert cost == modelled
    if n_fft == 1024 and n_mels == 64 and bank == 'htk':
        assert cost <= gn0 + gn1 + 1
