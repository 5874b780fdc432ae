"""tcgen05 GCC lag projection (seld_gcc_gemm) vs a float32 matmul and vs the irfft it replaces."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(a16, bt16, scale):
    """a16 [rows, 1024], bt16 [64, 1024] float16 CUDA tensors (row-major) -> packs the operand images, runs the kernel."""
    from seld_b200 import _lib, tables
    a_img = torch.from_numpy(tables.gcc_operand_image(a16.cpu().numpy())).cuda()
    bt_img = torch.from_numpy(tables.gcc_operand_image(bt16.cpu().numpy())).cuda()
    out = torch.empty(a16.shape[0], 64, dtype=torch.float32, device='cuda')
    _lib.check(_lib.load().seld_gcc_gemm(_lib.ptr(a_img), _lib.ptr(bt_img), a16.shape[0], float(scale), _lib.ptr(out),
                                         _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('rows', [128, 37, 1000, 128 * 148 * 5 + 5])
def test_gemm_matches_float32_matmul(rows):
    g = torch.Generator().manual_seed(rows)
    a = (torch.rand(rows, 1024, generator=g) * 2 - 1).to(torch.float16).cuda()
    bt = (torch.rand(64, 1024, generator=g) * 2 - 1).to(torch.float16).cuda()
    got = _run(a, bt, 0.25)
    want = 0.25 * (a.float() @ bt.float().t())
    assert torch.allclose(got, want, atol=2e-3, rtol=1e-4), float((got - want).abs().max())


def test_gemm_is_the_pruned_irfft():
    """Random unit phasors -> the 64 centre lags of irfft (reference feature_extractor.py:210-211)."""
    from seld_b200 import tables
    rows = 600
    g = torch.Generator().manual_seed(7)
    ph = torch.exp(1j * (torch.rand(rows, 513, generator=g, dtype=torch.float64) * 2 * np.pi))
    ph[:, 0] = torch.sign(ph[:, 0].real)
    ph[:, 512] = torch.sign(ph[:, 512].real)
    cc = torch.fft.irfft(ph, n=1024, dim=1)
    want = torch.cat([cc[:, -32:], cc[:, :32]], 1).float()
    a = torch.empty(rows, 1024, dtype=torch.float64)
    a[:, 0] = ph[:, 0].real
    a[:, 1] = ph[:, 512].real
    a[:, 2::2] = ph[:, 1:512].real
    a[:, 3::2] = ph[:, 1:512].imag
    bt = torch.from_numpy(tables.gcc_basis(1024, 64)).cuda()
    got = _run(a.to(torch.float16).cuda(), bt, 1.0 / tables.GCC_BASIS_SCALE).cpu()
    err = float((got - want).abs().max())
    assert err <= 2e-4, err                       # FP16 operands: ~5e-5 typical, tolerance of the path is 1e-3
