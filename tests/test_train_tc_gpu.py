"""GPU: the tcgen05 backward contraction (csrc/train_tc.cu: C (+)= A W and G += A^T X in one pass over A; K-major and
MN-major reads of the same shared-memory images) against float64 matmuls and against the fp32 FMA kernels it replaces."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_tc(A, W, X, col0, ldx_pad, a_rows, c_rows, x_rows, mask, C0, accumulate, G0, r_dev=None, R=None):
    from trackmpnn_b200 import _lib as L
    dev = A.device
    img = torch.empty(int(L.lib().tmpnn_bwd_tc_image_bytes()), dtype=torch.uint8, device=dev)
    part = torch.empty(int(L.lib().tmpnn_bwd_tc_partial_floats()), dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call('tmpnn_pack_w_tc', L.ptr(W), int(W.shape[1]), col0, L.ptr(img), L.stream())
    C, G = C0.clone(), G0.clone()
    L.call('tmpnn_rows_gemm_tc', L.ptr(r_dev), 0 if r_dev is not None else int(R), L.ptr(a_rows), L.ptr(c_rows), L.ptr(x_rows),
           L.ptr(mask), L.ptr(A), L.ptr(img), L.ptr(C), int(C.shape[1]), int(accumulate), L.ptr(X), int(X.shape[1]), L.ptr(part),
           L.ptr(G), int(G.shape[1]), L.ptr(status), L.stream())
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    return C, G


@pytest.mark.parametrize('n,scale', [(1, 1.0), (100, 1.0), (128, 1e-6), (1000, 1.0), (20001, 3e-4), (60000, 1.0)])
@pytest.mark.parametrize('accumulate', [0, 1])
def test_rows_gemm_tc_against_float64(n, scale, accumulate):
    """Plain rows (no gather, no mask), gradient-sized values down to 1e-6 (the bf16 split keeps the fp32 exponent range):
    C and G within 2^-14 of the float64 result relative to its largest entry; 60000 rows = 469 tiles > 3 per SM, so every
    barrier phase wraps."""
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(n)
    A = (torch.randn(n, 192, generator=g) * scale).to(dev)
    W = (torch.randn(192, 64, generator=g) * 0.01).to(dev)
    X = torch.randn(n, 64, generator=g).to(dev)
    C0 = torch.randn(n, 64, generator=g).to(dev) * scale * 0.1
    G0 = torch.randn(192, 64, generator=g).to(dev) * scale
    C, G = _run_tc(A, W, X, 0, 64, None, None, None, None, C0, accumulate, G0, R=n)
    Cw = A.double() @ W.double() + (C0.double() if accumulate else 0)
    Gw = G0.double() + A.double().t() @ X.double()
    tol = 2.0 ** -14
    assert float((C.double() - Cw).abs().max()) <= tol * float((A.double() @ W.double()).abs().max()) + 1e-7 * float(C0.abs().max())
    assert float((G.double() - Gw).abs().max()) <= tol * float((A.double().t() @ X.double()).abs().max()) + 1e-6 * float(G0.abs().max())


def test_rows_gemm_tc_gather_mask_and_device_count():
    """The calling forms of the backward pass: masked rows (detection rows inside the edge cell's sweep), gathered A / C / X
    rows (the detection list), a row count read from device memory, a 128-wide W taken as two 64-column passes with strided
    C / X / G -- against the FMA kernels tmpnn_rows_times_w / tmpnn_rows_outer on identical inputs."""
    from trackmpnn_b200 import _lib as L
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(7)
    n, nd = 5000, 700
    A = (torch.randn(n, 192, generator=g) * 1e-3).to(dev)
    mask = torch.where(torch.rand(n, generator=g) < 0.1, -1, 3).to(torch.int32).to(dev)
    st = L.stream()
    # (a) masked sweep over all rows, W 192 x 128, X with a leading dimension of 128
    W = (torch.randn(192, 128, generator=g) * 0.01).to(dev)
    X = torch.randn(n, 128, generator=g).to(dev)
    C_ref = torch.zeros(n, 128, device=dev); G_ref = torch.zeros(192, 128, device=dev)
    L.call('tmpnn_rows_times_w', None, n, None, None, L.ptr(mask), L.ptr(A), L.ptr(W), 128, L.ptr(C_ref), 128, 0, st)
    L.call('tmpnn_rows_outer', None, n, None, None, L.ptr(mask), L.ptr(A), L.ptr(X), 128, 128, L.ptr(G_ref), st)
    C = torch.zeros(n, 128, device=dev); G = torch.zeros(192, 128, device=dev)
    img = torch.empty(int(L.lib().tmpnn_bwd_tc_image_bytes()), dtype=torch.uint8, device=dev)
    part = torch.empty(int(L.lib().tmpnn_bwd_tc_partial_floats()), dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    for c0 in (0, 64):
        L.call('tmpnn_pack_w_tc', L.ptr(W), 128, c0, L.ptr(img), st)
        L.call('tmpnn_rows_gemm_tc', None, n, None, None, None, L.ptr(mask), L.ptr(A), L.ptr(img), L.ptr(C) + 4 * c0, 128, 0,
               L.ptr(X) + 4 * c0, 128, L.ptr(part), L.ptr(G) + 4 * c0, 128, L.ptr(status), st)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert float((C - C_ref).abs().max()) <= 2.0 ** -13 * float(C_ref.abs().max())
    assert float((G - G_ref).abs().max()) <= 2.0 ** -13 * float(G_ref.abs().max())
    assert float(C[mask < 0].abs().max()) == 0.0      # masked rows are not written
    # (b) the detection list: A gathered, C compact or gathered, X compact; the count lives on the device
    rows = torch.sort(torch.randperm(n, generator=g)[:nd]).values.to(torch.int32).to(dev)
    n_dev = torch.tensor([nd], dtype=torch.int32, device=dev)
    W2 = (torch.randn(192, 64, generator=g) * 0.01).to(dev)
    X2 = torch.randn(nd, 64, generator=g).to(dev)
    for c_rows, crows_n in ((None, nd), (rows, n)):
        C0 = torch.randn(crows_n, 64, generator=g).to(dev) * 1e-4
        C_ref = C0.clone(); G_ref = torch.zeros(192, 64, device=dev)
        L.call('tmpnn_rows_times_w', L.ptr(n_dev), 0, L.ptr(rows), L.ptr(c_rows), None, L.ptr(A), L.ptr(W2), 64, L.ptr(C_ref), 64, 1, st)
        L.call('tmpnn_rows_outer', L.ptr(n_dev), 0, L.ptr(rows), None, None, L.ptr(A), L.ptr(X2), 64, 64, L.ptr(G_ref), st)
        C, G = _run_tc(A, W2, X2, 0, 64, rows, c_rows, None, None, C0, 1, torch.zeros(192, 64, device=dev), r_dev=n_dev)
        assert float((C - C_ref).abs().max()) <= 2.0 ** -13 * float((C_ref - C0).abs().max())
        assert float((G - G_ref).abs().max()) <= 2.0 ** -13 * float(G_ref.abs().max())


def test_weight_gradients_are_reproducible():
    """The per-SM partials are summed in a fixed order: two runs give bit-identical G (the FMA kernel's float atomics do not)."""
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(3)
    n = 30000
    A = torch.randn(n, 192, generator=g).to(dev); W = torch.randn(192, 64, generator=g).to(dev); X = torch.randn(n, 64, generator=g).to(dev)
    z = torch.zeros(n, 64, device=dev); g0 = torch.zeros(192, 64, device=dev)
    _, G1 = _run_tc(A, W, X, 0, 64, None, None, None, None, z, 0, g0, R=n)
    _, G2 = _run_tc(A, W, X, 0, 64, None, None, None, None, z, 0, g0, R=n)
    assert torch.equal(G1, G2)
