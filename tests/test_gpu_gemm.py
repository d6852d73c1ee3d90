"""GPU: the tcgen05 training GEMMs (csrc/gemm_tc.cu) against fp64 torch matmuls of the same (dequantised) operands.

NT (forward / dgrad; K-major operands, multi-segment, bias / ReLU / ReLU-mask epilogues, packed + fp32 outputs) and
TN (wgrad; the same packed bytes read as MN-major operands, split-K partials + fixed-order reduce), x3 (hi + lo) and
single-plane modes.  Bars: x3 1e-5 of the output scale (fp32-grade), single plane 2e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-9)).item()


@pytest.mark.parametrize("x3", [True, False])
def test_pack_round_trip(built_lib, x3):
    lib = built_lib
    torch.manual_seed(0)
    src = torch.randn(300, 63, device=DEV)
    pk = lib.pack_rows(src, 300, 3, 64, 8.0, x3=x3)
    d = pk.to_dense() / 8.0
    assert d.shape == (384, 64)
    assert (d[300:] == 0).all() and (d[:, 63] == 0).all()
    assert _rel(d[:300, :63], src) < (2e-6 if x3 else 1e-3)
    # per-ray rows broadcast to their samples
    v = torch.randn(5, 27, device=DEV)
    pk = lib.pack_rows(v, 5 * 65, 3, 32, 1.0, row_div=65, x3=x3)
    d = pk.to_dense()
    assert _rel(d[:325, :27], v.repeat_interleave(65, 0)) < (2e-6 if x3 else 1e-3)


@pytest.mark.parametrize("x3", [True, False])
@pytest.mark.parametrize("N,K", [(256, 256), (128, 256), (256, 64), (16, 256), (256, 16), (128, 32)])
def test_gemm_nt_linear(built_lib, x3, N, K):
    lib = built_lib
    torch.manual_seed(1)
    M, tiles = 500, 4
    X = torch.randn(M, K, device=DEV)
    W = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV)
    A = lib.pack_rows(X, M, tiles, K, 8.0, x3=x3)
    B = lib.pack_linear(W, False, N, K, 64.0, x3=x3)
    out = lib.PK(tiles, N, DEV, x3)
    out32 = torch.zeros(tiles * 128, N, device=DEV)
    lib.gemm_nt([(A, 0, K, B, 0, 0)], N, tiles, DEV, bias=b, relu=True, inv_scale=1.0 / 512, out_f32=out32, out=out,
                out_scale=8.0, x3=x3)
    want = torch.relu(X.double() @ W.double().t() + b.double())
    tol = 1e-5 if x3 else 3e-3
    assert _rel(out32[:M], want) < tol
    assert _rel(out.to_dense()[:M] / 8.0, want) < (tol if x3 else 4e-3)


def test_gemm_nt_segments_mask_rows(built_lib):
    """Two segments (skip concat), a row window of the weight (dgrad w.r.t. the first 256 inputs of a 320-input layer through the
    transposed packing) and the ReLU-mask epilogue."""
    lib = built_lib
    torch.manual_seed(2)
    M, tiles = 384, 3
    H, E = torch.randn(M, 256, device=DEV), torch.randn(M, 63, device=DEV)
    W = torch.randn(256, 319, device=DEV) / 18
    Ah, Ae = lib.pack_rows(H, M, tiles, 256, 8.0), lib.pack_rows(E, M, tiles, 64, 8.0)
    B = lib.pack_linear(W, False, 256, 320, 64.0)
    out32 = torch.zeros(tiles * 128, 256, device=DEV)
    lib.gemm_nt([(Ah, 0, 256, B, 0, 0), (Ae, 0, 64, B, 256, 0)], 256, tiles, DEV, inv_scale=1.0 / 512, out_f32=out32)
    want = torch.cat([H, E], 1).double() @ W.double().t()
    assert _rel(out32, want) < 1e-5
    # dgrad: dX[:, :256] = dY @ W[:, :256], masked by an activation
    dY = torch.randn(M, 256, device=DEV) * 1e-4
    act = torch.relu(torch.randn(M, 256, device=DEV))
    G = lib.pack_rows(dY, M, tiles, 256, 4096.0)
    BT = lib.pack_linear(W, True, 320, 256, 64.0)
    mk = lib.pack_rows(act, M, tiles, 256, 8.0)
    dx = lib.PK(tiles, 256, DEV)
    cs = lib.gemm_nt([(G, 0, 256, BT, 0, 0)], 256, tiles, DEV, epi=lib.EPI_MASK, mask=(mk, 0), inv_scale=1.0 / 64, out=dx, colsum=True)
    want = (dY.double() @ W.double()[:, :256]) * (act > 0)
    assert _rel(dx.to_dense() / 4096.0, want) < 1e-5
    assert cs.shape[0] in (tiles, 148) and cs.shape[1] == 256 and _rel(cs.sum(0) / 4096.0, want.sum(0)) < 1e-5      # fused bias-gradient column sums
    # second row window: the 63 encoding inputs
    dxe = torch.zeros(tiles * 128, 64, device=DEV)
    lib.gemm_nt([(G, 0, 256, BT, 0, 256)], 64, tiles, DEV, epi=lib.EPI_MASK, inv_scale=1.0 / (64 * 4096.0), out_f32=dxe, n_valid=63)
    assert _rel(dxe[:M, :63], dY.double() @ W.double()[:, 256:]) < 1e-5


@pytest.mark.parametrize("x3", [True, False])
@pytest.mark.parametrize("n_out,n_in,splits", [(256, 256, 5), (128, 64, 3), (256, 16, 2), (128, 32, 40)])
def test_gemm_tn_wgrad(built_lib, x3, n_out, n_in, splits):
    lib = built_lib
    torch.manual_seed(3)
    tiles = 9
    M = tiles * 128 - 37
    dY = torch.randn(M, n_out, device=DEV) * 1e-4
    X = torch.randn(M, n_in, device=DEV)
    G = lib.pack_rows(dY, M, tiles, n_out, 4096.0, x3=x3)
    A = lib.pack_rows(X, M, tiles, n_in, 8.0, x3=x3)
    part = lib.gemm_tn(G, 0, n_out // 128, A, 0, n_in, splits, x3=x3)
    dW = torch.zeros(n_out, n_in + 3, device=DEV)
    lib.wgrad_reduce(part, 1.0 / (4096.0 * 8.0), dW, 3, n_out, n_in)
    want = dY.double().t() @ X.double()
    assert _rel(dW[:, 3:], want) < (1e-5 if x3 else 3e-3)
    assert (dW[:, :3] == 0).all()
    # transposed use (heads: 16 output features): rows = inputs, cols = outputs
    part = lib.gemm_tn(A if n_in >= 128 else G, 0, 1, G if n_in >= 128 else A, 0, 16, splits, x3=x3)
    if n_in >= 128:
        dWt = torch.zeros(16, n_in, device=DEV)
        lib.wgrad_reduce(part, 1.0 / (4096.0 * 8.0), dWt, 0, 128, 16, transpose=True)
        assert _rel(dWt[:, :128], want[:16, :128]) < (1e-5 if x3 else 3e-3)
    # bias gradient
    cs = lib.colsum_packed(G, 4).sum(0) / 4096.0
    assert _rel(cs, dY.double().sum(0)) < (1e-5 if x3 else 3e-3)


def test_gemm_nt_persistent_many_tiles(built_lib):
    """More row tiles than SMs: every persistent CTA walks 2-3 tiles with the ring and the double-buffered TMEM accumulator
    carried across tiles (forward epilogue, then the dgrad / mask / column-sum epilogue)."""
    lib = built_lib
    torch.manual_seed(5)
    tiles = 333
    M = tiles * 128 - 50
    X = torch.randn(M, 256, device=DEV)
    W = torch.randn(256, 256, device=DEV) / 16
    b = torch.randn(256, device=DEV)
    A = lib.pack_rows(X, M, tiles, 256, 8.0)
    B = lib.pack_linear(W, False, 256, 256, 64.0)
    out = lib.PK(tiles, 256, DEV)
    lib.gemm_nt([(A, 0, 256, B, 0, 0)], 256, tiles, DEV, bias=b, relu=True, inv_scale=1.0 / 512, out=out, out_scale=8.0)
    want = torch.relu(X.double() @ W.double().t() + b.double())
    got = out.to_dense()[:M] / 8.0
    assert _rel(got, want) < 1e-5
    dY = torch.randn(M, 256, device=DEV) * 1e-3
    G = lib.pack_rows(dY, M, tiles, 256, 1024.0)
    BT = lib.pack_linear(W, True, 256, 256, 64.0)
    dx = lib.PK(tiles, 256, DEV)
    assert out.bits is not None and out.bits.shape == (tiles * 128, 8)           # ReLU mask bit plane from the forward epilogue
    bits = ((out.bits[:M, :, None] >> torch.arange(32, device=DEV)) & 1).reshape(M, 256).bool()
    assert torch.equal(bits, got > 0)
    cs = lib.gemm_nt([(G, 0, 256, BT, 0, 0)], 256, tiles, DEV, epi=lib.EPI_MASK, mask=(out, 0), inv_scale=1.0 / 64, out=dx, colsum=True)
    # the mask is the sign of OUR forward output (a pre-activation within fp32 rounding of zero may round either way)
    wantd = (dY.double() @ W.double()) * (got > 0)
    assert _rel(dx.to_dense()[:M] / 1024.0, wantd) < 1e-5
    assert _rel(cs.sum(0) / 1024.0, wantd.sum(0)) < 1e-5
    # the same through a mask given as an activation plane (no bit plane) and two segments (skip concat shape)
    act = torch.relu(torch.randn(M, 256, device=DEV))
    mk = lib.pack_rows(act, M, tiles, 256, 8.0)
    E = torch.randn(M, 63, device=DEV)
    Ae = lib.pack_rows(E, M, tiles, 64, 8.0)
    W2 = torch.randn(256, 319, device=DEV) / 18
    B2 = lib.pack_linear(W2, False, 256, 320, 64.0)
    o2 = lib.PK(tiles, 256, DEV)
    lib.gemm_nt([(A, 0, 256, B2, 0, 0), (Ae, 0, 64, B2, 256, 0)], 256, tiles, DEV, epi=lib.EPI_MASK, mask=(mk, 0), inv_scale=1.0 / 512,
                out=o2, out_scale=8.0)
    want2 = (torch.cat([X, E], 1).double() @ W2.double().t()) * (act > 0)
    assert _rel(o2.to_dense()[:M] / 8.0, want2) < 1e-5
