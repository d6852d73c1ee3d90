import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aon_b200 import lib
DEV = "cuda:0"
torch.manual_seed(5)
tiles = 333
M = tiles * 128
X = torch.randn(M, 256, device=DEV)
W = torch.randn(256, 256, device=DEV) / 16
A = lib.pack_rows(X, M, tiles, 256, 8.0)
BT = lib.pack_linear(W, True, 256, 256, 64.0)
act = torch.relu(torch.randn(M, 256, device=DEV))
mk = lib.pack_rows(act, M, tiles, 256, 8.0)
want0 = X.double() @ W.double()
for name, kw, want in (("mask only", dict(epi=lib.EPI_MASK, mask=(mk, 0)), want0 * (act > 0)),
                       ("colsum only", dict(epi=lib.EPI_MASK, colsum=True), want0),
                       ("mask+colsum", dict(epi=lib.EPI_MASK, mask=(mk, 0), colsum=True), want0 * (act > 0)),
                       ("plain", dict(epi=lib.EPI_MASK), want0)):
    dx = lib.PK(tiles, 256, DEV)
    cs = lib.gemm_nt([(A, 0, 256, BT, 0, 0)], 256, tiles, DEV, inv_scale=1.0 / 512, out=dx, **kw)
    err = ((dx.to_dense().double() - want).abs().view(tiles, 128, 256).amax((1, 2)) / want.abs().max())
    bad = (err > 1e-4).nonzero().flatten().tolist()
    print(name, "max err %.2e" % err.max().item(), "bad tiles:", bad[:20], len(bad))
    if cs is not None:
        e = ((cs.double() - want.view(tiles, 128, 256).sum(1)).abs().amax(1) / want.abs().max())
        print("   colsum bad tiles:", (e > 1e-4).nonzero().flatten().tolist()[:20])
