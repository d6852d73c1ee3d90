"""One forward-shaped (NT) and one wgrad-shaped (TN) tcgen05 training GEMM at the fine-level size of a 2048-ray batch
(395 264 rows, 256 x 256, hi+lo operands), for ncu captures and stand-alone timing.

    python tools/prof_gemm.py [--reps 5]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aon_b200 import lib as L

ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=5); args = ap.parse_args()
dev = "cuda:0"
torch.manual_seed(0)
M = 2048 * 193
tiles = M // 128
X = torch.randn(M, 256, device=dev)
W = torch.randn(256, 256, device=dev) / 16
b = torch.zeros(256, device=dev)
A = L.pack_rows(X, M, tiles, 256, 8.0)
B = L.pack_linear(W, False, 256, 256, 64.0)
out = L.PK(tiles, 256, dev)
G = L.pack_rows(X * 1e-4, M, tiles, 256, 2048.0)
def t(fn, name, flop):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print("%-28s %.3f ms  %.0f algorithmic TFLOP/s (x3 executed: %.0f)" % (name, ms, flop / ms * 1e-9, 3 * flop / ms * 1e-9))
flop = 2.0 * M * 256 * 256
t(lambda: L.gemm_nt([(A, 0, 256, B, 0, 0)], 256, tiles, dev, bias=b, relu=True, inv_scale=1 / 512, out=out, out_scale=8.0), "NT forward 256x256", flop)
t(lambda: L.gemm_nt([(G, 0, 256, B, 0, 0)], 256, tiles, dev, epi=L.EPI_MASK, mask=(A, 0), inv_scale=1 / 64, out=out, colsum=True), "NT dgrad+mask+colsum", flop)
out2 = L.PK(tiles, 256, dev)
assert out.bits is not None            # ReLU bit plane written by the forward call above
t(lambda: L.gemm_nt([(G, 0, 256, B, 0, 0)], 256, tiles, dev, epi=L.EPI_MASK, mask=(out, 0), inv_scale=1 / 64, out=out2, colsum=True), "NT dgrad, bit mask + colsum", flop)
t(lambda: L.gemm_nt([(G, 0, 256, B, 0, 0)], 256, tiles, dev, epi=L.EPI_MASK, mask=(out, 0), inv_scale=1 / 64, out=out2), "NT dgrad, bit mask", flop)
t(lambda: L.gemm_nt([(G, 0, 256, B, 0, 0)], 256, tiles, dev, epi=L.EPI_MASK, inv_scale=1 / 64, out=out2, colsum=True), "NT dgrad, colsum only", flop)
t(lambda: L.gemm_nt([(G, 0, 256, B, 0, 0)], 256, tiles, dev, epi=L.EPI_MASK, inv_scale=1 / 64, out=out2), "NT dgrad, plain", flop)
t(lambda: L.gemm_tn(G, 0, 2, A, 0, 256, 148), "TN wgrad 256x256", flop)

# single-plane operands (x3 = False): one MMA per K step, half the operand bytes
A1 = L.pack_rows(X, M, tiles, 256, 8.0, x3=False); B1 = L.pack_linear(W, False, 256, 256, 64.0, x3=False); out1 = L.PK(tiles, 256, dev, x3=False)
G1 = L.pack_rows(X * 1e-4, M, tiles, 256, 2048.0, x3=False)
t(lambda: L.gemm_nt([(A1, 0, 256, B1, 0, 0)], 256, tiles, dev, bias=b, relu=True, inv_scale=1 / 512, out=out1, out_scale=8.0, x3=False), "NT forward f16 (1 plane)", flop)
t(lambda: L.gemm_tn(G1, 0, 2, A1, 0, 256, 74, x3=False), "TN wgrad f16 (1 plane)", flop)
t(lambda: L.gemm_tn(G, 0, 2, A, 0, 256, 74), "TN wgrad x3, 74 splits", flop)
B128 = L.pack_linear(W[:128].contiguous(), False, 128, 256, 64.0); out128 = L.PK(tiles, 128, dev)
t(lambda: L.gemm_nt([(A, 0, 256, B128, 0, 0)], 128, tiles, dev, inv_scale=1 / 512, out=out128, out_scale=8.0), "NT forward N=128", flop / 2)
