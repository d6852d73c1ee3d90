"""Fused training forward (aon_forward_train) vs the same level rendered by the eval kernel (aon_render_level): what the
activation dump costs.  python tools/prof_fwd_train.py [R]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_cpu as O
from aon_b200 import lib as L, nerf, synth
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
o, d = L.raygen(480, 640, synth.sapien_focal(480), synth.sapien_camera(0), dev)
idx = torch.randperm(o.shape[0], device=dev)[:R]
o, d = o[idx].contiguous(), d[idx].contiguous()
net = nerf.NeRF().to(dev)
lins = net.fine_mlp.linears()
W, B = [l.weight.detach() for l in lins], [l.bias.detach() for l in lins]
def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ONLY = os.environ.get("AON_PROF_ONLY", "")          # e.g. "train" under ncu: training forward launches only
for name, prec in (("f16x3", L.PREC_TC_F16X3), ("f16", L.PREC_TC_F16)):
    packed = L.pack_weights(L.KIND_VANILLA, prec, W, B)
    for S in (65, 193):
        t = (2.0 + 4.0 * torch.rand(R, S, device=dev)).sort(-1).values.contiguous()
        if ONLY == "train":                  # profiling: just the training forward (coarse then fine), once
            L.forward_train(L.KIND_VANILLA, prec, packed, None, o, d, d, t, S)
            torch.cuda.synchronize()
            continue
        ms_eval = timeit(lambda: L.render_level(L.KIND_VANILLA, prec, packed, None, o, d, d, t, True))
        ms_train = timeit(lambda: L.forward_train(L.KIND_VANILLA, prec, packed, None, o, d, d, t, S))
        planes = (2 if name == "f16x3" else 1)
        gb = L.load().aon_train_tiles(R, S) * 128 * (9 * 256 + 128 + 64) * 2 * planes / 1e9
        print("%-6s R=%d S=%3d: eval level %.3f ms, training forward %.3f ms (%.2f GB of activation planes -> %.0f GB/s)"
              % (name, R, S, ms_eval, ms_train, gb, gb / (ms_train * 1e-3)))
