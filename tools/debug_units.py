import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import torch.nn.functional as F
from oracle import ref_cpu as O
from tests.test_gpu_parity import _make_net
from tests.test_gpu_tc import _preacts
from aon_b200 import lib, nerf
dev = torch.device("cuda:0")
kind, mode = "autodecoder", "f16"
prec = lib.PRECISIONS[mode]
sd = O.make_state_dict(kind, 0, sharp=True)
net = _make_net(nerf, kind, sd, dev)
k = net.coarse_mlp.KIND
info = lib.debug_program_info(k, prec)
rays = O.sapien_rays(10, 16, seed=3)
o, d, v = (rays[x].to(dev) for x in ("rays_o", "rays_d", "viewdirs"))
lat = {kk: vv.to(dev) for kk, vv in O.code_library(sd, torch.tensor([0]), torch.tensor([5]), is_test=True).items()}
lins = net.coarse_mlp.linears()
packed = lib.pack_weights(k, prec, [l.weight for l in lins], [l.bias for l in lins])
folded = lib.fold_latents(k, prec, packed, lat["density"], lat["color"], lat["articulation"])
for rep in range(2):
    dbg = torch.zeros(info["n_units"], 128, 256, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.debug_set_buffers(dbg, err)
    t = lib.sample_along_rays(2.0, 6.0, 65, o.shape[0], dev)
    lib.render_level(k, prec, packed, folded, o, d, v, t, True)
    torch.cuda.synchronize()
    lib.debug_set_buffers(None, None)
    with torch.no_grad():
        pts = o[:128] + t[0] * d[:128]
        want = _preacts(net.coarse_mlp, kind, pts, v[:128], lat)
    for ui in range(len(want)):
        got = dbg[ui][:, :want[ui].shape[1]]
        diff = (got - want[ui]).abs()
        print(rep, "unit", ui, "N", want[ui].shape[1], "max err %.4f" % diff.max().item(), "max ref %.3f" % want[ui].abs().max().item(),
              "err by 32-col block:", ["%.3f" % diff[:, j:j+32].max().item() for j in range(0, want[ui].shape[1], 32)],
              "err by row quadrant:", ["%.3f" % diff[j:j+32].max().item() for j in range(0, 128, 32)])
    # partial-K hypothesis for unit 2: input h1 = relu(unit1 preact)
    W2 = lins[2].weight.detach(); b2 = lins[2].bias.detach()
    h1 = F.relu(want[1])
    for blk in [(0, 64), (64, 128)]:
        part = h1[:, blk[0]:blk[1]] @ W2[:, blk[0]:blk[1]].T
        print("   unit2 minus K-block", blk, "residual %.4f" % ((dbg[2][:, :128] - (want[2] - part)).abs().max().item()))
