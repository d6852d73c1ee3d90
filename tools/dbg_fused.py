"""Debug aid: fused training forward (aon_forward_train) vs per-layer GEMMs -- planes and gradients side by side."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from oracle import ref_cpu as O
from aon_b200 import nerf, train_tc, lib as L
from tests.test_gpu_parity import _make_net
DEV = "cuda:0"
R, S, x3 = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 65, True
sd = O.make_state_dict("vanilla", 0, sharp=False)
rays = {k: v[:R].contiguous().to(DEV) for k, v in O.sapien_rays(20, 24, seed=4).items()}
o, d, v = rays["rays_o"], rays["rays_d"], rays["viewdirs"]
g = torch.Generator().manual_seed(2)
t_vals = (2.0 + 4.0 * torch.rand(R, S, generator=g)).sort(-1).values.to(DEV)
g_up = (torch.randn(R * S, 4, generator=g) / R).to(DEV)        # O(1 / rays) like a mean-over-rays loss (train_tc.py scaling)
view_enc = nerf.pos_enc_cuda(v, 0, 4)
net = _make_net(nerf, "vanilla", sd, torch.device(DEV)).train()
mlp = net.fine_mlp
W = [l.weight.detach().contiguous() for l in mlp.linears()]
B = [l.bias.detach().contiguous() for l in mlp.linears()]
packed = L.pack_weights(L.KIND_VANILLA, L.PREC_TC_F16X3, W, B)
acts, enc, raw, _ = L.forward_train(L.KIND_VANILLA, L.PREC_TC_F16X3, packed, None, o, d, v, t_vals, S)
torch.cuda.synchronize()
samples = o[:, None, :] + t_vals[..., None] * d[:, None, :]
E_ref = nerf.pos_enc_cuda(samples, 0, 10).reshape(R * S, 63)
def untile(x):   # [tiles*128, C] tile order -> [R*S, C]
    return L.unpack_rows_tiled(x.contiguous(), R, S)
E_f = untile(enc.to_dense())[:, :63] / 8.0
print("E max abs diff", (E_f - E_ref).abs().max().item(), "per column max", (E_f - E_ref).abs().max(0).values[:8].tolist())
# reference activations through torch fp32
x = E_ref
hs = []
for i in range(8):
    inp = x if i != 5 else torch.cat([x, E_ref], -1)
    x = torch.relu(inp @ W[i].t() + B[i]); hs.append(x)
for i in range(8):
    a = untile(acts[i].to_dense()) / 8.0
    bits = acts[i].bits
    bt = untile(torch.stack([((bits >> k) & 1).float() for k in range(32)], -1).reshape(bits.shape[0], -1))
    print("h%d max abs diff %.3e   mask mismatches %d of %d (|h|>1e-4)" % (i, (a - hs[i]).abs().max().item(),
          int(((bt > 0) != (hs[i] > 0))[hs[i].abs() > 1e-4].sum()), hs[i].numel()))
# padded rows of the planes: finite?
for i, a in enumerate(acts):
    assert torch.isfinite(a.to_dense()).all(), i
out = {}
for fwd in ("layers", "fused"):
    for p in mlp.parameters():
        p.grad = None
    if fwd == "fused":
        rgb, sig = train_tc.vanilla_fused(o, d, v, t_vals, view_enc, mlp, x3=x3)
    else:
        rgb, sig = train_tc.vanilla_mlp(nerf.pos_enc_cuda(samples, 0, 10), view_enc, S, mlp, x3=x3)
    r = torch.cat([rgb, sig], -1).reshape(R * S, 4)
    r.backward(g_up)
    out[fwd] = {n: p.grad.clone() for n, p in mlp.named_parameters()}
# fp32 autograd yardstick
for p in mlp.parameters():
    p.grad = None
rgb, sig = mlp(nerf.pos_enc_cuda(samples, 0, 10).reshape(R, S, 63), view_enc)
torch.cat([rgb, sig], -1).reshape(R * S, 4).backward(g_up)
ref = {n: p.grad.clone() for n, p in mlp.named_parameters()}
for n in ref:
    m = ref[n].abs().max().item()
    print("%-28s fused-ref %.2e  layers-ref %.2e  fused-layers %.2e" % (n, (out["fused"][n] - ref[n]).abs().max().item() / m,
          (out["layers"][n] - ref[n]).abs().max().item() / m, (out["fused"][n] - out["layers"][n]).abs().max().item() / m))
gw = out["fused"]["pts_linears.0.weight"] - ref["pts_linears.0.weight"]
print("pts_linears.0.weight error per input column:", (gw.abs().max(0).values / ref["pts_linears.0.weight"].abs().max()).tolist())
