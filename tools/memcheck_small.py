"""Small end-to-end workload for compute-sanitizer (memcheck / synccheck): the fused image kernel (both model kinds, parity and
one-pass modes; fused launch + sample-segmented tail), the camera entry point, and one training step of each kind on the
tcgen05 GEMMs (fp16 hi+lo and single-plane), all at sizes a sanitizer run finishes in minutes.

    compute-sanitizer --tool memcheck python tools/memcheck_small.py"""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aon_b200 import lib as L, lit, nerf, synth

dev = torch.device("cuda:0")
H, W = 20, 31                                   # 620 rays: 3 CTA pairs, ragged
for kind in ("vanilla", "autodecoder"):
    sd = synth.make_state_dict(kind, 0, True)
    net = nerf.NeRF() if kind == "vanilla" else nerf.NeRF_AE_Art()
    net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
    net = net.to(dev).eval()
    lat = None
    if kind != "vanilla":
        z = torch.zeros(1, 128, device=dev)
        lat = {"density": z + 0.01, "color": z - 0.01, "articulation": torch.zeros(1, 32, device=dev) + 0.02}
    o, d = L.raygen(H, W, synth.sapien_focal(H), synth.sapien_camera(1), dev)
    rays = {"rays_o": o, "rays_d": d, "viewdirs": d}
    for mode in ("f16x3", "f16"):
        net.precision = L.PRECISIONS[mode]
        with torch.no_grad():
            out = net(rays, False, True, 2.0, 6.0) if lat is None else net(rays, False, True, 2.0, 6.0, lat)
            L.debug_no_tail_split(True)          # everything through the fused launch
            out2 = net(rays, False, True, 2.0, 6.0) if lat is None else net(rays, False, True, 2.0, 6.0, lat)
            L.debug_no_tail_split(False)
        torch.cuda.synchronize()
        assert torch.equal(out[1][0], out2[1][0]) and torch.isfinite(out[1][0]).all()
        print("render %s %s ok" % (kind, mode), flush=True)
    if kind == "vanilla":
        k = net.coarse_mlp.KIND
        pc = net._cache["coarse"].get(net.coarse_mlp, net.precision)
        pf = net._cache["fine"].get(net.fine_mlp, net.precision)
        img, _ = L.render_image(k, net.precision, pc, pf, None, None, synth.sapien_camera(1), synth.sapien_focal(H), H, W, 2.0, 6.0, True)
        torch.cuda.synchronize()
        print("render_image ok", flush=True)
for exp, gemm in (("vanilla", "tc"), ("vanilla_autodecoder", "tc"), ("vanilla", "tc16")):
    torch.manual_seed(0)
    s = lit.build_system(SimpleNamespace(exp_type=exp, run_max_steps=10, white_back=True, N_max_objs=1, N_obj_code_length=128)).to(dev)
    s.train()
    s.model.train_gemm = gemm
    o, d = L.raygen(8, 9, synth.sapien_focal(8), synth.sapien_camera(0), dev)
    batch = {"rays_o": o[None], "rays_d": d[None], "viewdirs": d[None], "target": torch.rand(1, 72, 3, device=dev)}
    if exp != "vanilla":
        batch.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([3], device=dev))
    opt = s.configure_optimizers()
    s.trainer = SimpleNamespace(global_step=0, is_global_zero=True)
    opt.zero_grad()
    loss = s.training_step(batch, 0)
    loss.backward()
    s.optimizer_step(0, 0, opt, 0, None, False, False, False)
    torch.cuda.synchronize()
    print("training step %s %s ok: loss %.4f" % (exp, gemm, loss.item()), flush=True)
