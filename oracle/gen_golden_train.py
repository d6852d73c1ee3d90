"""TEST INFRASTRUCTURE ONLY -- pins the oracle's TRAINING path (randomized sampling + autograd of loss0 + loss1) to the
reference and writes tests/golden/train_*.npz.

Runs ONLY in the build container (needs /root/reference).  For each model kind it loads the synthetic sharp weights into
the UNMODIFIED reference module, seeds torch's global generator, runs the reference's own forward with randomized=True
(helper.sample_along_rays / sorted_piecewise_constant_pdf draw with torch.rand, helper.py:122-127,228-231), takes
img2mse(coarse) + img2mse(fine) (model.py:271-276) and backpropagates with torch autograd; then re-seeds, runs the oracle
the same way and asserts loss and EVERY parameter gradient agree bit for bit.  The draws themselves are recovered by
replaying the generator (t_rand [R,65] then u [R,128]) and verified by injecting them into the oracle.  Stored: rays,
target, t_rand, u, the loss, a checksum (sum |g|) of every gradient and a few whole gradients.

    python -m oracle.gen_golden_train            # from the repo root
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_cpu as O
from oracle.gen_golden import OUT, import_reference, pick_rays

SEED = 7
KEEP = ("coarse_mlp.pts_linears.0.weight", "coarse_mlp.density_layer.weight", "fine_mlp.pts_linears.5.bias",
        "fine_mlp.views_linear.0.bias", "fine_mlp.rgb_layer.weight", "fine_mlp.deformation_layer.weight",
        "code_library.embedding_instance_articulation.weight")


def main():
    torch.set_num_threads(8)
    ref = import_reference()
    R = 33
    for kind in ("vanilla", "autodecoder"):
        sd = O.make_state_dict(kind, 0, sharp=True)
        rays = pick_rays(24, 32, R, 3)
        g = torch.Generator().manual_seed(11)
        target = torch.rand(R, 3, generator=g)
        # ---- the reference itself ----
        if kind == "vanilla":
            net = ref.M.NeRF()
            net.load_state_dict(sd)
            lat = None
        else:
            net = ref.MA.NeRF_AE_Art()
            net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
            lib = ref.CodeLibraryArticulated(type("H", (), {"N_max_objs": 1, "N_obj_code_length": 128})())
            lib.load_state_dict({k[len("code_library."):]: v for k, v in sd.items() if k.startswith("code_library.")})
            lat = lib({"instance_id": torch.tensor([0]), "articulation_id": torch.tensor([3])})
        net.train()
        torch.manual_seed(SEED)
        out = net(rays, True, True, 2.0, 6.0) if lat is None else net(rays, True, True, 2.0, 6.0, lat)
        loss_ref = ref.helper.img2mse(out[0][0], target) + ref.helper.img2mse(out[1][0], target)
        loss_ref.backward()
        g_ref = {k: p.grad for k, p in net.named_parameters()}
        if lat is not None:
            g_ref.update({"code_library." + k: p.grad for k, p in lib.named_parameters()})
        # ---- the oracle, same generator state ----
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lat_o = None if lat is None else O.code_library(p, torch.tensor([0]), torch.tensor([3]))
        torch.manual_seed(SEED)
        out_o = O.nerf_forward(p, rays, True, True, 2.0, 6.0, latents=lat_o)
        loss_o = O.img2mse(out_o[0][0], target) + O.img2mse(out_o[1][0], target)
        loss_o.backward()
        assert torch.equal(loss_o.detach(), loss_ref.detach()), (kind, loss_o.item(), loss_ref.item())
        worst = 0.0
        for k, gr in g_ref.items():
            go = p[k].grad
            assert go is not None and gr is not None, k
            assert torch.equal(go, gr), "%s %s: oracle gradient != reference gradient, max abs diff %g" % (kind, k, (go - gr).abs().max().item())
            worst = max(worst, (go - gr).abs().max().item())
        # ---- recover the draws and check that injecting them reproduces the run ----
        torch.manual_seed(SEED)
        t_rand, u = torch.rand(R, 65), torch.rand(R, 128)
        p2 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lat2 = None if lat is None else O.code_library(p2, torch.tensor([0]), torch.tensor([3]))
        out2 = O.nerf_forward(p2, rays, True, True, 2.0, 6.0, latents=lat2, t_rand=t_rand, u=u)
        loss2 = O.img2mse(out2[0][0], target) + O.img2mse(out2[1][0], target)
        loss2.backward()
        assert torch.equal(loss2.detach(), loss_ref.detach())
        for k, gr in g_ref.items():
            assert torch.equal(p2[k].grad, gr), k
        store = {"rays_o": rays["rays_o"], "rays_d": rays["rays_d"], "viewdirs": rays["viewdirs"], "target": target,
                 "t_rand": t_rand, "u": u, "loss": loss_ref.detach(), "sd_checksum": torch.tensor(sum(v.double().abs().sum().item() for v in sd.values()))}
        names = sorted(g_ref)
        store["grad_names"] = np.array(names)
        store["grad_abs_sums"] = torch.tensor([g_ref[k].double().abs().sum().item() for k in names], dtype=torch.float64)
        for k in KEEP:
            if k in g_ref:
                store["grad/" + k] = g_ref[k]
        path = os.path.join(OUT, "train_%s_sharp_R%d.npz" % (kind, R))
        np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in store.items()})
        print("%s: loss %.8f, %d gradients bit-equal to the reference -> %s (%.0f KB)" % (kind, loss_ref.item(), len(names), path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
