"""PSNR protocol of SURVEY.md 8(d): same weights, two renderers.

1. write a synthetic SAPIEN-format scene (aon_b200.data.write_synthetic_scene);
2. train the vanilla model on it through the Lightning-surface module + minimal trainer (kernels for the sampling
   stages, torch autograd for the MLP -- the native backward is 8f/F1);
3. render the test split from that checkpoint with the REFERENCE path (the CPU oracle, pinned bit-for-bit to the
   reference's own code) and with the fused kernels in every precision mode;
4. report PSNR of each against the ground-truth images (interface.py:54-62 formula) and the difference (bar: 0.1 dB).

    python tools/psnr_protocol.py [--steps 1500] [--wh 64 48] [--out profiles/r1_psnr_protocol.md]
Test infrastructure (imports oracle/).
"""
import argparse
import os
import sys
import tempfile
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from aon_b200 import data, lib as L, lit
from oracle import ref_cpu as O


def psnr(img, gt):
    mse = torch.mean((torch.clip(img, 0, 1) - torch.clip(gt, 0, 1)) ** 2)
    return float(-10.0 * torch.log10(mse))


def run(steps, wh, modes=("fp32", "f16x3", "f16", "bf16"), n_train=40, n_test=4, seed=0, log=print):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    root = data.write_synthetic_scene(tempfile.mkdtemp(prefix="aon_scene_"), tuple(wh), n_train=n_train, n_val=1, n_test=n_test, seed=seed)
    train, test = data.SapienDataset(root, "train", tuple(wh)), data.SapienDataset(root, "test_val", tuple(wh))
    hp = SimpleNamespace(exp_type="vanilla", run_max_steps=steps, img_wh=tuple(wh), white_back=True, N_max_objs=1, N_obj_code_length=128)
    system = lit.LitNeRF(hp, lr_delay_steps=min(200, steps // 5)).to(dev)
    system.setup(datasets={"train": train, "test": test})
    t0 = time.time()
    lit.Trainer(max_steps=steps).fit(system, train.ray_batches(2048, seed))
    torch.cuda.synchronize()
    log("trained %d steps in %.1f s: train/psnr1 %.2f dB" % (steps, time.time() - t0, system.logged["train/psnr1"]))
    sd = {k: v.detach().cpu() for k, v in system.model.state_dict().items()}
    system.eval()
    rows, ref_imgs = [], []
    for i in range(len(test)):
        b = test[i]
        rays_cpu = {k: b[k].cpu() for k in ("rays_o", "rays_d", "viewdirs")}
        with torch.no_grad():
            ref = O.render_chunked(sd, rays_cpu, True, 2.0, 6.0)["comp_rgb"]
        ref_imgs.append(ref)
    gts = [test[i]["target"].cpu() for i in range(len(test))]
    p_ref = float(np.mean([psnr(r, g) for r, g in zip(ref_imgs, gts)]))
    rows.append(("reference path (CPU oracle, fp32)", p_ref, 0.0, float("inf")))
    for mode in modes:
        system.model.precision = L.PRECISIONS[mode]
        ps, cross = [], []
        for i in range(len(test)):
            out = system.render_rays_test(test[i])["rgb"].cpu()
            ps.append(psnr(out, gts[i]))
            cross.append(psnr(out, ref_imgs[i]))
        rows.append(("fused kernels, %s" % mode, float(np.mean(ps)), float(np.mean(ps)) - p_ref, float(np.mean(cross))))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--wh", type=int, nargs=2, default=[64, 48])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = run(a.steps, a.wh)
    lines = ["# PSNR protocol (SURVEY 8d): synthetic SAPIEN-format scene %dx%d, vanilla model trained %d steps, test split" % (a.wh[0], a.wh[1], a.steps), "",
             "| renderer | PSNR vs ground truth (dB) | difference to the reference path (dB) | PSNR vs the reference render (dB) |", "|---|---|---|---|"]
    for name, p, dlt, cross in rows:
        lines.append("| %s | %.3f | %+.4f | %s |" % (name, p, dlt, "-" if cross == float("inf") else "%.1f" % cross))
    txt = "\n".join(lines) + "\n"
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)
    bad = [r for r in rows[1:] if abs(r[2]) > 0.1 and "bf16" not in r[0]]
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
