"""Training-path protocol (last sentence): a short training run WITH OUR KERNELS IN THE LOOP next to the same
run on library fp32 GEMMs under autograd (the reference's arithmetic on the GPU), same initial weights, same ray batches,
same random draws; test-split PSNR (interface.py:54-62 formula) at equal step counts.  Trajectories of a non-convex
optimisation separate after the first rounding difference, so the yardstick is the spread between two library-path runs
that differ only in the seed of the stratified / inverse-CDF draws.

    python tools/train_protocol.py [--steps 1000] [--wh 64 48] [--out profiles/r1_train_protocol.md]
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


def psnr(img, gt):
    mse = torch.mean((torch.clip(img, 0, 1) - torch.clip(gt, 0, 1)) ** 2)
    return float(-10.0 * torch.log10(mse))


def run(gemm, draw_seed, steps, marks, wh, root, dev):
    train, test = data.SapienDataset(root, "train", tuple(wh)), data.SapienDataset(root, "test_val", tuple(wh))
    torch.manual_seed(0)                                      # identical initial weights in every run
    hp = SimpleNamespace(exp_type="vanilla", run_max_steps=steps, img_wh=tuple(wh), white_back=True, N_max_objs=1, N_obj_code_length=128)
    system = lit.LitNeRF(hp, lr_delay_steps=min(200, steps // 5)).to(dev)
    system.model.train_gemm = gemm
    system.model.precision = L.PREC_TC_F16X3                  # eval renders in the parity mode
    system.setup(datasets={"train": train, "test": test})
    torch.manual_seed(1000 + draw_seed)                       # the sampling draws (t_rand, u) of training_step
    batches = train.ray_batches(2048, 0)                      # identical ray batches in every run
    tr = lit.Trainer(max_steps=0)
    out, t_train = {}, 0.0
    for m in marks:
        tr.max_steps = m
        torch.cuda.synchronize(); t0 = time.time()
        tr.fit(system, batches)
        torch.cuda.synchronize(); t_train += time.time() - t0
        system.eval()
        ps = [psnr(system.render_rays_test(test[i])["rgb"], test[i]["target"]) for i in range(len(test))]
        out[m] = (float(np.mean(ps)), system.logged["train/psnr1"])
    return out, t_train


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--wh", type=int, nargs=2, default=[64, 48])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    root = data.write_synthetic_scene(tempfile.mkdtemp(prefix="aon_scene_"), tuple(a.wh), n_train=40, n_val=1, n_test=4, seed=0)
    marks = [a.steps // 4, a.steps // 2, a.steps]
    runs = [("tcgen05 GEMMs + native adjoints (ours)", "tc", 0), ("library fp32 GEMMs under autograd", "torch", 0),
            ("library fp32 GEMMs, other sampling seed", "torch", 1)]
    res = [(name,) + run(g, s, a.steps, marks, a.wh, root, dev) for name, g, s in runs]
    lines = ["# Training protocol: synthetic SAPIEN-format scene %dx%d, vanilla model, 2048-ray batches, test-split PSNR (dB) at equal steps"
             % (a.wh[0], a.wh[1]), "",
             "| training path | " + " | ".join("step %d" % m for m in marks) + " | train wall time |", "|---|" + "---|" * (len(marks) + 1)]
    for name, out, t in res:
        lines.append("| %s | " % name + " | ".join("%.3f" % out[m][0] for m in marks) + " | %.1f s |" % t)
    d_ours = [res[0][1][m][0] - res[1][1][m][0] for m in marks]
    d_seed = [res[2][1][m][0] - res[1][1][m][0] for m in marks]
    lines += ["", "ours - library path, same draws: " + ", ".join("%+.3f dB" % d for d in d_ours),
              "library path, other sampling seed - library path: " + ", ".join("%+.3f dB" % d for d in d_seed), ""]
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)


if __name__ == "__main__":
    main()
