"""Full-size parity: ALL 307 200 rays of the bench's 640x480 view rendered by the reference's own modules on the host (oracle/_ref,
3840-ray chunks like model.py:421-428) and by the fused kernel in the parity mode, compared pixel by pixel.
(The GPU tests compare at most 3840 rays with the oracle because the reference needs ~2.5 min of 16 cores per image.)

    python tools/parity_full_image.py [--kind vanilla|autodecoder] [--out profiles/r2_parity_full_image.md]
Test infrastructure (imports oracle/ and bench.py's reference renderer)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from oracle import ref_cpu as O
from aon_b200 import lib as L, nerf

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="vanilla")
ap.add_argument("--out", default="")
ap.add_argument("--rays", type=int, default=0, help="first N rays only (0 = the whole image)")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)
fn, _, which, sd, lat = bench.cpu_renderer(args.kind)
rays = O.sapien_rays(bench.H, bench.W, seed=0)
n_rays = args.rays or bench.H * bench.W
rays = {k: v[:n_rays].contiguous() for k, v in rays.items()}
t0 = time.perf_counter()
ref = [[], [], []]
for lo in range(0, n_rays, 3840):
    out = fn({k: v[lo:lo + 3840] for k, v in rays.items()})[1]
    for j in range(3):
        ref[j].append(out[j])
ref = [torch.cat(x) for x in ref]
t_ref = time.perf_counter() - t0
net = (nerf.NeRF() if args.kind == "vanilla" else nerf.NeRF_AE_Art())
net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
net = net.to(dev).eval()
rd = {k: v.to(dev) for k, v in rays.items()}
lines = ["# Full-image parity, %s: %d rays of the bench's 640x480 view, fused kernel vs the %s on the host (%d cores, %.0f s)"
         % (args.kind, n_rays, "reference's own modules" if which == "reference" else "oracle port", os.cpu_count() or 1, t_ref), "",
         "| mode | max rel rgb | max rel acc | max rel depth | rays above 1e-4 (any output) | 99.99th percentile (worst output) | PSNR(ours, reference) |",
         "|---|---|---|---|---|---|---|"]
outlier_note = ""
for mode in ("f16x3", "fp32", "f16"):
    net.precision = L.PRECISIONS[mode]
    with torch.no_grad():
        a = (rd, False, True, 2.0, 6.0) + (() if lat is None else ({k: v.to(dev) for k, v in lat.items()},))
        got = [x.cpu() for x in net(*a)[1]]
    rel = []
    for j in range(3):
        g, r = got[j].double().reshape(n_rays, -1), ref[j].double().reshape(n_rays, -1)
        rel.append(((g - r).abs() / r.abs().clamp_min(1e-2)).max(1).values)          # per ray
    worst = torch.stack(rel, 1).max(1).values
    mse = ((got[0].double() - ref[0].double()) ** 2).mean().item()
    lines.append("| %s | %.2e | %.2e | %.2e | %d of %d | %.2e | %.1f dB |" % (
        mode, rel[0].max(), rel[1].max(), rel[2].max(), int((worst > 1e-4).sum()), n_rays,
        torch.quantile(worst, 0.9999).item() if n_rays <= 16_000_000 else float("nan"), -10 * torch.log10(torch.tensor(max(mse, 1e-30))).item()))
    if mode == "f16x3":
        bad = (worst > 1e-4).nonzero().flatten()
        n_bad = bad.numel()
        if n_bad > 2048:                         # float64 is slow on the host: a fixed random subset of the rays above the bar
            bad = bad[torch.randperm(n_bad, generator=torch.Generator().manual_seed(0))[:2048]]
        if n_bad > 0:
            # the rays above the bar, evaluated in float64 (oracle restatement of the reference): how far is the reference's OWN
            # fp32 result from that truth on the same rays?
            sd64 = {k: t.double() for k, t in sd.items()}
            l64 = None if lat is None else {k: t.double() for k, t in lat.items()}
            with torch.no_grad():
                t64 = O.nerf_forward(sd64, {k: v[bad].double() for k, v in rays.items()}, False, True, 2.0, 6.0, latents=l64)[1]
            r64 = lambda x, j: ((x[bad].double().reshape(bad.numel(), -1) - t64[j].reshape(bad.numel(), -1)).abs()
                                / t64[j].reshape(bad.numel(), -1).abs().clamp_min(1e-2)).max(1).values
            ref_far = torch.stack([r64(ref[j], j) for j in range(3)], 1).max(1).values
            our_far = torch.stack([r64(got[j], j) for j in range(3)], 1).max(1).values
            outlier_note = ("%d rays are above 1e-4 in `f16x3`.  " % n_bad) + ("The %d of them evaluated in float64 (the same network, oracle restatement): the reference's own "
                            "fp32 result is off by up to %.2e (median %.2e, %d of them above 1e-4), ours by up to %.2e (median %.2e) -- these are the "
                            "rays on which the fine level is discontinuous in the coarse weights (an importance sample crossing a bin edge), for "
                            "ANY fp32 evaluation." % (bad.numel(), ref_far.max(), ref_far.median(), int((ref_far > 1e-4).sum()), our_far.max(), our_far.median()))
        else:
            outlier_note = ""
lines += ["", outlier_note] if outlier_note else []
lines += ["", "Relative error = |ours - reference| / max(|reference|, 1e-2) (tests/test_gpu_parity.py::relerr).  Weights: the bench's synthetic "
          "state dict with the sharpened density head (oracle.make_state_dict(sharp=True)), the case in which the reference's own fp32 "
          "evaluation is furthest from an fp64 one (bench `parity` block: its fp32-vs-fp64 floor on this view is 3.2e-5)."]
text = "\n".join(lines) + "\n"
print(text)
if args.out:
    open(args.out, "w").write(text)
