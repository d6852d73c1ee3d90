"""PSNR protocol of SURVEY.md 8(d) at BASELINE scale: same weights, two renderers, both model kinds.

For the vanilla model (sapien single scene) and the auto-decoder (sapien_multi, 2-part "scissor", 10 articulation states):
1. write a synthetic scene in the reference's on-disk format (aon_b200.data.write_synthetic_scene / _articulated);
2. train on it through the Lightning-surface module + minimal trainer with OUR kernels in the loop (tcgen05 training GEMMs,
   hand-written adjoints, flat Adam);
3. from that checkpoint render held-out views with the REFERENCE -- the reference's own modules imported from oracle/_ref
   (a byte copy of models/vanilla_nerf/model.py ... made by oracle/make_ref.py; the bit-pinned port if absent), on the host
   CPU in the reference's 3840-ray chunks -- and with the fused kernels in every precision mode;
4. report PSNR of each against the ground-truth images (interface.py:54-62 formula: clip to [0,1], -10 log10 mse, mean over
   images), the difference to the reference (bar: 0.1 dB), and PSNR(ours, reference render).
For the auto-decoder the held-out views are validation-style frames (a stored view of a known articulation state: ground
truth exists) plus test-path frames with INTERPOLATED articulation codes (odd rows of the 19-row table,
models/code_library.py:55-71; renderer-vs-renderer only: the reference's test split has no matching ground truth either).

    python tools/psnr_protocol.py [--steps 5000] [--wh 320 240] [--kinds vanilla autodecoder] [--out profiles/r2_psnr_protocol.md]
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


def reference_renderer(kind, state_dict):
    """-> (render(rays_cpu, latents | None) -> rgb [R,3], "reference" | "port")"""
    sd = {k[len("model."):]: v for k, v in state_dict.items() if k.startswith("model.")}
    try:
        from oracle import ref_import
        root = ref_import.available_root()
        ref = ref_import.import_reference(root) if root else None
    except Exception as e:
        print("reference modules not importable (%s); using the oracle port" % e)
        ref = None
    if ref is None:
        def render(rays, lat):
            with torch.no_grad():
                return O.render_chunked(sd, rays, True, 2.0, 6.0, latents=lat)["comp_rgb"]
        return render, "port"
    net = (ref.M.NeRF() if kind == "vanilla" else ref.MA.NeRF_AE_Art()).eval()
    net.load_state_dict(sd)

    def render(rays, lat, chunk=3840):
        out = []
        R = rays["rays_o"].shape[0]
        with torch.no_grad():
            for i in range(0, R, chunk):
                sub = {k: v[i:i + chunk] for k, v in rays.items()}
                r = net(sub, False, True, 2.0, 6.0) if lat is None else net(sub, False, True, 2.0, 6.0, lat)
                out.append(r[1][0])
        return torch.cat(out, 0)

    return render, "reference"


def run(steps, wh, modes=("fp32", "f16x3", "f16", "bf16"), n_train=40, n_test=4, seed=0, log=print, kind="vanilla"):
    """short form used by tests/test_gpu_data.py: rows (name, PSNR, difference to the reference, min cross-PSNR | inf)"""
    table, _ = run_kind(kind, steps, wh, modes, n_test, seed, log, n_train=n_train)
    return [(n, p, d, float("inf") if c is None else c) for n, p, d, c in table]


def run_kind(kind, steps, wh, modes, n_eval, seed, log, n_train=60):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    tmp = tempfile.mkdtemp(prefix="aon_scene_")
    if kind == "vanilla":
        root = data.write_synthetic_scene(tmp, tuple(wh), n_train=n_train, n_val=1, n_test=n_eval, seed=seed)
        train, test = data.SapienDataset(root, "train", tuple(wh)), data.SapienDataset(root, "test_val", tuple(wh))
        hp = SimpleNamespace(exp_type="vanilla", run_max_steps=steps, img_wh=tuple(wh), white_back=True, N_max_objs=1, N_obj_code_length=128)
        system = lit.LitNeRF(hp, lr_delay_steps=min(500, steps // 5)).to(dev)
        batches = train.ray_batches(2048, seed)
    else:
        root = data.write_synthetic_articulated(tmp, tuple(wh), n_states=10, n_images=10, seed=seed)
        train = data.SapienDatasetMulti(root, "train", tuple(wh), seed=seed)
        hp = SimpleNamespace(exp_type="vanilla_autodecoder", run_max_steps=steps, img_wh=tuple(wh), white_back=True, N_max_objs=1, N_obj_code_length=128)
        system = lit.LitNeRF_AutoDecoder(hp, lr_delay_steps=min(500, steps // 5)).to(dev)
        batches = train.ray_batches()
    system.setup(datasets={"train": train})
    t0 = time.time()
    lit.Trainer(max_steps=steps).fit(system, batches)
    torch.cuda.synchronize()
    train_s = time.time() - t0
    log("%s: trained %d steps in %.1f s (%.2f ms/step): train/psnr1 %.2f dB" % (kind, steps, train_s, 1e3 * train_s / steps, system.logged["train/psnr1"]))
    sd_full = {k: v.detach().cpu() for k, v in system.state_dict().items()}
    system.eval()
    ref_render, ref_kind = reference_renderer(kind, sd_full)
    # ---- held-out frames: (rays on the GPU, ground truth or None, latents batch or None, label)
    frames = []
    if kind == "vanilla":
        for i in range(n_eval):
            b = test[i]
            frames.append((b, b["target"].cpu(), None, "test view %d" % i))
    else:
        val = data.SapienDatasetMulti(root, "val", tuple(wh), seed=seed + 1)
        for i in range(n_eval):
            b = val[i]
            frames.append((b, b["target"].cpu(), False, "stored view, articulation state %d" % int(b["articulation_id"])))
        te = data.SapienDatasetMulti(root, "test_val", tuple(wh), eval_inference="x", seed=seed)
        for idx in (7, 13):
            frames.append((te[idx], None, True, "test-path frame %d (interpolated articulation row %d)" % (idx, idx)))
    rows = {m: {"gt": [], "cross": []} for m in modes}
    ref_gt, t_ref = [], 0.0
    for b, gt, is_test, label in frames:
        rays_cpu = {k: b[k].cpu() for k in ("rays_o", "rays_d", "viewdirs")}
        lat = lat_cpu = None
        if kind != "vanilla":
            with torch.no_grad():
                lat = system.code_library({"instance_id": b["instance_id"], "articulation_id": b["articulation_id"]}, is_test=bool(is_test))
            lat_cpu = {k: v.cpu() for k, v in lat.items()}
        t0 = time.time()
        ref_img = ref_render(rays_cpu, lat_cpu)
        t_ref += time.time() - t0
        if gt is not None:
            ref_gt.append(psnr(ref_img, gt))
        for m in modes:
            system.model.precision = L.PRECISIONS[m]
            with torch.no_grad():
                rays_gpu = {k: b[k] for k in ("rays_o", "rays_d", "viewdirs")}
                out = (system.model(rays_gpu, False, True, 2.0, 6.0) if lat is None else system.model(rays_gpu, False, True, 2.0, 6.0, lat))[1][0].cpu()
            if gt is not None:
                rows[m]["gt"].append(psnr(out, gt))
            rows[m]["cross"].append(psnr(out, ref_img))
        log("  %s: reference %s dB" % (label, "%.3f" % ref_gt[-1] if gt is not None else "-"))
    p_ref = float(np.mean(ref_gt))
    table = [("reference (%s, host CPU, %d frames in %.0f s)" % ("its own modules from oracle/_ref" if ref_kind == "reference" else "oracle port", len(frames), t_ref), p_ref, 0.0, None)]
    for m in modes:
        p = float(np.mean(rows[m]["gt"]))
        table.append(("fused kernels, %s" % m, p, p - p_ref, float(np.min(rows[m]["cross"]))))
    return table, train_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--wh", type=int, nargs=2, default=[320, 240])
    ap.add_argument("--kinds", nargs="+", default=["vanilla", "autodecoder"])
    ap.add_argument("--modes", nargs="+", default=["fp32", "f16x3", "f16", "bf16"])
    ap.add_argument("--n-eval", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lines = ["# PSNR protocol (SURVEY 8d): synthetic SAPIEN-format scenes %dx%d, %d training steps with our kernels in the loop, "
             "held-out frames rendered by the reference and by the fused kernels from the same checkpoint" % (a.wh[0], a.wh[1], a.steps), ""]
    bad = []
    for kind in a.kinds:
        table, train_s = run_kind(kind, a.steps, a.wh, a.modes, a.n_eval, 0, print)
        lines += ["## %s (trained in %.0f s)" % (kind, train_s), "",
                  "| renderer | PSNR vs ground truth (dB) | difference to the reference (dB) | min PSNR vs the reference render (dB) |", "|---|---|---|---|"]
        for name, p, dlt, cross in table:
            lines.append("| %s | %.3f | %+.4f | %s |" % (name, p, dlt, "-" if cross is None else "%.1f" % cross))
            if cross is not None and abs(dlt) > 0.1 and "bf16" not in name:
                bad.append((kind, name, dlt))
        lines.append("")
    txt = "\n".join(lines) + "\n"
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
