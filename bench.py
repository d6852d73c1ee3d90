#!/usr/bin/env python
"""bench.py -- rays/sec of the volume-rendering hot path (640x480, 65+193 samples/ray) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|f16x3|f16|bf16]
                    [--kind vanilla|autodecoder] [--impl ours|reference]

One "step" = one full coarse+fine render (A3 -> A4/A5/A6 -> A7 -> A4/A5/A6) of one
synthetic SAPIEN-shaped 640x480 view (307 200 rays) per GPU; ranks render contiguous ray blocks of an
N-image batch and exchange the rendered pixels with ONE all-gather (weak scaling).

* value     rays/s, inputs resident in HBM, CUDA-event timed, max over ranks
* e2e       same metric through the C-ABI host call aon_render_image_host() (pinned HOST rays in,
            HOST pixels out, H2D + D2H inside the timed region)
* roofline  dominant kernel = the fine-level fused render kernel; achieved = algorithmic MLP FLOP of one
            launch / its mean duration, timed live with CUDA events on the launching stream
* cpu_baseline  the oracle (CPU restatement of the reference, pinned bit-for-bit to it) timed on the
            host cores on a bounded ray sample (rank 0, N=1 only)

--impl reference times that same oracle port on the host cores with all threads (the reference is
100 % Python/PyTorch and cannot travel to the GPU box; oracle/ref_cpu.py is its bit-checked
restatement).  Only this leg and cpu_baseline execute anything under oracle/.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 480, 640
NEAR, FAR = 2.0, 6.0
S0, NF = 65, 128
S1 = S0 + NF
FLOP_PER_SAMPLE = {"vanilla": 1186816, "autodecoder": 1589760}   # SURVEY.md 8(d), 2*MAC of the reference shapes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("AON_BENCH_PRECISION", "auto"))
    ap.add_argument("--kind", default="vanilla", choices=["vanilla", "autodecoder"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the extra training-step block")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,power.draw"

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)

    def _once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            f = [x.strip() for x in out.strip().split(",")]
            self.samples.append(float(f[0]))
            self.max_mhz = float(f[1])
            bits = int(f[2], 16) if f[2].startswith("0x") else 0
            names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                     0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                     0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
            for b, n in names.items():
                if bits & b and n != "gpu_idle":
                    self.reasons.add(n)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(0.2)

    def __enter__(self):
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def report(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference)
# ---------------------------------------------------------------------------------------------------
def cpu_rays_per_sec(kind: str, budget_s: float, steps: int = 0, warmup: int = 0):
    """Times oracle.nerf_forward (full coarse+fine, no grad, deterministic) on chunks of the bench's rays.
    budget mode (steps=0): ~budget_s seconds of work.  step mode: exactly `steps` timed samples."""
    import torch
    from oracle import ref_cpu as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict(kind, seed=0, sharp=True)
    lat = None
    if kind != "vanilla":
        lat = O.code_library(sd, torch.tensor([0]), torch.tensor([6]), is_test=True)
    rays = O.sapien_rays(H, W, seed=0)
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(H * W, generator=g)

    def run(n, off):
        idx = perm[off:off + n]
        sub = {k: v[idx].contiguous() for k, v in rays.items()}
        t0 = time.perf_counter()
        with torch.no_grad():
            O.nerf_forward(sd, sub, False, True, NEAR, FAR, latents=lat)
        return time.perf_counter() - t0

    run(256, 0)                                  # warm the thread pool / allocator
    dt = run(512, 256)
    rate = 512 / dt
    if steps:
        per_step_s = max(1.0, min(6.0, 150.0 / max(1, steps + warmup)))
        n = int(min(3840, max(256, rate * per_step_s)))
        for i in range(warmup):
            run(n, (i * n) % (H * W - n))
        t = [run(n, ((warmup + i) * n) % (H * W - n)) for i in range(steps)]
        total = sum(t)
        return {"value": n * steps / total, "cores": cores, "ms_per_step": 1e3 * total / steps,
                "sample": "%d steps x %d random rays of the 640x480 view (reference chunk is 3840), full 65+193 path" % (steps, n)}
    n = int(min(3840, max(256, rate * budget_s / 3)))
    t = [run(n, 1024 + i * n) for i in range(3)]
    med = sorted(t)[1]
    return {"value": n / med, "cores": cores, "ms_per_step": 1e3 * med,
            "sample": "median of 3 x %d random rays of the 640x480 view, full 65+193 path" % n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_rays_per_sec(args.kind, 0.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "rays/sec", "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, "cpu"),
            "cpu_baseline": {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, precision):
    return {"workload": "sapien single-scene %s 640x480 eval render, 65 coarse + 193 fine samples/ray (the reference's "
                        "'64+128'), 307200 rays per GPU per step, white background, near 2 far 6" % args.kind,
            "rays_per_gpu_per_step": H * W, "samples_per_ray": S0 + S1, "precision": precision,
            "weights": "synthetic xavier-like init, density head sharpened (oracle.make_state_dict(sharp=True))",
            "parallelism": "rays sharded contiguously across %d GPU(s); one all-gather of [rays,5] pixels per step" % args.gpus,
            "l2": "256 MiB buffer overwritten between timed iterations (L2 flush)"}


def train_block(args, dev, world, rank, rays_o, rays_d):
    """training_step + backward + gradient all-reduce + optimizer_step of the Lightning-surface module at the reference's
    per-GPU batch (2048 rays vanilla, model.py:426; 4096 rays sapien_multi, sapien_multi.py:235), randomized sampling,
    3 warm-up + 10 timed steps, CUDA events, max over ranks.  MLP contractions: tcgen05 GEMMs (train_tc.py)."""
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from aon_b200 import dist as D
    from aon_b200 import lib as L
    from aon_b200 import lit
    torch.manual_seed(1234)                       # identical initial weights on every rank (DDP semantics) ...
    exp = "vanilla" if args.kind == "vanilla" else "vanilla_autodecoder"
    Rb = 2048 if args.kind == "vanilla" else 4096
    s = lit.build_system(SimpleNamespace(exp_type=exp, run_max_steps=100000, white_back=True, N_max_objs=1, N_obj_code_length=128)).to(dev)
    s.train()
    s.trainer = SimpleNamespace(global_step=0, is_global_zero=rank == 0)
    opt = s.configure_optimizers()
    torch.manual_seed(1234 + rank)                # ... and a different ray batch / different sampling draws per rank
    idx = torch.randperm(rays_o.shape[0], device=dev)[:Rb]
    batch = {"rays_o": rays_o[idx][None], "rays_d": rays_d[idx][None], "viewdirs": rays_d[idx][None],
             "target": torch.rand(1, Rb, 3, device=dev)}
    if args.kind != "vanilla":
        batch.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([3], device=dev))

    def step(i):
        opt.zero_grad()
        loss = s.training_step(batch, i)
        loss.backward()
        if world > 1:
            D.allreduce_mean_(opt.flat_grad)
        s.optimizer_step(0, i, opt, 0, None, False, False, False)
        s.trainer.global_step += 1

    for i in range(3):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    L.launch_count(reset=True)
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(3 + i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / K
    flop = 3 * Rb * (S0 + S1) * FLOP_PER_SAMPLE[args.kind]
    return {"value": world * Rb / (ms * 1e-3), "unit": "rays/s (training: forward + backward + all-reduce + Adam)", "ms_per_step": ms,
            "rays_per_gpu_per_step": Rb, "steps": K, "warmup": 3, "algorithmic_tflops_per_gpu": flop / (ms * 1e-3) / 1e12,
            "gemm": "tcgen05 kind::f16, fp16 hi+lo operands (3 MMAs per K step), fp32 accumulate" if s.model.train_gemm == "tc" else "library",
            "aon_launches_per_step": L.launch_count() // K, "grad_allreduce_bytes": opt.flat_grad.numel() * 4 if world > 1 else 0}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle to the real stdout and point fd 1 at stderr, so anything a
    library writes to stdout (NCCL prints its version banner there at NCCL_DEBUG >= VERSION, torchrun helpers, ...) cannot
    get in front of it."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _OUT


def emit(line: dict) -> None:
    out = claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from aon_b200 import lib as L
    from aon_b200 import nerf as NF_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the render path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = L.load()
    kind = L.KIND_VANILLA if args.kind == "vanilla" else L.KIND_AUTODECODER
    prec_name = args.precision
    if prec_name == "auto":
        prec_name = "f16x3" if lib.aon_packed_bytes(kind, L.PREC_TC_F16X3) > 0 else "fp32"
    prec = L.PRECISIONS[prec_name]

    # ---- synthetic scene: weights + this rank's view ------------------------------------------------
    # (rays come from our raygen kernel for a SAPIEN-shaped camera; weights are seeded random tensors)
    from aon_b200.synth import make_state_dict, sapien_camera
    sd = make_state_dict(args.kind, seed=0, sharp=True)
    if args.kind == "vanilla":
        net = NF_.NeRF()
        net.load_state_dict(sd)
        lat = None
    else:
        net = NF_.NeRF_AE_Art()
        net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
        from types import SimpleNamespace
        codes = NF_.CodeLibraryArticulated(SimpleNamespace(N_max_objs=1, N_obj_code_length=128))
        codes.load_state_dict({k[len("code_library."):]: v for k, v in sd.items() if k.startswith("code_library.")})
        with torch.no_grad():
            lat = {k: v.to(dev) for k, v in codes({"instance_id": torch.tensor([0]), "articulation_id": torch.tensor([6])},
                                                   is_test=True).items()}
    net = net.to(dev).eval()
    net.precision = prec
    focal = 0.5 * H / math.tan(math.radians(17.5))
    c2w = sapien_camera(seed=rank)
    rays_o, rays_d = L.raygen(H, W, focal, c2w, dev)
    R = H * W
    pc = net._cache["coarse"].get(net.coarse_mlp, prec)
    pf = net._cache["fine"].get(net.fine_mlp, prec)
    fc = ff = None
    if lat is not None:
        la = (lat["density"].contiguous(), lat["color"].contiguous(), lat["articulation"].contiguous())
        fc, ff = L.fold_latents(kind, prec, pc, *la), L.fold_latents(kind, prec, pf, *la)
    t0_tab = L.sample_along_rays(NEAR, FAR, S0, R, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world, R, 5, dtype=torch.float32, device=dev) if world > 1 else None
    pix = torch.empty(R, 5, dtype=torch.float32, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    fine_ev, coarse_ev, pdf_ev = [], [], []

    def step(record: bool):
        """one full render of this rank's 307 200 rays, device resident."""
        e = [ev() for _ in range(4)] if record else None
        if record: e[0].record()
        rgb0, acc0, dep0, w0 = L.render_level(kind, prec, pc, fc, rays_o, rays_d, rays_d, t0_tab, True, True)
        if record: e[1].record()
        t1 = L.sample_pdf(t0_tab, w0, NF)
        if record: e[2].record()
        rgb1, acc1, dep1, _ = L.render_level(kind, prec, pf, ff, rays_o, rays_d, rays_d, t1, True, False)
        if record: e[3].record()
        if world > 1:
            pix[:, :3] = rgb1
            pix[:, 3] = acc1
            pix[:, 4] = dep1
            dist.all_gather_into_tensor(gathered, pix)
        if record:
            coarse_ev.append((e[0], e[1])); pdf_ev.append((e[1], e[2])); fine_ev.append((e[2], e[3]))
        return rgb1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step(False)
    barrier()
    L.launch_count(reset=True)
    step_ev = []
    with ClockSampler(local) as clk:
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            a, b = ev(), ev()
            a.record()
            step(True)
            b.record()
            step_ev.append((a, b))
        barrier()
    launches = L.launch_count()
    ms = sum(a.elapsed_time(b) for a, b in step_ev)           # device time of the K steps on this rank
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    value = world * R * args.steps / (ms_total * 1e-3)

    fine_ms = sum(a.elapsed_time(b) for a, b in fine_ev) / len(fine_ev)
    coarse_ms = sum(a.elapsed_time(b) for a, b in coarse_ev) / len(coarse_ev)
    pdf_ms = sum(a.elapsed_time(b) for a, b in pdf_ev) / len(pdf_ev)

    # ---- the single-pass fp16 mode of the same kernel (PSNR-level agreement, not the 1e-4 parity mode): 3 steps ----
    fast = None
    if prec_name == "f16x3" and world == 1 and lib.aon_packed_bytes(kind, L.PREC_TC_F16) > 0:
        fprec = L.PREC_TC_F16
        fpc = L.pack_weights(kind, fprec, [l.weight for l in net.coarse_mlp.linears()], [l.bias for l in net.coarse_mlp.linears()])
        fpf = L.pack_weights(kind, fprec, [l.weight for l in net.fine_mlp.linears()], [l.bias for l in net.fine_mlp.linears()])
        ffc = fff = None
        if lat is not None:
            ffc, fff = L.fold_latents(kind, fprec, fpc, *la), L.fold_latents(kind, fprec, fpf, *la)

        def fstep():
            _, _, _, w0 = L.render_level(kind, fprec, fpc, ffc, rays_o, rays_d, rays_d, t0_tab, True, True)
            t1 = L.sample_pdf(t0_tab, w0, NF)
            a, b = ev(), ev()
            a.record()
            L.render_level(kind, fprec, fpf, fff, rays_o, rays_d, rays_d, t1, True, False)
            b.record()
            return a, b

        fstep()
        torch.cuda.synchronize()
        f0, f1, fe = ev(), ev(), []
        f0.record()
        for _ in range(3):
            flush.zero_()
            fe.append(fstep())
        f1.record()
        torch.cuda.synchronize()
        ffine = sum(a.elapsed_time(b) for a, b in fe) / len(fe)
        fast = {"precision": "f16 (single tcgen05 pass; not the parity mode)", "value": R * 3 / (f0.elapsed_time(f1) * 1e-3), "unit": "rays/s",
                "fine_kernel_ms": ffine, "achieved_tflops": FLOP_PER_SAMPLE[args.kind] * S1 * R / (ffine * 1e-3) / 1e12}

    # ---- e2e: host rays -> host pixels through the C-ABI call ------------------------------------------
    ho = rays_o.cpu().pin_memory(); hd = rays_d.cpu().pin_memory()
    hout = torch.empty(R, 5, dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    L.render_image_host(kind, prec, pc, pf, fc, ff, ho, hd, hd, NEAR, FAR, True, out=hout)
    barrier()
    t_e2e = time.perf_counter()
    for _ in range(e2e_steps):
        L.render_image_host(kind, prec, pc, pf, fc, ff, ho, hd, hd, NEAR, FAR, True, out=hout)   # synchronises
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * R * e2e_steps / dt.item()

    # ---- training step (BASELINE.json configs[3]: ray-sharded training; extra block, not the headline) -------------
    train = None
    if not args.no_train:
        try:
            train = train_block(args, dev, world, rank, rays_o, rays_d)
        except Exception as e:       # never lose the headline line to the extra block
            train = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tensor_mode = prec != L.PREC_FP32
        if tensor_mode:
            peak, peak_src = peaks.get("bf16_tflops_sustained", 1400.0), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured"
            if "bf16_tflops_sustained" not in peaks:
                peak_src = "fallback of B200_PROFILING.md (no MEASURED_PEAKS.json): 1.4 PFLOP/s sustained, of fallback"
        else:
            peak, peak_src = 148 * 128 * 2 * 1.965e-3, "fp32 FFMA peak 148 SM x 128 lanes x 2 x 1.965 GHz (CUDA-core mode; not a tensor-pipe number)"
        flop_fine = FLOP_PER_SAMPLE[args.kind] * S1 * R
        achieved = flop_fine / (fine_ms * 1e-3) / 1e12
        passes = 3 if prec_name == "f16x3" else 1
        traffic = None
        try:   # per-launch DRAM bytes of the fine kernel from the committed ncu --set full capture of this command
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_fine_%s_%s.json" % (args.kind, prec_name))))["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "f16x3": "f32 (3x fp16-split tcgen05, fp32 accumulate)",
                                           "f16": "f16", "bf16": "bf16"}[prec_name],
            "data": "synthetic", "config": workload_config(args, prec_name),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 3 * R * 12, "d2h_bytes_per_step": R * 20,
                    "steps": e2e_steps, "api": "aon_render_image_host (C ABI, pinned host buffers)"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "render_level (fine, %d samples/ray)" % S1, "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "mma_passes": passes, "executed_tflops": achieved * passes, "executed_frac": achieved * passes / peak,
                         "peak_source": peak_src, "kernel_ms": fine_ms,
                         "algorithmic_flop_per_launch": flop_fine,
                         "step_share": {"coarse_ms": coarse_ms, "sample_pdf_ms": pdf_ms, "fine_ms": fine_ms}},
            "clocks": clk.report(),
        }
        if fast is not None:
            fast["frac"] = fast["achieved_tflops"] / peak
            line["fast_mode"] = fast
        if train is not None:
            line["train"] = train
        if not args.no_cpu_baseline and world == 1:
            r = cpu_rays_per_sec(args.kind, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
