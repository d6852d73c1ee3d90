#!/usr/bin/env python
"""bench.py -- rays/sec of the volume-rendering hot path (640x480, 65+193 samples/ray) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|f16x3|f16|bf16]
                    [--kind vanilla|autodecoder] [--impl ours|reference]

One "step" = one full render (ray generation A1/A2 -> coarse level A3-A6 -> hierarchical sampling A7 -> fine level,
SURVEY.md 8a) of one synthetic SAPIEN-shaped 640x480 view (307 200 rays) per GPU through ONE fused kernel launch
(aon_render_image); at N > 1 every rank renders its own view and the pixels are exchanged with ONE all-gather
(weak scaling).  The JSON line carries, next to the contract keys:

* value     rays/s, inputs resident in HBM, CUDA-event timed, max over ranks
* e2e       same metric through the C-ABI host call aon_render_image_host() (pinned HOST rays in, HOST pixels out,
            H2D + D2H inside the timed region)
* roofline  dominant kernel = the fused render kernel; achieved = algorithmic MLP FLOP of one launch / its mean
            duration, timed live with CUDA events on the launching stream (a launch of exactly 16 full waves of CTA pairs)
* cpu_baseline / parity   the reference's own modules (oracle/_ref, a byte copy made by oracle/make_ref.py; else the
            bit-pinned oracle port) timed on the host cores on a bounded ray sample (rank 0, N=1 only), and the GPU
            render of those same rays compared with what the CPU produced
* sharded   ONE 640x480 image split over the N ranks (contiguous pixel blocks, one all-gather): ms / image, and an
            in-line check that the gathered image equals the single-GPU render bit for bit
* c2        BASELINE configs[2]: auto-decoder (2-part articulated model) at 320x240 and 640x480 on one GPU
* c4        BASELINE configs[4]: auto-decoder, 4 articulation states (ids 0, 6, 12, 18 of the 19-row test table) x one
            640x480 view each, every image sharded over the N ranks
* train     BASELINE configs[3]: training step (forward + backward + gradient all-reduce + Adam), 2048 rays per GPU;
            train_fast = the same step on single fp16 operand planes (not the fp32-grade default)
* fast_mode the single-pass fp16 mode of the same kernel (PSNR-level agreement, not the 1e-4 parity mode)

--impl reference times the reference's own CPU implementation (same modules as cpu_baseline) with all host threads.
Only that arm and the cpu_baseline / parity leg execute anything under oracle/.
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
C4_ARTICULATIONS = (0, 6, 12, 18)                                 # of the 19-row test table (code_library.py:55-71)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("AON_BENCH_PRECISION", "auto"))
    ap.add_argument("--kind", default="vanilla", choices=["vanilla", "autodecoder"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step block")
    ap.add_argument("--no-extras", action="store_true", help="skip the sharded / c2 / c4 / fast_mode blocks")
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
# CPU legs: the reference's own modules (oracle/_ref) when present, else the oracle port
# ---------------------------------------------------------------------------------------------------
def cpu_renderer(kind: str):
    """(fn(rays) -> [(rgb, acc, depth)] x 2, no grad, deterministic;  train_fn(rays, target) -> loss after backward;
    "reference" | "port";  state dict;  latents)."""
    import torch
    from oracle import ref_cpu as O
    sd = O.make_state_dict(kind, seed=0, sharp=True)
    lat = None
    if kind != "vanilla":
        lat = O.code_library(sd, torch.tensor([0]), torch.tensor([6]), is_test=True)
    root = None
    try:
        from oracle import ref_import
        root = ref_import.available_root()
        ref = ref_import.import_reference(root) if root else None
    except Exception as e:          # a missing third-party module on this box: fall back to the bit-pinned port, say so
        sys.stderr.write("bench: reference modules not importable (%s: %s); CPU arm uses the oracle port\n" % (type(e).__name__, e))
        ref = None
    if ref is not None:
        if kind == "vanilla":
            net = ref.M.NeRF()
            net.load_state_dict(sd)
        else:
            net = ref.MA.NeRF_AE_Art()
            net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})

        def fn(rays):
            net.eval()
            with torch.no_grad():
                return net(rays, False, True, NEAR, FAR) if lat is None else net(rays, False, True, NEAR, FAR, lat)

        def train_fn(rays, target):
            net.train()
            net.zero_grad()
            out = net(rays, True, True, NEAR, FAR) if lat is None else net(rays, True, True, NEAR, FAR, lat)
            loss = ref.helper.img2mse(out[0][0], target) + ref.helper.img2mse(out[1][0], target)
            loss.backward()
            return loss.item()

        return fn, train_fn, "reference", sd, lat

    def fn(rays):
        with torch.no_grad():
            return O.nerf_forward(sd, rays, False, True, NEAR, FAR, latents=lat)

    def train_fn(rays, target):
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        out = O.nerf_forward(p, rays, True, True, NEAR, FAR, latents=lat)
        loss = O.img2mse(out[0][0], target) + O.img2mse(out[1][0], target)
        loss.backward()
        return loss.item()

    return fn, train_fn, "port", sd, lat


def cpu_rays_per_sec(kind: str, budget_s: float, steps: int = 0, warmup: int = 0, keep_output: bool = False):
    """Times the full coarse+fine render (no grad, deterministic, the reference's 3840-ray chunk at most) on samples of the
    bench's rays.  budget mode (steps=0): ~budget_s seconds of work.  step mode: exactly `steps` timed samples."""
    import torch
    from oracle import ref_cpu as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, train_fn, which, sd, lat = cpu_renderer(kind)
    rays = O.sapien_rays(H, W, seed=0)
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(H * W, generator=g)
    last = {}

    def run(n, off):
        idx = perm[off:off + n]
        sub = {k: v[idx].contiguous() for k, v in rays.items()}
        t0 = time.perf_counter()
        out = fn(sub)
        dt = time.perf_counter() - t0
        if keep_output:
            last.update(idx=idx, rays=sub, out=out)
        return dt

    run(256, 0)                                  # warm the thread pool / allocator
    dt = run(512, 256)
    rate = 512 / dt
    res = {"cores": cores, "kind": which}
    if steps:
        per_step_s = max(1.0, min(6.0, 150.0 / max(1, steps + warmup)))
        n = int(min(3840, max(256, rate * per_step_s)))
        for i in range(warmup):
            run(n, (i * n) % (H * W - n))
        t = [run(n, ((warmup + i) * n) % (H * W - n)) for i in range(steps)]
        total = sum(t)
        res.update(value=n * steps / total, ms_per_step=1e3 * total / steps, rays_per_step=n,
                   sample="%d steps x %d random rays of the 640x480 view (the reference's render chunk is 3840 rays), full 65+193 path" % (steps, n))
    else:
        n = int(min(3840, max(256, rate * budget_s / 3)))
        t = [run(n, 1024 + i * n) for i in range(3)]
        med = sorted(t)[1]
        res.update(value=n / med, ms_per_step=1e3 * med, rays_per_step=n,
                   sample="median of 3 x %d random rays of the 640x480 view, full 65+193 path" % n)
    res["_last"] = last
    res["_sd"], res["_lat"], res["_train_fn"], res["_rays"], res["_perm"] = sd, lat, train_fn, rays, perm
    return res


def cpu_train_rays_per_sec(r, kind: str, seconds: float = 12.0):
    """one reference training step (randomized sampling, loss0 + loss1, autograd backward) on a bounded ray batch."""
    import torch
    rays, perm, train_fn = r["_rays"], r["_perm"], r["_train_fn"]
    n = 128
    idx = perm[:n]
    sub = {k: v[idx].contiguous() for k, v in rays.items()}
    tgt = torch.rand(n, 3)
    t0 = time.perf_counter()
    train_fn(sub, tgt)
    rate = n / (time.perf_counter() - t0)
    n = int(min(2048 if kind == "vanilla" else 4096, max(128, rate * seconds / 2)))
    idx = perm[n:2 * n]
    sub = {k: v[idx].contiguous() for k, v in rays.items()}
    tgt = torch.rand(n, 3)
    t0 = time.perf_counter()
    train_fn(sub, tgt)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "rays/s (training step: forward + autograd backward, no optimizer)", "ms_per_step": 1e3 * dt,
            "rays_per_step": n, "cores": r["cores"], "kind": r["kind"],
            "sample": "1 step of %d rays (the reference trains on %d rays per GPU per step)" % (n, 2048 if kind == "vanilla" else 4096)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_rays_per_sec(args.kind, 0.0, steps=args.steps, warmup=args.warmup)
    cfg = workload_config(args, "cpu fp32 (torch, %d threads)" % r["cores"])
    cfg["rays_per_gpu_per_step"] = r["rays_per_step"]
    cfg["reference_arm"] = ("the same workload sampled: each step renders %d random rays of the 640x480 view on the host cores "
                            "(rays are independent; rays/s is per ray)" % r["rays_per_step"])
    line = {"impl": "reference", "metric": "rays/sec", "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_train:
        try:
            line["train"] = cpu_train_rays_per_sec(r, args.kind)
        except Exception as e:
            line["train"] = {"error": "%s: %s" % (type(e).__name__, e)}
    emit(line)


def workload_config(args, precision):
    return {"workload": "sapien single-scene %s 640x480 eval render, 65 coarse + 193 fine samples/ray (the reference's "
                        "'64+128'), 307200 rays per GPU per step, white background, near 2 far 6" % args.kind,
            "rays_per_gpu_per_step": H * W, "samples_per_ray": S0 + S1, "precision": precision,
            "weights": "synthetic xavier-like init, density head sharpened (oracle.make_state_dict(sharp=True))",
            "parallelism": "one 640x480 view per GPU per step on %d GPU(s) (weak scaling); one all-gather of [rays,5] pixels per step" % args.gpus,
            "l2": "256 MiB buffer overwritten between timed iterations (L2 flush)"}


# ---------------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------------
class Scene:
    """weights of one model kind packed for one precision (+ folded latents per articulation id)."""

    def __init__(self, kind_name: str, prec: int, dev):
        import torch
        from types import SimpleNamespace
        from aon_b200 import lib as L
        from aon_b200 import nerf as NF_
        from aon_b200.synth import make_state_dict
        self.kind_name, self.prec, self.dev = kind_name, prec, dev
        self.kind = L.KIND_VANILLA if kind_name == "vanilla" else L.KIND_AUTODECODER
        sd = make_state_dict(kind_name, seed=0, sharp=True)
        self.codes = None
        if kind_name == "vanilla":
            net = NF_.NeRF()
            net.load_state_dict(sd)
        else:
            net = NF_.NeRF_AE_Art()
            net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
            self.codes = NF_.CodeLibraryArticulated(SimpleNamespace(N_max_objs=1, N_obj_code_length=128))
            self.codes.load_state_dict({k[len("code_library."):]: v for k, v in sd.items() if k.startswith("code_library.")})
            self.codes = self.codes.to(dev)
        self.net = net.to(dev).eval()
        self.net.precision = prec
        self.pc = self.net._cache["coarse"].get(self.net.coarse_mlp, prec)
        self.pf = self.net._cache["fine"].get(self.net.fine_mlp, prec)
        self.fc = self.ff = None
        if self.codes is not None:
            self.fold(6)

    def repack(self, prec: int):
        from aon_b200 import lib as L
        lc, lf = self.net.coarse_mlp.linears(), self.net.fine_mlp.linears()
        s = Scene.__new__(Scene)
        s.__dict__.update(self.__dict__)
        s.prec = prec
        s.pc = L.pack_weights(self.kind, prec, [l.weight for l in lc], [l.bias for l in lc])
        s.pf = L.pack_weights(self.kind, prec, [l.weight for l in lf], [l.bias for l in lf])
        if self.codes is not None:
            s.fold(6)
        return s

    def fold(self, articulation_id: int):
        """latents of (instance 0, test-time articulation row `articulation_id`) -> per-call folded biases, both MLPs"""
        import torch
        from aon_b200 import lib as L
        with torch.no_grad():
            lat = self.codes({"instance_id": torch.tensor([0], device=self.dev),
                              "articulation_id": torch.tensor([articulation_id], device=self.dev)}, is_test=True)
        la = (lat["density"].contiguous(), lat["color"].contiguous(), lat["articulation"].contiguous())
        self.fc, self.ff = L.fold_latents(self.kind, self.prec, self.pc, *la), L.fold_latents(self.kind, self.prec, self.pf, *la)

    def image(self, c2w, focal, h, w, ray0=0, R=None, out=None):
        from aon_b200 import lib as L
        return L.render_image(self.kind, self.prec, self.pc, self.pf, self.fc, self.ff, c2w, focal, h, w, NEAR, FAR, True,
                              ray0=ray0, R=R, out=out)[0]


def timed(fn, steps, warmup, flush=None, sync=None):
    """mean ms per call over `steps` calls, CUDA events on the current stream, L2 flushed between calls."""
    import torch
    for _ in range(warmup):
        if flush is not None:
            flush.zero_()
        fn()
    if sync:
        sync()
    torch.cuda.synchronize()
    ev = []
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        ev.append((a, b))
    if sync:
        sync()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / steps


def max_over_ranks(ms, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def train_block(args, dev, world, rank, rays_o, rays_d, gemm="tc"):
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
    s.model.train_gemm = gemm
    s.trainer = SimpleNamespace(global_step=0, is_global_zero=rank == 0)
    opt = s.configure_optimizers()
    torch.manual_seed(1234 + rank)                # ... and a different ray batch / different sampling draws per rank
    idx = torch.randperm(rays_o.shape[0], device=dev)[:Rb]
    batch = {"rays_o": rays_o[idx][None], "rays_d": rays_d[idx][None], "viewdirs": rays_d[idx][None],
             "target": torch.rand(1, Rb, 3, device=dev)}
    if args.kind != "vanilla":
        batch.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([3], device=dev))

    sync = lit.GradSync(list(s.named_parameters()), opt.flat_grad)   # per-MLP slices all-reduced while the backward still runs
    # the whole step (zero_grad .. Adam, with the overlapped per-MLP all-reduces when there are several ranks) is captured once
    # in a CUDA graph and replayed (lit.GraphedStep), like lit.Trainer.fit does
    graphed = None
    if (world == 1 or os.environ.get("AON_TRAIN_GRAPH_NCCL", "1") == "1") and os.environ.get("AON_TRAIN_GRAPH", "1") == "1":
        try:
            graphed = lit.GraphedStep(s, opt, batch, sync if world > 1 else None)
        except Exception as e:
            sys.stderr.write("bench: CUDA-graph capture of the training step failed (%s); eager steps\n" % str(e).splitlines()[0])

    def step(i):
        if graphed is not None:
            graphed(batch)
            return
        opt.zero_grad()
        sync.start()
        loss = s.training_step(batch, i)
        loss.backward()
        opt.grad_scale = sync.finish()
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
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / K
    flop = 3 * Rb * (S0 + S1) * FLOP_PER_SAMPLE[args.kind]
    # HBM roofline of the step (it is bound by operand planes, not by the tensor pipe): algorithmic bytes =
    # every 16-bit activation / gradient plane the design moves, counted once per pass that must touch HBM -- the fused forward
    # WRITES each layer output once (no layer re-reads its input), the dgrad chain reads and writes each delta plane once, every
    # wgrad reads its two operand planes once.  Feature columns per sample (vanilla): forward 2496, dgrad 2432 + 2208,
    # wgrad 5696; auto-decoder: forward 3392, dgrad 3328 + 3504, wgrad 7520 (the fp32 encoding gradient and the 16 / 32-wide
    # position / view planes are left out).
    cols = {"vanilla": 2496 + 2432 + 2208 + 5696, "autodecoder": 3392 + 3328 + 3504 + 7520}[args.kind]
    planes = 2 if gemm == "tc" else 1
    step_bytes = Rb * (S0 + S1) * cols * 2 * planes
    try:
        hbm_peak, hbm_src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs, of measured"
    except Exception:
        hbm_peak, hbm_src = 6400.0, "fallback of B200_PROFILING.md, of fallback"
    roof = {"bound": "hbm", "achieved": step_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": step_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_step": step_bytes, "peak_source": hbm_src,
            "what": "whole training step (all kernels): operand-plane bytes the design must move through HBM / step time"}
    return {"value": world * Rb / (ms * 1e-3), "roofline": roof,
            "forward": "fused: one launch per level (aon_forward_train), each layer output written to HBM once" if getattr(s.model, "train_fwd", "") == "fused" else "one GEMM launch per nn.Linear",
            "sampling_draws": "Philox4x32-10 inside the sampling kernels (no uniform tensors in HBM)", "unit": "rays/s (training: forward + backward + all-reduce + Adam)", "ms_per_step": ms,
            "rays_per_gpu_per_step": Rb, "steps": K, "warmup": 3, "algorithmic_tflops_per_gpu": flop / (ms * 1e-3) / 1e12,
            "gemm": {"tc": "tcgen05 kind::f16, fp16 hi+lo operands (3 MMAs per K step), fp32 accumulate",
                     "tc16": "tcgen05 kind::f16, single fp16 operand planes (1 MMA per K step; fast training mode, ~1e-3 gradient noise), fp32 accumulate"}.get(s.model.train_gemm, "library"),
            "aon_launches_per_step": graphed.launches if graphed is not None else L.launch_count() // K,
            "cuda_graph": graphed is not None,
            "grad_allreduce_bytes": opt.flat_grad.numel() * 4 if world > 1 else 0,
            "grad_allreduce": "%d asynchronous all-reduces per step (fine MLP, coarse MLP%s), launched from post-accumulate hooks while the "
                              "backward runs" % (len(sync.groups), ", code tables" if len(sync.groups) > 2 else "") if world > 1 else "none (1 rank)"}


def sharded_block(scene, focal, c2w, world, rank, dev, flush, steps=5, warmup=2):
    """ONE 640x480 image over the N ranks: contiguous pixel blocks (dist.shard_bounds), each rendered by the fused kernel,
    ONE all-gather (dist.render_image_sharded).  ms per image = max over ranks; the gathered image is compared, bit for
    bit, with the same image rendered by a single GPU (every rank renders the full image once, untimed)."""
    import torch
    import torch.distributed as dist
    from aon_b200 import dist as D
    R = H * W
    block = lambda lo, hi: scene.image(c2w, focal, H, W, ray0=lo, R=hi - lo)
    box = {}

    def one():
        box["img"] = D.render_image_sharded(block, R, device=dev)

    ms = timed(one, steps, warmup, flush, sync=(dist.barrier if world > 1 else None))
    ms = max_over_ranks(ms, world, dev)
    single = scene.image(c2w, focal, H, W)
    equal = bool(torch.equal(box["img"], single))
    eq = torch.tensor([1 if equal else 0], device=dev)
    if world > 1:
        dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    return {"what": "one 640x480 image, pixels sharded contiguously over %d rank(s), one all-gather of [rays,5]" % world,
            "ms_per_image": ms, "value": R / (ms * 1e-3), "unit": "rays/s", "rays": R, "steps": steps, "warmup": warmup,
            "gather_bytes": R * 20, "equals_single_gpu_render": bool(eq.item() == 1), "scaling": "strong"}


def c2_block(prec, prec_name, dev, flush, peak):
    """BASELINE configs[2]: sapien_multi auto-decoder (2-part articulated model) on ONE GPU, at the configured 320x240 and at
    640x480.  Device-resident rays/s, e2e through aon_render_image_host, and the fused kernel's algorithmic roofline."""
    import torch
    from aon_b200 import lib as L
    from aon_b200.synth import sapien_camera, sapien_focal
    scene = Scene("autodecoder", prec, dev)
    out = {"model": "vanilla_autodecoder (NeRF_AE_Art), instance 0, test articulation row 6", "precision": prec_name}
    passes = 3 if prec_name == "f16x3" else 1
    for (h, w) in ((240, 320), (480, 640)):
        focal, c2w = sapien_focal(h), sapien_camera(seed=2)
        R = h * w
        ms = timed(lambda: scene.image(c2w, focal, h, w), 5, 3, flush)
        o, d = L.raygen(h, w, focal, c2w, dev)
        ho, hd = o.cpu().pin_memory(), d.cpu().pin_memory()
        hout = torch.empty(R, 5, dtype=torch.float32).pin_memory()
        host = lambda: L.render_image_host(scene.kind, prec, scene.pc, scene.pf, scene.fc, scene.ff, ho, hd, hd, NEAR, FAR, True, out=hout)
        host()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            host()
        e2e = R * 3 / (time.perf_counter() - t0)
        flop = FLOP_PER_SAMPLE["autodecoder"] * (S0 + S1) * R
        tf = flop / (ms * 1e-3) / 1e12
        out["%dx%d" % (w, h)] = {"value": R / (ms * 1e-3), "unit": "rays/s", "ms_per_image": ms, "rays": R,
                                 "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": 3 * R * 12, "d2h_bytes_per_step": R * 20},
                                 "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                              "mma_passes": passes, "executed_frac": tf * passes / peak,
                                              "algorithmic_flop_per_image": flop, "flop_per_sample": FLOP_PER_SAMPLE["autodecoder"]}}
    return out


def c4_block(prec, prec_name, world, rank, dev, flush):
    """BASELINE configs[4]: sapien_multi articulated eval -- 4 articulation states (rows 0, 6, 12, 18 of the 19-row
    interpolated test table, models/code_library.py:55-71) x one 640x480 view each, every image sharded over the N ranks
    (render_rays_test of model_autodecoder.py:514-541 without its chunk loop).  One pass = 4 x (latent fold, sharded
    render, all-gather)."""
    import torch
    import torch.distributed as dist
    from aon_b200 import dist as D
    from aon_b200.synth import sapien_camera, sapien_focal
    scene = Scene("autodecoder", prec, dev)
    focal = sapien_focal(H)
    cams = [sapien_camera(seed=10 + i) for i in range(len(C4_ARTICULATIONS))]
    R = H * W
    imgs = {}

    def one_pass():
        for i, art in enumerate(C4_ARTICULATIONS):
            scene.fold(art)
            imgs[art] = D.render_image_sharded(lambda lo, hi: scene.image(cams[i], focal, H, W, ray0=lo, R=hi - lo), R, device=dev)

    ms = timed(one_pass, 2, 1, flush, sync=(dist.barrier if world > 1 else None))
    ms = max_over_ranks(ms, world, dev)
    # in-line check: state 12 rendered by this rank alone must equal the gathered image; states must differ from each other
    scene.fold(12)
    single = scene.image(cams[2], focal, H, W)
    eq = torch.tensor([1 if torch.equal(single, imgs[12]) else 0], device=dev)
    if world > 1:
        dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    differ = float((imgs[0][:, :3] - imgs[18][:, :3]).abs().mean().item())
    n = len(C4_ARTICULATIONS)
    return {"what": "vanilla_autodecoder 640x480, %d articulation states %s, each image sharded over %d rank(s) + one all-gather"
                    % (n, list(C4_ARTICULATIONS), world), "precision": prec_name,
            "ms_per_pass": ms, "images_per_s": n / (ms * 1e-3), "value": n * R / (ms * 1e-3), "unit": "rays/s", "rays_per_pass": n * R,
            "steps": 2, "warmup": 1, "equals_single_gpu_render": bool(eq.item() == 1), "mean_abs_rgb_difference_state0_vs_state18": differ,
            "scaling": "strong"}


def parity_block(scene_ctx, cpu):
    """The GPU render (default precision of the bench line) of exactly the rays the CPU leg rendered, against what the CPU
    produced (the reference's own modules when cpu['kind'] == 'reference'): max relative error of rgb / acc / depth at the
    fine level, relative error floor 1e-2 and the BASELINE 4.4 floor 1e-6 both printed, next to the reference's own fp32
    noise floor (its distance from an fp64 evaluation of the same network; first 256 rays)."""
    import torch
    from oracle import ref_cpu as O
    L, kind, prec, pc, pf, fc, ff, dev = scene_ctx
    last = cpu["_last"]
    if not last:
        return None
    rays, want = last["rays"], last["out"]
    o, d, v = (rays[k].to(dev) for k in ("rays_o", "rays_d", "viewdirs"))
    fine, coarse = L.render_rays(kind, prec, pc, pf, fc, ff, o, d, v, NEAR, FAR, True)
    fine = fine.cpu()
    got = (fine[:, :3], fine[:, 3], fine[:, 4])

    def rel(a, b, floor):
        a, b = a.double(), b.double()
        return ((a - b).abs() / b.abs().clamp_min(floor)).max().item()

    n = min(256, o.shape[0])
    sd64 = {k: t.double() for k, t in cpu["_sd"].items()}
    l64 = None if cpu["_lat"] is None else {k: t.double() for k, t in cpu["_lat"].items()}
    with torch.no_grad():
        t64 = O.nerf_forward(sd64, {k: t[:n].double() for k, t in rays.items()}, False, True, NEAR, FAR, latents=l64)
    floor = max(rel(want[1][j][:n], t64[1][j], 1e-2) for j in range(3))
    ours64 = max(rel(got[j][:n], t64[1][j], 1e-2) for j in range(3))
    return {"against": cpu["kind"], "rays": int(o.shape[0]), "level": "fine, end to end (coarse -> sample_pdf -> fine)",
            "e2e_max_rel": max(rel(got[j], want[1][j], 1e-2) for j in range(3)),
            "e2e_max_rel_floor_1e-6": max(rel(got[j], want[1][j], 1e-6) for j in range(3)),
            "reference_fp32_vs_fp64_floor": floor, "ours_vs_fp64": ours64,
            "bar": "max(1e-4, 5 x reference_fp32_vs_fp64_floor) -- tests/test_gpu_parity.py; stage-wise bar 1e-4 in tests/test_gpu_tc.py"}


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


def guarded(name, fn, *a):
    """an extra block never costs the headline line"""
    try:
        return fn(*a)
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from aon_b200 import lib as L
    from aon_b200.synth import sapien_camera

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
    tensor_mode = prec != L.PREC_FP32

    # ---- synthetic scene: weights + this rank's view (rays are generated inside the fused kernel) -----------------
    scene = Scene(args.kind, prec, dev)
    focal = 0.5 * H / math.tan(math.radians(17.5))
    c2w = sapien_camera(seed=rank)
    R = H * W
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world, R, 5, dtype=torch.float32, device=dev) if world > 1 else None
    pix = torch.empty(R, 5, dtype=torch.float32, device=dev)

    def step():
        """one full render of this rank's 307 200 rays, device resident: ONE fused launch (+ the sample-segmented tail)."""
        scene.image(c2w, focal, H, W, out=pix)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pix)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    L.launch_count(reset=True)
    step_ev = []
    with ClockSampler(local) as clk:
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            step_ev.append((a, b))
        barrier()
    launches = L.launch_count()
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in step_ev), world, dev)
    value = world * R * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel: a launch of exactly 16 full waves of CTA pairs = the fused kernel alone -------------
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    R_k = (R // (sms // 2 * 256)) * (sms // 2 * 256) if tensor_mode else R
    kern_ms = timed(lambda: scene.image(c2w, focal, H, W, ray0=0, R=R_k), 3, 1, flush)

    # ---- e2e: host rays -> host pixels through the C-ABI call ------------------------------------------
    rays_o, rays_d = L.raygen(H, W, focal, c2w, dev)
    ho = rays_o.cpu().pin_memory(); hd = rays_d.cpu().pin_memory()
    hout = torch.empty(R, 5, dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    host = lambda: L.render_image_host(kind, prec, scene.pc, scene.pf, scene.fc, scene.ff, ho, hd, hd, NEAR, FAR, True, out=hout)
    host()
    barrier()
    t_e2e = time.perf_counter()
    for _ in range(e2e_steps):
        host()   # synchronises
    torch.cuda.synchronize()
    e2e_value = world * R * e2e_steps / max_over_ranks(time.perf_counter() - t_e2e, world, dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if tensor_mode:
        peak, peak_src = peaks.get("bf16_tflops_sustained", 1400.0), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured"
        if "bf16_tflops_sustained" not in peaks:
            peak_src = "fallback of B200_PROFILING.md (no MEASURED_PEAKS.json): 1.4 PFLOP/s sustained, of fallback"
    else:
        peak, peak_src = 148 * 128 * 2 * 1.965e-3, "fp32 FFMA peak 148 SM x 128 lanes x 2 x 1.965 GHz (CUDA-core mode; not a tensor-pipe number)"

    # ---- extra blocks (BASELINE configs[2..4], the sharded single image, the one-pass mode) ---------------------------
    fast = sharded = c2 = c4 = train = None
    if not args.no_extras and tensor_mode:
        sharded = guarded("sharded", sharded_block, scene, focal, sapien_camera(seed=0), world, rank, dev, flush)
        if world == 1 and prec_name == "f16x3":
            def fast_block():
                fs = scene.repack(L.PREC_TC_F16)
                ms = timed(lambda: fs.image(c2w, focal, H, W), 3, 1, flush)
                kms = timed(lambda: fs.image(c2w, focal, H, W, ray0=0, R=R_k), 3, 1, flush)
                tf = FLOP_PER_SAMPLE[args.kind] * (S0 + S1) * R_k / (kms * 1e-3) / 1e12
                return {"precision": "f16 (single tcgen05 pass; not the parity mode)", "value": R / (ms * 1e-3), "unit": "rays/s",
                        "kernel_ms": kms, "achieved_tflops": tf, "frac": tf / peak}
            fast = guarded("fast", fast_block)
        if world == 1 and args.kind == "vanilla":
            c2 = guarded("c2", c2_block, prec, prec_name, dev, flush, peak)
        c4 = guarded("c4", c4_block, prec, prec_name, world, rank, dev, flush)
    train_fast = None
    if not args.no_train:
        train = guarded("train", train_block, args, dev, world, rank, rays_o, rays_d)
        if not args.no_extras:
            train_fast = guarded("train_fast", train_block, args, dev, world, rank, rays_o, rays_d, "tc16")

    if rank == 0:
        flop_k = FLOP_PER_SAMPLE[args.kind] * (S0 + S1) * R_k
        achieved = flop_k / (kern_ms * 1e-3) / 1e12
        passes = 3 if prec_name == "f16x3" else 1
        traffic = None
        try:   # per-launch DRAM bytes of the fused kernel from the committed ncu --set full capture of this command
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_fused_%s_%s.json" % (args.kind, prec_name))))["dram_bytes_per_launch"]
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
            "roofline": {"bound": "tensor", "kernel": "render_tc_kernel, fused image kernel (raygen + coarse 65 + sample_pdf + fine 193 samples/ray)"
                         if tensor_mode else "render_simt_kernel (three-launch path)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": 20 * R_k,
                         "mma_passes": passes, "executed_tflops": achieved * passes, "executed_frac": achieved * passes / peak,
                         "peak_source": peak_src, "kernel_ms": kern_ms, "rays_per_launch": R_k,
                         "algorithmic_flop_per_launch": flop_k,
                         "step_share": {"fused_kernel_ms_scaled_to_full_image": kern_ms * R / R_k, "step_ms": ms_total / args.steps}},
            "clocks": clk.report(),
        }
        for k, v in (("fast_mode", fast), ("sharded", sharded), ("c2", c2), ("c4", c4), ("train", train), ("train_fast", train_fast)):
            if v is not None:
                line[k] = v
        if not args.no_cpu_baseline and world == 1:
            r = cpu_rays_per_sec(args.kind, args.cpu_seconds, keep_output=True)
            line["cpu_baseline"] = {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            line["parity"] = guarded("parity", parity_block, (L, kind, prec, scene.pc, scene.pf, scene.fc, scene.ff, dev), r)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
