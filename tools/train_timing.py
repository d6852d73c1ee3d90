"""Training-step timing on one GPU (BASELINE.json configs[3] per-GPU share: 2048-ray batches, vanilla; 4096 for sapien_multi).

    python tools/train_timing.py [--steps 20]

One step = lit.Trainer's loop body: zero_grad, training_step (native sampling + pos_enc + compositing with their hand-written
adjoints, MLP contractions as library GEMMs under autograd), backward, optimizer_step (LR schedule + one aon_adam_step).
Prints ms/step, rays/s and the share of the three native stages."""
import argparse
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aon_b200 import lib as L, lit, synth

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
dev = torch.device("cuda:0")
ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph replay of lit.GraphedStep")
ap.add_argument("--only", default="", help="exp:gemm, e.g. vanilla:tc -- run one configuration only (profiling)")
args = ap.parse_args()
CONFIGS = (("vanilla", 2048, "tc"), ("vanilla_autodecoder", 4096, "tc"), ("vanilla", 2048, "tc16"), ("vanilla_autodecoder", 4096, "tc16"))
if args.only:
    CONFIGS = tuple(c for c in CONFIGS if "%s:%s" % (c[0], c[2]) == args.only)
for exp, R, gemm in CONFIGS:
    torch.manual_seed(0)
    s = lit.build_system(SimpleNamespace(exp_type=exp, run_max_steps=1000, white_back=True, N_max_objs=1, N_obj_code_length=128)).to(dev)
    s.train()
    s.model.train_gemm = gemm            # "tc": fp16 hi+lo operand planes (fp32-grade); "tc16": single fp16 planes (fast mode)
    o, d = L.raygen(480, 640, synth.sapien_focal(480), synth.sapien_camera(0), dev)
    idx = torch.randperm(o.shape[0], device=dev)[:R]
    batch = {"rays_o": o[idx][None], "rays_d": d[idx][None], "viewdirs": d[idx][None], "target": torch.rand(1, R, 3, device=dev)}
    if exp != "vanilla":
        batch.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([3], device=dev))
    opt = s.configure_optimizers()
    s.trainer = SimpleNamespace(global_step=0, is_global_zero=True)

    graphed = None if args.no_graph else lit.GraphedStep(s, opt, batch, None)

    def step(i):
        if graphed is not None:
            return graphed(batch)
        opt.zero_grad()
        loss = s.training_step(batch, i)
        loss.backward()
        s.optimizer_step(0, i, opt, 0, None, False, False, False)
        s.trainer.global_step += 1
        return loss

    for i in range(3):
        step(i)
    L.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    import time
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for i in range(2):               # 2 steps = ~600 launches: below the launch queue depth, so this is pure host cost
        step(i)
    t_host2 = (time.perf_counter() - t2) / 2 * 1e3
    torch.cuda.synchronize()
    e0.record()
    t_host = time.perf_counter()
    for i in range(args.steps):
        loss = step(i)
    t_host = (time.perf_counter() - t_host) / args.steps * 1e3      # host time to ENQUEUE a step (no sync inside the loop ...
    e1.record()                                                      # ... except the float() of the logged scalars)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flop = 3 * R * 258 * (1186816 if exp == "vanilla" else 1589760)
    print("%-20s %-4s %-5s %5d rays/step: %.2f ms/step -> %.0f rays/s training, %.1f algorithmic TFLOP/s (fwd+bwd = 3x fwd), "
          "%d aon kernel launches/step, host enqueue %.2f ms/step (%.2f ms/step with an empty launch queue), loss %.4f" % (exp, gemm, "eager" if graphed is None else "graph", R, ms, R / ms * 1e3, flop / ms * 1e-9, L.launch_count() // (args.steps + 2), t_host, t_host2, loss.item()))
