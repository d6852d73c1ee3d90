"""Debug aid: eager vs CUDA-graph-replayed training steps (lit.GraphedStep) must leave identical parameters."""
import sys, os, traceback
sys.path.insert(0, os.getcwd())
from types import SimpleNamespace
import torch
from oracle import ref_cpu as O
from aon_b200 import lit
dev = torch.device("cuda:0")
rays = {k: v.to(dev) for k, v in O.sapien_rays(24, 32, seed=5).items()}
g = torch.Generator().manual_seed(1)
target = torch.rand(rays["rays_o"].shape[0], 3, generator=g).to(dev)
def batches(n, exp):
    for i in range(n):
        sl = slice(128 * i, 128 * (i + 1))
        b = {k: v[sl][None] for k, v in rays.items()}
        b["target"] = target[sl][None]
        if exp != "vanilla":
            b.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([i % 10], device=dev))
        yield b
def run(graph, n, exp):
    torch.manual_seed(0)
    s = lit.build_system(SimpleNamespace(exp_type=exp, run_max_steps=200, white_back=True, N_max_objs=1, N_obj_code_length=128)).to(dev)
    s.randomized = False
    s.lr_delay_steps = 4
    if graph and exp != "vanilla":      # show the capture error in full
        opt = s.configure_optimizers(); s._optimizer = opt
        try:
            b = next(batches(1, exp))
            lit.GraphedStep(s, opt, b, None)
            print("capture ok", exp)
        except Exception:
            traceback.print_exc()
        return None
    tr = lit.Trainer(max_steps=n, cuda_graph=graph)
    tr.fit(s, batches(n, exp))
    o = s._optimizer
    return o.flat.clone(), o.flat_grad.clone(), o.exp_avg.clone(), s.logged["train/loss"]
for n in (3, 6):
    a = run(False, n, "vanilla"); c = run(True, n, "vanilla")
    print(n, "eager-graph params %.3e grad %.3e m %.3e loss %r %r" % ((a[0]-c[0]).abs().max().item(), (a[1]-c[1]).abs().max().item(), (a[2]-c[2]).abs().max().item(), a[3], c[3]))
run(True, 1, "vanilla_autodecoder")
