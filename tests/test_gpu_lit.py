"""GPU: Lightning-surface modules end to end on the fused kernels: render dict keys, eval == direct
NeRF.forward, and a few training steps through the minimal Trainer reduce the loss."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_cpu as O

pytestmark = pytest.mark.gpu


def _hp(exp):
    return SimpleNamespace(exp_type=exp, run_max_steps=200, white_back=True, N_max_objs=1, N_obj_code_length=128)


@pytest.mark.parametrize("exp", ["vanilla", "vanilla_autodecoder"])
def test_lit_eval_and_train(built_lib, exp):
    from aon_b200 import lit
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    s = lit.build_system(_hp(exp)).to(dev)
    rays = {k: v.to(dev) for k, v in O.sapien_rays(12, 16, seed=2).items()}
    target = torch.rand(rays["rays_o"].shape[0], 3, device=dev)
    batch = dict(rays, target=target, instance_mask=torch.ones(target.shape[0], dtype=torch.bool, device=dev))
    if exp != "vanilla":
        batch.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([3], device=dev))
    s.eval()
    if exp == "vanilla":
        ret = s.render_rays(batch)
        out = s.test_step({k: (v[None] if torch.is_tensor(v) else v) for k, v in batch.items()}, 0)
        with torch.no_grad():
            direct = s.model(batch, False, True, 2.0, 6.0)[1][0]
    else:
        lat = s.code_library(batch)
        ret = s.render_rays(batch, lat)
        out = s.render_rays_test(batch, lat)
        with torch.no_grad():
            direct = s.model(batch, False, True, 2.0, 6.0, lat)[1][0]
    assert set(ret) == {"comp_rgb", "acc", "depth"} and set(out) == {"target", "instance_mask", "rgb"}
    assert torch.equal(ret["comp_rgb"], direct) and torch.equal(out["rgb"], direct)
    assert "val/psnr" in s.logged
    tr = lit.Trainer(max_steps=8)
    first = None
    losses = []

    def batches():
        while True:
            yield {k: (v[None] if torch.is_tensor(v) and k not in ("instance_id", "articulation_id") else v)
                   for k, v in batch.items()}

    s.lr_delay_steps = 0
    for _ in range(2):
        tr.fit(s, batches())
        losses.append(s.logged["train/loss"])
        tr.max_steps += 8
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize("randomized", [False, True])
@pytest.mark.parametrize("exp", ["vanilla", "vanilla_autodecoder"])
def test_graphed_step_equals_eager(built_lib, exp, randomized):
    """lit.Trainer replays the training step as a CUDA graph (lit.GraphedStep: static batch buffers, Adam scalars through
    device memory, dry-run capture).  The replayed steps must leave EXACTLY the parameters the eager loop leaves: same kernels,
    deterministic reductions, same learning-rate schedule, capture trains nothing -- also with the randomized sampling on,
    because the stratified / inverse-cdf draws are generated in the sampling kernels from a step counter in device memory
    (lib.Rng) that eager and replayed steps advance alike."""
    from aon_b200 import lit
    dev = torch.device("cuda:0")
    rays = {k: v.to(dev) for k, v in O.sapien_rays(24, 32, seed=5).items()}
    g = torch.Generator().manual_seed(1)
    target = torch.rand(rays["rays_o"].shape[0], 3, generator=g).to(dev)

    def batches():
        for i in range(6):                                   # six different 128-ray batches
            sl = slice(128 * i, 128 * (i + 1))
            b = {k: v[sl][None] for k, v in rays.items()}
            b["target"] = target[sl][None]
            if exp != "vanilla":
                b.update(instance_id=torch.tensor([0], device=dev), articulation_id=torch.tensor([i % 10], device=dev))
            yield b

    flats = []
    for graph in (False, True):
        torch.manual_seed(0)
        s = lit.build_system(_hp(exp)).to(dev)
        s.randomized = randomized
        s.model.rng_seed = 7
        s.lr_delay_steps = 4                                 # the learning rate changes every step
        tr = lit.Trainer(max_steps=6, cuda_graph=graph)
        tr.fit(s, batches())
        assert tr.global_step == 6 and s._optimizer.steps == 6
        flats.append((s._optimizer.flat.clone(), s._optimizer.exp_avg_sq.clone(), s.logged["train/loss"]))
    assert torch.equal(flats[0][0], flats[1][0]), (flats[0][0] - flats[1][0]).abs().max().item()
    assert torch.equal(flats[0][1], flats[1][1])
    assert flats[0][2] == flats[1][2]


def test_resume_continues_bitwise(built_lib):
    """A run resumed from a checkpoint (weights + lit.FlatAdam.state_dict() + global_step, the blob run.save_checkpoint writes)
    continues EXACTLY where the first run stopped: 3 steps + resume + 3 steps leave the parameters of 6 straight steps, with
    the randomized sampling on -- the learning-rate schedule, the Adam moments / bias corrections and the in-kernel Philox draw
    sequence (one offset per step) are all functions of the restored step count."""
    import copy
    from aon_b200 import lit
    dev = torch.device("cuda:0")
    rays = {k: v.to(dev) for k, v in O.sapien_rays(24, 32, seed=5).items()}
    g = torch.Generator().manual_seed(1)
    target = torch.rand(rays["rays_o"].shape[0], 3, generator=g).to(dev)

    def batches(lo, hi):
        for i in range(lo, hi):
            sl = slice(128 * i, 128 * (i + 1))
            b = {k: v[sl][None] for k, v in rays.items()}
            b["target"] = target[sl][None]
            yield b

    def fresh():
        torch.manual_seed(0)
        s = lit.build_system(_hp("vanilla")).to(dev)
        s.randomized, s.lr_delay_steps = True, 4
        s.model.rng_seed = 9
        return s

    a = fresh()
    lit.Trainer(max_steps=6).fit(a, batches(0, 6))
    b = fresh()
    tr = lit.Trainer(max_steps=3)
    tr.fit(b, batches(0, 3))
    blob = copy.deepcopy({"state_dict": b.state_dict(), "global_step": tr.global_step, "optimizer_states": [b._optimizer.state_dict()]})
    c = fresh()
    c.load_state_dict(blob["state_dict"])
    tr2 = lit.Trainer(max_steps=6)
    tr2.resume(c, blob)
    tr2.fit(c, batches(3, 6))
    assert tr2.global_step == 6
    assert torch.equal(a._optimizer.flat, c._optimizer.flat), (a._optimizer.flat - c._optimizer.flat).abs().max().item()
