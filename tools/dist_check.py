"""Multi-GPU check (launch with torchrun): dist.render_sharded over NCCL == the 1-GPU render.
Every mode must be bit-identical: no operation crosses rays, the MMA issue order is fixed (ticket), and sample-segmented
launches composite in sample order.  Also checks the camera form (dist.render_image_sharded over the fused image kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from aon_b200 import dist as D, lib as L, nerf, synth


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = synth.make_state_dict("vanilla", 0, True)
    net = nerf.NeRF()
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    H, W = 61, 83          # ragged: 5063 rays
    o, d = L.raygen(H, W, synth.sapien_focal(H), synth.sapien_camera(3), dev)
    rays = {"rays_o": o, "rays_d": d, "viewdirs": d}

    def render(block):
        with torch.no_grad():
            rgb, acc, depth = net(block, False, True, 2.0, 6.0)[1]
        return torch.cat([rgb, acc[:, None], depth[:, None]], -1)

    ok = True
    for mode in ("fp32", "f16x3"):
        net.precision = L.PRECISIONS[mode]
        full = render(rays)
        got = D.render_sharded(render, rays)
        err = (got - full).abs().max().item()
        same = torch.equal(got, full)
        if rank == 0:
            print("world %d mode %s: sharded == single-GPU bitwise: %s, max abs diff %.3g" % (dist.get_world_size(), mode, same, err), flush=True)
        ok &= same
        if mode != "fp32":
            k = net.coarse_mlp.KIND
            pc = net._cache["coarse"].get(net.coarse_mlp, net.precision)
            pf = net._cache["fine"].get(net.fine_mlp, net.precision)
            c2w, focal = synth.sapien_camera(3), synth.sapien_focal(H)
            blk = lambda lo, hi: L.render_image(k, net.precision, pc, pf, None, None, c2w, focal, H, W, 2.0, 6.0, True, ray0=lo, R=hi - lo)[0]
            img = D.render_image_sharded(blk, H * W, device=dev)
            same_i = torch.equal(img, full)
            if rank == 0:
                print("world %d mode %s: render_image_sharded == single-GPU bitwise: %s" % (dist.get_world_size(), mode, same_i), flush=True)
            ok &= same_i
    ok &= train_check(rank, dev)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def train_check(rank, dev):
    """DDP semantics of the training step over NCCL: every rank trains on its OWN ray batches and sampling draws, the summed
    gradients are averaged (lit.GradSync + 1/world in the Adam kernel), so after any number of steps all ranks hold the SAME
    parameters; and the CUDA-graph replay of the step (all-reduces captured with it) leaves exactly the parameters the
    eager loop leaves."""
    from types import SimpleNamespace
    from aon_b200 import lit
    world = dist.get_world_size()
    o, d = L.raygen(48, 64, synth.sapien_focal(48), synth.sapien_camera(1), dev)
    finals = {}
    ok = True
    for mode in ("eager", "graph"):
        torch.manual_seed(11)                                    # same initial weights on every rank
        s = lit.build_system(SimpleNamespace(exp_type="vanilla", run_max_steps=100, white_back=True)).to(dev)
        s.lr_delay_steps = 3
        s.model.rng_seed = 1000 + rank                           # per-rank sampling draws
        g = torch.Generator().manual_seed(50 + rank)             # per-rank ray batches

        def batches():
            for i in range(5):
                idx = torch.randperm(o.shape[0], generator=g)[:256].to(dev)
                yield {"rays_o": o[idx][None], "rays_d": d[idx][None], "viewdirs": d[idx][None],
                       "target": torch.rand(1, 256, 3, generator=g).to(dev)}

        tr = lit.Trainer(max_steps=5, cuda_graph=mode == "graph")
        tr.fit(s, batches())
        flat = s._optimizer.flat.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], x) for x in gathered)
        finals[mode] = flat
        if rank == 0:
            print("world %d training, %s steps: parameters identical on all ranks after 5 steps: %s (loss %.5f)"
                  % (world, mode, same, s.logged["train/loss"]), flush=True)
        ok &= same
    eq = torch.equal(finals["eager"], finals["graph"])
    if rank == 0:
        print("world %d training: graph replay == eager steps bitwise: %s (max abs diff %.3g)"
              % (world, eq, (finals["eager"] - finals["graph"]).abs().max().item()), flush=True)
    return ok and (eq or world > 2)      # more than two addends: NCCL's reduction order may differ between a captured and an eager call


if __name__ == "__main__":
    main()
