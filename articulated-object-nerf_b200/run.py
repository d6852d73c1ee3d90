"""Thin re-host of the reference's run.py for the hot-path experiment types (run.py:20-171, opt.py):

    python -m aon_b200.run --config config/nerf_training.json [--exp_type vanilla] [--run_eval] ...

JSON keys override flags attribute by attribute like opt.py:212-222.  Only the flags the render path reads are kept
(grep of hparams.* uses): dataset_name, root_dir, img_wh, white_back, chunk, num_gpus, run_eval,
render_name, exp_name, exp_type, run_max_steps, N_max_objs, N_obj_code_length, output_path, ckpt_path.
Training drives the Lightning-surface module with the minimal trainer (lit.Trainer) and saves a PL-style checkpoint
{"state_dict": ...} to {output_path}/{exp_name}/last.ckpt; --run_eval loads it (run.py:156-163), renders the test split
and writes ckpts/{exp_name}/{render_name}/imageNNN.jpg + results.json (model.py:459-507)."""
from __future__ import annotations

import argparse
import json
import os

import torch

from . import data, lit


def get_opts(argv=None):
    """The flags the render path reads, with the reference's defaults (opt.py); every other reference flag (the README
    commands pass --batch_size, --num_epochs, --lr ...) is accepted and ignored -- the hot path does not read them."""
    p = argparse.ArgumentParser()
    p.add_argument("--config", default=None)
    p.add_argument("--exp_type", default="vanilla", choices=["vanilla", "vanilla_autodecoder"])
    p.add_argument("--dataset_name", default="sapien")
    p.add_argument("--root_dir", default=None)
    p.add_argument("--exp_name", default="exp")
    p.add_argument("--render_name", default="render")
    p.add_argument("--output_path", default="./results")
    p.add_argument("--ckpt_path", default=None)
    p.add_argument("--img_wh", nargs=2, type=int, default=[640, 480])
    p.add_argument("--white_back", action="store_true", default=False)                 # opt.py: default False
    p.add_argument("--chunk", type=int, default=16 * 240)
    p.add_argument("--num_gpus", type=int, default=1)
    p.add_argument("--run_max_steps", type=int, default=100000)
    p.add_argument("--run_eval", action="store_true", default=False)
    p.add_argument("--N_max_objs", type=int, default=151)                              # opt.py default
    p.add_argument("--N_obj_code_length", type=int, default=128)
    p.add_argument("--seed", type=int, default=0, help="base seed of the ray sampler; rank r draws with seed + r")
    p.add_argument("--ckpt_every", type=int, default=5000, help="steps between checkpoints (0 = only at the end)")
    p.add_argument("--precision", default=None, help="fp32 | f16x3 | f16 | bf16 (default: AON_PRECISION or fp32)")
    a, ignored = p.parse_known_args(argv)
    a.ignored_flags = ignored
    if a.config:
        with open(a.config) as f:
            for k, v in json.load(f).items():
                setattr(a, k, v)
    return a


def _init_distributed(dev):
    """One process per GPU under torchrun (run.py:109-111,151-153: PL DDPPlugin): join the NCCL group when WORLD_SIZE > 1."""
    import torch.distributed as dist
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if dev.type == "cuda":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def save_checkpoint(path, system, step):
    """PL-style checkpoint + what a resumed run needs: Adam moments / step count (lit.FlatAdam.state_dict) and global_step.
    Written by rank 0 only, atomically (tmp + rename)."""
    opt = getattr(system, "_optimizer", None)
    blob = {"state_dict": system.state_dict(), "global_step": int(step),
            "optimizer_states": [opt.state_dict()] if opt is not None else []}
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, path)


def main(hparams):
    if hparams.dataset_name not in ("sapien", "sapien_multi"):
        raise SystemExit("dataset_name must be 'sapien' or 'sapien_multi' (datasets/__init__.py:4)")
    multi = hparams.dataset_name == "sapien_multi"
    Dataset = data.SapienDatasetMulti if multi else data.SapienDataset
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    rank, world = _init_distributed(dev)
    result = os.path.join(hparams.output_path, hparams.exp_name)
    os.makedirs(result, exist_ok=True)
    ckpt = os.path.join(result, hparams.ckpt_path or "last.ckpt")
    blob = torch.load(ckpt, map_location=dev) if (hparams.run_eval or (getattr(hparams, "resume", True) and os.path.exists(ckpt))) else None
    if blob is not None and "code_library.embedding_instance_shape.weight" in blob["state_dict"]:
        # the code tables' row count comes from the checkpoint (a reference checkpoint carries N_max_objs = 151 rows)
        hparams.N_max_objs = blob["state_dict"]["code_library.embedding_instance_shape.weight"].shape[0]
    system = lit.build_system(hparams).to(dev)
    if hparams.precision:
        from . import lib
        system.model.precision = lib.PRECISIONS[hparams.precision]
    if hparams.run_eval:
        system.load_state_dict(blob["state_dict"])
        test = Dataset(hparams.root_dir, "test_val", tuple(hparams.img_wh), white_back=hparams.white_back,
                                  eval_inference=hparams.render_name, device=dev)
        system.setup(datasets={"test": test})
        tr = lit.Trainer()
        keep = ("instance_id", "articulation_id")
        # every rank renders whole images: rank r takes images r, r + world, ...
        idx = range(rank, len(test), world)
        tr.test(system, ({k: (v[None] if torch.is_tensor(v) and k not in keep else v) for k, v in test[i].items()} for i in idx))
        print("test", {k: round(v, 4) for k, v in system.logged.items() if k.startswith("test/")})
        return system
    train = Dataset(hparams.root_dir, "train", tuple(hparams.img_wh), white_back=hparams.white_back, device=dev,
                    **({"seed": hparams.seed + rank} if multi else {}))
    system.setup(datasets={"train": train})
    trainer = lit.Trainer(max_steps=hparams.run_max_steps, log_every=max(1, hparams.run_max_steps // 10))
    if blob is not None:                       # resume: weights, Adam moments, step count (LR schedule + bias correction)
        system.load_state_dict(blob["state_dict"])
        trainer.resume(system, blob)
    if rank == 0:
        trainer.on_checkpoint = lambda step: save_checkpoint(ckpt, system, step)
        trainer.ckpt_every = int(getattr(hparams, "ckpt_every", 0) or 0)
    # DDP semantics: identical initial weights (Trainer.fit broadcasts rank 0's), a different ray stream per rank
    batches = train.ray_batches() if multi else train.ray_batches(2048, seed=hparams.seed + rank)
    trainer.fit(system, batches)
    if rank == 0:
        save_checkpoint(ckpt, system, trainer.global_step)
    if world > 1:
        torch.distributed.barrier()
    return system


if __name__ == "__main__":
    main(get_opts())
