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
    p.add_argument("--white_back", action="store_true", default=True)
    p.add_argument("--chunk", type=int, default=16 * 240)
    p.add_argument("--num_gpus", type=int, default=1)
    p.add_argument("--run_max_steps", type=int, default=100000)
    p.add_argument("--run_eval", action="store_true", default=False)
    p.add_argument("--N_max_objs", type=int, default=1)
    p.add_argument("--N_obj_code_length", type=int, default=128)
    p.add_argument("--precision", default=None, help="fp32 | f16x3 | f16 | bf16 (default: AON_PRECISION or fp32)")
    a = p.parse_args(argv)
    if a.config:
        with open(a.config) as f:
            for k, v in json.load(f).items():
                setattr(a, k, v)
    return a


def main(hparams):
    if hparams.dataset_name not in ("sapien", "sapien_multi"):
        raise SystemExit("dataset_name must be 'sapien' or 'sapien_multi' (datasets/__init__.py:4)")
    multi = hparams.dataset_name == "sapien_multi"
    Dataset = data.SapienDatasetMulti if multi else data.SapienDataset
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    system = lit.build_system(hparams).to(dev)
    if hparams.precision:
        from . import lib
        system.model.precision = lib.PRECISIONS[hparams.precision]
    result = os.path.join(hparams.output_path, hparams.exp_name)
    os.makedirs(result, exist_ok=True)
    ckpt = os.path.join(result, hparams.ckpt_path or "last.ckpt")
    if hparams.run_eval:
        system.load_state_dict(torch.load(ckpt, map_location=dev)["state_dict"])
        test = Dataset(hparams.root_dir, "test_val", tuple(hparams.img_wh), white_back=hparams.white_back,
                                  eval_inference=hparams.render_name, device=dev)
        system.setup(datasets={"test": test})
        tr = lit.Trainer()
        keep = ("instance_id", "articulation_id")
        tr.test(system, ({k: (v[None] if torch.is_tensor(v) and k not in keep else v) for k, v in test[i].items()} for i in range(len(test))))
        print("test", {k: round(v, 4) for k, v in system.logged.items() if k.startswith("test/")})
        return system
    train = Dataset(hparams.root_dir, "train", tuple(hparams.img_wh), white_back=hparams.white_back, device=dev)
    system.setup(datasets={"train": train})
    lit.Trainer(max_steps=hparams.run_max_steps, log_every=max(1, hparams.run_max_steps // 10)).fit(system, train.ray_batches() if multi else train.ray_batches(2048))
    torch.save({"state_dict": system.state_dict(), "global_step": hparams.run_max_steps}, ckpt)
    return system


if __name__ == "__main__":
    main(get_opts())
