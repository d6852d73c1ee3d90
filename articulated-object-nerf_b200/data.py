"""SAPIEN single-scene dataset in the reference's on-disk format + a synthetic scene writer.

* ``SapienDataset``  <- datasets/sapien.py:11-157.  Same layout and semantics:
  ``{root}/{train,val,test}/rgb/r_{i}.png`` (RGBA, alpha-blended onto white, sapien.py:97-99) and
  ``{root}/{split}/transforms.json`` = ``{"focal": f | "camera_angle_x": a, "frames": {"r_i": 4x4 c2w}}``;
  focal rule of sapien.py:62-69; ``near = 2.0, far = 6.0`` (sapien.py:72-73); val/test files sorted by index
  (sapien.py:46-47); sample dict keys ``rays_o, rays_d, viewdirs, target, instance_mask``.
  Differences, all result-preserving: rays come from the CUDA ray-generation kernel (aon_raygen: A1+A2 of the hot
  path; ``rays_d`` is unit-norm and equals ``viewdirs`` exactly as in the reference, ray_utils.py:146-147) and the
  whole training set is ONE GPU-resident ray table sampled on the GPU (``ray_batches``) instead of a CPU
  ``DataLoader`` over per-ray samples (the reference pre-expands 14 floats per ray on the host).
* ``write_synthetic_scene``: an analytic two-sphere + slab scene ray-traced into that format with the datagen camera
  distribution (datagen/data_utils.py:66-80; fovy 35 deg, data_gen.py:64), because there is no SAPIEN simulator and no
  network in the image.  Used by the PSNR protocol (tools/psnr_protocol.py) and the tests.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Iterator, Optional, Tuple

import numpy as np
import torch

from . import lib as L
from .synth import sapien_camera, sapien_focal

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------------------------------
def _load_rgba(path: str, img_wh: Tuple[int, int]) -> Tensor:
    """PIL open + LANCZOS resize + ToTensor (x/255) as sapien.py:94-99; returns [H*W,4] float32."""
    from PIL import Image
    img = Image.open(path)
    if img.size != tuple(img_wh):
        img = img.resize(tuple(img_wh), Image.LANCZOS)
    a = np.asarray(img.convert("RGBA"), dtype=np.float32) / 255.0
    return torch.from_numpy(a.reshape(-1, 4))


class SapienDataset:
    def __init__(self, root_dir: str, split: str = "train", img_wh: Tuple[int, int] = (320, 240), model_type=None,
                 white_back: Optional[bool] = True, eval_inference=None, device="cuda"):
        self.root_dir, self.split, self.img_wh, self.white_back = root_dir, split, tuple(img_wh), white_back
        self.device = torch.device(device)
        sub = split if split in ("train", "val") else "test"          # 'test_val' etc. fall to test/ (sapien.py:51-58)
        self.base_dir = os.path.join(root_dir, sub)
        with open(os.path.join(self.base_dir, "transforms.json")) as f:
            self.meta = json.load(f)
        files = [x for x in os.listdir(os.path.join(self.base_dir, "rgb")) if x.endswith(".png")]
        self.img_files = sorted(files, key=lambda fn: int(fn.split("_")[1].split(".")[0]))
        w, h = self.img_wh
        if self.meta.get("camera_angle_x", False):
            self.focal = 0.5 * h / np.tan(0.5 * self.meta["camera_angle_x"]) * (w / 320)
        else:
            self.focal = self.meta.get("focal", None)
            if self.focal is None:
                raise ValueError("focal length not found in transforms.json")
        self.near, self.far = 2.0, 6.0
        self.image_sizes = np.array([[h, w] for _ in range(len(self.img_files) if eval_inference is not None else 1)])
        if split == "train":
            rays_o, rays_d, rgbs = [], [], []
            for fn in self.img_files:
                s = self._image_sample(fn)
                rays_o.append(s["rays_o"]); rays_d.append(s["rays_d"]); rgbs.append(s["target"])
            self.all_rays_o, self.all_rays_d, self.all_rgbs = torch.cat(rays_o), torch.cat(rays_d), torch.cat(rgbs)

    def _image_sample(self, fn: str) -> Dict[str, Tensor]:
        w, h = self.img_wh
        c2w = torch.tensor(self.meta["frames"][fn.split(".")[0]], dtype=torch.float32)[:3, :4]
        rgba = _load_rgba(os.path.join(self.base_dir, "rgb", fn), self.img_wh).to(self.device)
        target = rgba[:, :3] * rgba[:, 3:] + (1 - rgba[:, 3:])
        o, d = L.raygen(h, w, float(self.focal), c2w, self.device)
        return {"rays_o": o, "rays_d": d, "viewdirs": d, "instance_mask": rgba[:, 3] > 0, "target": target.contiguous()}

    def __len__(self):
        if self.split == "train":
            return self.all_rays_o.shape[0]
        return 1 if self.split == "val" else len(self.img_files)

    def __getitem__(self, idx) -> Dict[str, Tensor]:
        if self.split == "train":
            return {"rays_o": self.all_rays_o[idx], "rays_d": self.all_rays_d[idx], "viewdirs": self.all_rays_d[idx],
                    "target": self.all_rgbs[idx]}
        return self._image_sample(self.img_files[idx])

    def ray_batches(self, batch_size: int = 2048, seed: int = 0) -> Iterator[Dict[str, Tensor]]:
        """Endless stream of random training ray batches (model.py:421-428: shuffle=True, batch 2048), drawn on the GPU."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        n = self.all_rays_o.shape[0]
        while True:
            idx = torch.randint(0, n, (batch_size,), generator=g, device=self.device)
            yield {"rays_o": self.all_rays_o[idx], "rays_d": self.all_rays_d[idx], "viewdirs": self.all_rays_d[idx],
                   "target": self.all_rgbs[idx]}


# --------------------------------------------------------------------------------------------------------------
# synthetic scene writer (analytic ray tracing on the host)
# --------------------------------------------------------------------------------------------------------------
_SPHERES = [((-0.35, 0.0, 0.05), 0.45, (0.85, 0.25, 0.2)), ((0.45, 0.1, 0.15), 0.35, (0.2, 0.35, 0.85))]
_SLAB = ((0.0, 0.0, -0.5), (0.8, 0.55, 0.08), (0.25, 0.7, 0.3))      # centre, half extents, colour
_LIGHT = np.array([0.5, 0.6, 0.8]) / np.linalg.norm([0.5, 0.6, 0.8])


def _trace(o: np.ndarray, d: np.ndarray) -> np.ndarray:
    """o [3], d [N,3] unit -> RGBA [N,4] in [0,1] (Lambert + ambient, alpha = hit)."""
    n = d.shape[0]
    t_best = np.full(n, np.inf)
    col = np.zeros((n, 3))
    nrm = np.zeros((n, 3))
    for c, r, rgb in _SPHERES:
        oc = o - np.asarray(c)
        b = d @ oc
        disc = b * b - (oc @ oc - r * r)
        t = -b - np.sqrt(np.maximum(disc, 0))
        hit = (disc > 0) & (t > 0) & (t < t_best)
        t_best = np.where(hit, t, t_best)
        p = o + t[:, None] * d
        nrm[hit] = ((p - np.asarray(c)) / r)[hit]
        col[hit] = rgb
    c, h, rgb = (np.asarray(x) for x in _SLAB)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (c - h - o) / d
        t1 = (c + h - o) / d
    tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
    tnear, tfar = tn.max(1), tf.min(1)
    hit = (tnear < tfar) & (tnear > 0) & (tnear < t_best)
    axis = tn.argmax(1)
    t_best = np.where(hit, tnear, t_best)
    bn = np.zeros((n, 3))
    bn[np.arange(n), axis] = -np.sign(d[np.arange(n), axis])
    nrm[hit] = bn[hit]
    col[hit] = rgb
    alpha = np.isfinite(t_best)
    shade = 0.3 + 0.7 * np.clip(nrm @ _LIGHT, 0, 1)
    return np.concatenate([col * shade[:, None] * alpha[:, None], alpha[:, None].astype(np.float64)], 1)


def write_synthetic_scene(root: str, img_wh: Tuple[int, int] = (64, 48), n_train: int = 40, n_val: int = 2, n_test: int = 4,
                          seed: int = 0) -> str:
    """Writes {root}/{train,val,test}/{rgb/r_i.png, transforms.json} (datagen/data_utils.py:189-242 layout)."""
    from PIL import Image
    w, h = img_wh
    focal = sapien_focal(h)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    dirs = np.stack([(xs - w / 2) / focal, -(ys - h / 2) / focal, -np.ones_like(xs)], -1).reshape(-1, 3)   # ray_utils.py:86-88
    k = seed * 100003
    for split, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        os.makedirs(os.path.join(root, split, "rgb"), exist_ok=True)
        frames = {}
        for i in range(n):
            c2w = sapien_camera(seed=k, radius=4.0).numpy().astype(np.float64)
            k += 1
            d = dirs @ c2w[:, :3].T
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            rgba = _trace(c2w[:, 3], d)
            Image.fromarray((np.clip(rgba, 0, 1) * 255 + 0.5).astype(np.uint8).reshape(h, w, 4), "RGBA").save(
                os.path.join(root, split, "rgb", "r_%d.png" % i))
            frames["r_%d" % i] = np.vstack([c2w, [0, 0, 0, 1]]).tolist()
        with open(os.path.join(root, split, "transforms.json"), "w") as f:
            json.dump({"focal": focal, "frames": frames}, f)
    return root
