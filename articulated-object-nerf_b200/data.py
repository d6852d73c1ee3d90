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


# --------------------------------------------------------------------------------------------------------------
# articulated multi-instance format (datasets/sapien_multi.py) + synthetic two-part "scissor" writer
# --------------------------------------------------------------------------------------------------------------
IDX_TO_DEG = {0: 0, 1: 10, 2: 20, 3: 30, 4: 40, 5: 50, 6: 60, 7: 70, 8: 80, 9: 90}     # sapien_multi.py:11-14 ("train")


def create_spheric_poses(radius: float = 4.0) -> Tensor:
    """sapien_multi.py:29-72: 40 poses on a circle, elevation -30 deg, [40,4,4]."""
    def pose(theta, phi, r):
        t = torch.eye(4); t[2, 3] = r
        ph, th = math.radians(phi), math.radians(theta)
        rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0], [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1]], dtype=torch.float32)
        rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]], dtype=torch.float32)
        fix = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
        return fix @ (rt @ (rp @ t))
    return torch.stack([pose(a, -30.0, radius) for a in np.linspace(-180, 180, 41)[:-1]], 0)


class SapienDatasetMulti:
    """datasets/sapien_multi.py:124-479: {root}/{instance}/train/{deg}_degree/{rgb,seg}/r_i.png + transforms.json with
    ``camera_angle_x``; train samples = 4096 random pixels of a random (instance, articulation state, image); val = one
    full image; test = 19 frames on the spherical path (articulation_id = frame index -> interpolated code table)."""

    def __init__(self, root_dir, split="train", img_wh=(320, 240), model_type=None, white_back=True, eval_inference=None,
                 device="cuda", seed=0):
        self.root_dir, self.split, self.img_wh, self.white_back = root_dir, split, tuple(img_wh), white_back
        self.device = torch.device(device)
        self.ids = sorted(f.name for f in os.scandir(root_dir) if f.is_dir())
        self.samples_per_epoch = 4000
        self.near, self.far = 2.0, 6.0
        w, h = self.img_wh
        self.image_sizes = np.array([[h, w] for _ in range(19 if eval_inference is not None else 1)])
        self.poses_test = create_spheric_poses(4.0)
        self.rng = np.random.RandomState(seed)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self._cache = {}

    def _states(self, inst):
        d = [f.name for f in os.scandir(os.path.join(self.root_dir, inst, "train")) if f.is_dir()]
        return sorted(d, key=lambda s: int(s.split("_")[0]))

    def _frame(self, inst, state, image_id, c2w=None):
        key = (inst, state, image_id, c2w is None)
        if key in self._cache:
            return self._cache[key]
        from PIL import Image
        base = os.path.join(self.root_dir, inst, "train", state)
        meta = json.load(open(os.path.join(base, "transforms.json")))
        files = sorted(os.listdir(os.path.join(base, "rgb")), key=lambda fn: int(fn.split("_")[1].split(".")[0]))
        fn = files[image_id % len(files)]
        w, h = self.img_wh
        focal = 0.5 * h / np.tan(0.5 * meta["camera_angle_x"]) * (w / 320)            # sapien_multi.py:281-283
        if c2w is None:
            c2w = torch.tensor(meta["frames"][fn.split(".")[0]], dtype=torch.float32)
        img = np.asarray(Image.open(os.path.join(base, "rgb", fn)).convert("RGB").resize((w, h), Image.LANCZOS), dtype=np.float32) / 255.0
        seg = np.asarray(Image.open(os.path.join(base, "seg", fn)).resize((w, h), Image.LANCZOS)) > 0
        bg = 1.0 if self.white_back else 0.0
        img = np.where(seg[..., None], img, bg).astype(np.float32)                       # get_masked_img_seg
        o, d = L.raygen(h, w, float(focal), c2w[:3, :4], self.device)
        out = (o, d, torch.from_numpy(img.reshape(-1, 3)).to(self.device), torch.from_numpy(seg.reshape(-1)).to(self.device))
        if len(self._cache) < 4096:
            self._cache[key] = out
        return out

    def __len__(self):
        return self.samples_per_epoch if self.split == "train" else (1 if self.split == "val" else 19)

    def __getitem__(self, idx) -> Dict[str, object]:
        w, h = self.img_wh
        inst_idx = int(self.rng.randint(0, len(self.ids)))
        inst = self.ids[inst_idx]
        states = self._states(inst)
        if self.split in ("train", "val"):
            deg_idx = int(self.rng.randint(0, len(states)))
            o, d, rgb, seg = self._frame(inst, states[deg_idx], int(self.rng.randint(0, 59)))
            if self.split == "train":
                pix = torch.randint(0, h * w, (4096,), generator=self.gen, device=self.device)
                o, d, rgb, seg = o[pix], d[pix], rgb[pix], seg[pix]
            s = {"deg": np.float32(np.deg2rad(IDX_TO_DEG[deg_idx])), "articulation_id": torch.tensor([deg_idx], device=self.device)}
        else:
            o, d, rgb, seg = self._frame(inst, states[0], idx, c2w=self.poses_test[idx])
            s = {"articulation_id": torch.tensor([idx], device=self.device)}
        s.update(rays_o=o, rays_d=d, viewdirs=d, target=rgb, instance_mask=seg, instance_id=torch.tensor([inst_idx], device=self.device),
                 img_wh=np.array((w, h)))
        return s

    def ray_batches(self) -> Iterator[Dict[str, object]]:
        while True:
            yield self[0]


def _rot_z(deg):
    a = math.radians(deg)
    return np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1.0]])


def _trace_scissor(o, d, deg):
    """Two slabs hinged at the origin: a fixed one along +x and one rotated by `deg` about z.  RGB [N,3], mask [N]."""
    n = d.shape[0]
    t_best = np.full(n, np.inf); col = np.zeros((n, 3)); nrm = np.zeros((n, 3))
    for rot, rgb in ((np.eye(3), (0.8, 0.3, 0.25)), (_rot_z(deg), (0.25, 0.4, 0.85))):
        ol, dl = rot.T @ o, d @ rot                          # ray in the part's frame
        c, h = np.array([0.55, 0.0, 0.0]), np.array([0.6, 0.12, 0.1])
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (c - h - ol) / dl, (c + h - ol) / dl
        tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
        tnear, tfar = tn.max(1), tf.min(1)
        hit = (tnear < tfar) & (tnear > 0) & (tnear < t_best)
        axis = tn.argmax(1)
        bn = np.zeros((n, 3)); bn[np.arange(n), axis] = -np.sign(dl[np.arange(n), axis])
        t_best = np.where(hit, tnear, t_best)
        nrm[hit] = (bn @ rot.T)[hit]; col[hit] = rgb
    mask = np.isfinite(t_best)
    return col * (0.3 + 0.7 * np.clip(nrm @ _LIGHT, 0, 1))[:, None], mask


def write_synthetic_articulated(root: str, img_wh=(64, 48), n_states: int = 4, n_images: int = 8, instances=("0001",), seed: int = 0) -> str:
    """{root}/{inst}/train/{deg}_degree/{rgb,seg}/r_i.png + transforms.json {"camera_angle_x", "frames"}."""
    from PIL import Image
    w, h = img_wh
    focal = sapien_focal(h)
    cax = 2 * math.atan(0.5 * h * (w / 320) / focal)          # inverse of the reader's focal rule (sapien_multi.py:281-283)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    dirs = np.stack([(xs - w / 2) / focal, -(ys - h / 2) / focal, -np.ones_like(xs)], -1).reshape(-1, 3)
    k = seed * 7919
    for inst in instances:
        for si in range(n_states):
            deg = IDX_TO_DEG[si]
            base = os.path.join(root, inst, "train", "%d_degree" % deg)
            os.makedirs(os.path.join(base, "rgb"), exist_ok=True); os.makedirs(os.path.join(base, "seg"), exist_ok=True)
            frames = {}
            for i in range(n_images):
                c2w = sapien_camera(seed=k, radius=4.0).numpy().astype(np.float64); k += 1
                dw = dirs @ c2w[:, :3].T
                dw /= np.linalg.norm(dw, axis=1, keepdims=True)
                rgb, mask = _trace_scissor(c2w[:, 3], dw, deg)
                Image.fromarray((np.clip(rgb, 0, 1) * 255 + 0.5).astype(np.uint8).reshape(h, w, 3), "RGB").save(os.path.join(base, "rgb", "r_%d.png" % i))
                Image.fromarray((mask.reshape(h, w) * 255).astype(np.uint8), "L").save(os.path.join(base, "seg", "r_%d.png" % i))
                frames["r_%d" % i] = np.vstack([c2w, [0, 0, 0, 1]]).tolist()
            with open(os.path.join(base, "transforms.json"), "w") as f:
                json.dump({"camera_angle_x": cax, "frames": frames}, f)
    return root
