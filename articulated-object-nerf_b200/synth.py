"""Deterministic synthetic inputs for benches / smoke runs: SAPIEN-shaped cameras (distribution of
datagen/data_utils.py:66-80, fovy 35 deg of datagen/data_gen.py:64) and seeded random parameters
with the reference's state_dict names and shapes.  Pure input generation -- no rendering arithmetic
lives here; rays come from the ray-generation kernel (``lib.raygen``).

tests/test_oracle.py checks that these generators produce the same tensors as the oracle's own
copies, so parity tests and bench runs see identical scenes.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import lib as L


def layer_names(kind: str):
    """state_dict layer names in aon_layer_shape order."""
    if kind == "vanilla":
        return (["pts_linears.%d" % i for i in range(8)]
                + ["views_linear.0", "bottleneck_layer", "density_layer", "rgb_layer"])
    return (["deformations_linear.%d" % i for i in range(4)] + ["deformation_layer"]
            + ["pts_linears.%d" % i for i in range(8)] + ["views_linear.%d" % i for i in range(4)]
            + ["bottleneck_layer", "density_layer", "rgb_layer"])


def make_state_dict(kind: str = "vanilla", seed: int = 0, sharp: bool = False) -> Dict[str, torch.Tensor]:
    """Uniform(+-sqrt(6/(in+out))) weights (xavier-like), small uniform biases; ``sharp`` scales the
    density head x60 with bias -3 to mimic a trained, peaky field."""
    g = torch.Generator().manual_seed(seed)
    kid = L.KIND_VANILLA if kind == "vanilla" else L.KIND_AUTODECODER
    shapes = L.layer_shapes(kid)
    sd: Dict[str, torch.Tensor] = {}
    for mlp in ("coarse_mlp.", "fine_mlp."):
        for name, (o, i) in zip(layer_names(kind), shapes):
            bound = math.sqrt(6.0 / (i + o))
            w = (torch.rand(o, i, generator=g) * 2 - 1) * bound
            b = (torch.rand(o, generator=g) * 2 - 1) / math.sqrt(i)
            if sharp and name == "density_layer":
                w = w * 60.0
                b = b - 3.0
            sd[mlp + name + ".weight"] = w
            sd[mlp + name + ".bias"] = b
    if kind != "vanilla":
        pre = "code_library.embedding_instance_"
        for nm, rows, cols in (("shape", 1, 128), ("appearance", 1, 128), ("articulation", 10, 32)):
            bound = math.sqrt(6.0 / (rows + cols))
            sd[pre + nm + ".weight"] = (torch.rand(rows, cols, generator=g) * 2 - 1) * bound
    return sd


def sapien_camera(seed: int = 0, radius: Optional[float] = None) -> torch.Tensor:
    """OpenGL c2w [3,4]: camera on a sphere r~U(3.5,4.5) looking at the origin."""
    rs = np.random.RandomState(seed)
    r = rs.uniform(3.5, 4.5) if radius is None else radius
    theta = rs.uniform(0, 2 * np.pi)
    phi = rs.uniform(0.15 * np.pi, 0.85 * np.pi)
    pos = np.array([r * np.sin(phi) * np.cos(theta), r * np.sin(phi) * np.sin(theta), r * np.cos(phi)])
    back = pos / np.linalg.norm(pos)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    return torch.tensor(np.stack([right, upv, back, pos], 1), dtype=torch.float32)


def sapien_focal(H: int) -> float:
    return 0.5 * H / math.tan(math.radians(17.5))
