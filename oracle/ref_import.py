"""TEST / BASELINE INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference modules of the hot path from a directory
that holds the reference's ``models/``, ``datasets/``, ``utils/`` packages and ``opt.py``:

* ``/root/reference`` in the build container (``oracle/gen_golden*.py``: pinning the oracle, writing the goldens);
* ``oracle/_ref/`` on the GPU box -- a git-ignored copy made by ``oracle/make_ref.py`` that travels with the snapshot
  like a built ``.so`` -- for the CPU arm of ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg).

Six third-party modules the reference imports at module scope are absent from the image;
they are stubbed.  None touches the arithmetic of the path except ``kornia.create_meshgrid``, restated below from
kornia 0.6.1 for ``normalized_coordinates=False``.  Nothing in the product package imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL = os.path.join(HERE, "_ref")
BUILD_CONTAINER = "/root/reference"


def available_root():
    """The directory to import the reference from, or None: oracle/_ref first (works on the GPU box), else /root/reference."""
    for root in (LOCAL, BUILD_CONTAINER):
        if os.path.isfile(os.path.join(root, "models", "vanilla_nerf", "model.py")):
            return root
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference(root: str):
    class _LM(torch.nn.Module):                      # stands in for pl.LightningModule
        @property
        def hparams(self):
            if not hasattr(self, "_hp"):
                object.__setattr__(self, "_hp", {})
            return self._hp

    def create_meshgrid(H, W, normalized_coordinates=False):
        xs = torch.linspace(0, W - 1, W)
        ys = torch.linspace(0, H - 1, H)
        g = torch.stack(torch.meshgrid([xs, ys], indexing="ij")).transpose(1, 2)
        return g.unsqueeze(0).permute(0, 2, 3, 1)

    for name, attrs in (("pytorch_lightning", {"LightningModule": _LM}), ("piqa", {}), ("piqa.lpips", {"LPIPS": object}),
                        ("piqa.ssim", {"SSIM": object}), ("kornia", {"create_meshgrid": create_meshgrid}),
                        ("matplotlib", {}), ("matplotlib.pyplot", {}), ("imageio", {}), ("torch_optimizer", {})):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name, **attrs)
    # the reference's packages are called models / datasets / utils: make sure they resolve to `root`
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "datasets", "utils", "opt")]:
        del sys.modules[k]
    sys.path.insert(0, root)
    argv, sys.argv = sys.argv, sys.argv[:1]          # opt.py parses sys.argv at import time
    try:
        import models.vanilla_nerf.helper as helper
        import models.vanilla_nerf.model as M
        import models.vanilla_nerf.model_autodecoder as MA
        from models.code_library import CodeLibraryArticulated
        from datasets.ray_utils import get_ray_directions, get_rays
    finally:
        sys.argv = argv
    return SimpleNamespace(root=root, helper=helper, M=M, MA=MA, CodeLibraryArticulated=CodeLibraryArticulated,
                           get_ray_directions=get_ray_directions, get_rays=get_rays)
