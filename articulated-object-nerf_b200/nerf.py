"""Host-side mirror of the reference's render-core modules, backed by libaon_b200.so.

Same class names, constructor defaults, ``forward`` signatures, return structure and
``state_dict`` keys as the reference, so checkpoints and callers carry over unchanged:

* ``NeRFMLP`` / ``NeRF``            <- models/vanilla_nerf/model.py:39-199
* ``NeRFMLP_AE`` / ``NeRF_AE_Art``  <- models/vanilla_nerf/model_autodecoder.py:60-337
* ``CodeLibraryArticulated``        <- models/code_library.py:12-71

``forward`` under ``torch.no_grad()`` (validation / ``--run_eval``) is ONE fused CUDA kernel launch per call
(``aon_render_rays``: coarse level, hierarchical sampling and fine level; outputs are column views of two [R,5] tensors).  When gradients are required (``training_step``) the level is evaluated stage by stage so that it can be
differentiated: sampling, positional encoding and activations + compositing run in our kernels with
hand-written adjoints (csrc/train_ops.cu); the vanilla MLP's contractions -- forward, dgrad and wgrad -- run as tcgen05
GEMMs (csrc/gemm_tc.cu via train_tc.py; ``train_gemm = "torch"`` selects library GEMMs under autograd instead, which is
also what the auto-decoder MLP still uses).  There is no CPU path: every call needs CUDA tensors and the built library.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

from . import lib as L
from . import train_tc

Tensor = torch.Tensor


def default_precision() -> int:
    return L.PRECISIONS[os.environ.get("AON_PRECISION", "fp32")]


# ------------------------------------------------------------------------------------------------
# parameter containers (state_dict ABI of the reference)
# ------------------------------------------------------------------------------------------------


class NeRFMLP(nn.Module):
    """Parameters of models/vanilla_nerf/model.py:39-93 (same names, shapes and init)."""

    KIND = L.KIND_VANILLA

    def __init__(self, min_deg_point, max_deg_point, deg_view, netdepth: int = 8, netwidth: int = 256,
                 netdepth_condition: int = 1, netwidth_condition: int = 128, skip_layer: int = 4,
                 input_ch: int = 3, input_ch_view: int = 3, num_rgb_channels: int = 3,
                 num_density_channels: int = 1):
        super().__init__()
        if (min_deg_point, max_deg_point, deg_view, netdepth, netwidth, netdepth_condition,
                netwidth_condition, skip_layer, input_ch, input_ch_view, num_rgb_channels,
                num_density_channels) != (0, 10, 4, 8, 256, 1, 128, 4, 3, 3, 3, 1):
            raise L.AonError("libaon_b200 implements the reference's hard-coded NeRFMLP configuration only "
                             "(model.py:124-135,218: the CLI never changes it)")
        pos_size = ((max_deg_point - min_deg_point) * 2 + 1) * input_ch
        view_pos_size = (deg_view * 2 + 1) * input_ch_view
        pts = [nn.Linear(pos_size, netwidth)]
        for idx in range(netdepth - 1):
            pts.append(nn.Linear(netwidth + pos_size if (idx % skip_layer == 0 and idx > 0) else netwidth, netwidth))
        for m in pts:
            init.xavier_uniform_(m.weight)
        self.pts_linears = nn.ModuleList(pts)
        self.views_linear = nn.ModuleList([nn.Linear(netwidth + view_pos_size, netwidth_condition)])
        self.bottleneck_layer = nn.Linear(netwidth, netwidth)
        self.density_layer = nn.Linear(netwidth, num_density_channels)
        self.rgb_layer = nn.Linear(netwidth_condition, num_rgb_channels)
        for m in (self.bottleneck_layer, self.density_layer, self.rgb_layer):
            init.xavier_uniform_(m.weight)

    def linears(self) -> List[nn.Linear]:
        """state_dict order == aon_layer_shape order."""
        return list(self.pts_linears) + list(self.views_linear) + [self.bottleneck_layer, self.density_layer, self.rgb_layer]

    def forward(self, x: Tensor, condition: Tensor) -> Tuple[Tensor, Tensor]:
        """torch-autograd evaluation (training only; model.py:95-120).  x [R,S,63], condition [R,27]."""
        S = x.shape[1]
        h = x.reshape(-1, x.shape[-1])
        inputs = h
        for i, lin in enumerate(self.pts_linears):
            h = F.relu(lin(h))
            if i % 4 == 0 and i > 0:
                h = torch.cat([h, inputs], -1)
        raw_density = self.density_layer(h).reshape(-1, S, 1)
        bott = self.bottleneck_layer(h)
        cond = condition[:, None, :].expand(-1, S, -1).reshape(-1, condition.shape[-1])
        h = F.relu(self.views_linear[0](torch.cat([bott, cond], -1)))
        return self.rgb_layer(h).reshape(-1, S, 3), raw_density


class NeRFMLP_AE(nn.Module):
    """Parameters of models/vanilla_nerf/model_autodecoder.py:60-169 (deformation_mlp=True,
    enc_after=True, embed_deg=False -- the configuration NeRF_AE_Art always builds)."""

    KIND = L.KIND_AUTODECODER

    def __init__(self, min_deg_point=0, max_deg_point=10, deg_view=4):
        super().__init__()
        if (min_deg_point, max_deg_point, deg_view) != (0, 10, 4):
            raise L.AonError("libaon_b200 implements the reference's hard-coded auto-decoder configuration only")
        d = [nn.Linear(3 + 128 + 32, 128)] + [nn.Linear(128, 128) for _ in range(3)]
        for m in d:
            init.xavier_uniform_(m.weight)
        self.deformations_linear = nn.ModuleList(d)
        self.deformation_layer = nn.Linear(128, 3)
        init.xavier_uniform_(self.deformation_layer.weight)
        pos_size = 63 + 128
        pts = [nn.Linear(pos_size, 256)]
        for idx in range(7):
            pts.append(nn.Linear(256 + pos_size if (idx % 4 == 0 and idx > 0) else 256, 256))
        for m in pts:
            init.xavier_uniform_(m.weight)
        self.pts_linears = nn.ModuleList(pts)
        views = [nn.Linear(256 + 27 + 128, 128)] + [nn.Linear(128, 128) for _ in range(3)]
        for m in views[1:]:
            init.xavier_uniform_(m.weight)
        self.views_linear = nn.ModuleList(views)
        self.bottleneck_layer = nn.Linear(256, 256)
        self.density_layer = nn.Linear(256, 1)
        self.rgb_layer = nn.Linear(128, 3)
        for m in (self.bottleneck_layer, self.density_layer, self.rgb_layer):
            init.xavier_uniform_(m.weight)

    def linears(self) -> List[nn.Linear]:
        return (list(self.deformations_linear) + [self.deformation_layer] + list(self.pts_linears)
                + list(self.views_linear) + [self.bottleneck_layer, self.density_layer, self.rgb_layer])

    def forward(self, pos: Tensor, condition: Tensor, latents: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
        """torch-autograd evaluation (training only; model_autodecoder.py:171-239)."""
        R, S, _ = pos.shape
        x0 = pos.reshape(-1, 3)
        n = R * S
        shape = latents["density"].expand(n, -1)
        app = latents["color"].expand(n, -1)
        art = latents["articulation"].expand(n, -1)
        h = torch.cat([x0, shape, art], -1)
        for lin in self.deformations_linear:
            h = F.relu(lin(h))
        warped = self.deformation_layer(h) + x0
        h = torch.cat([pos_enc_cuda(warped, 0, 10), shape], -1)
        inputs = h
        for i, lin in enumerate(self.pts_linears):
            h = F.relu(lin(h))
            if i % 4 == 0 and i > 0:
                h = torch.cat([h, inputs], -1)
        raw_density = self.density_layer(h).reshape(-1, S, 1)
        bott = self.bottleneck_layer(h)
        cond = condition[:, None, :].expand(-1, S, -1).reshape(-1, condition.shape[-1])
        h = torch.cat([bott, cond, app], -1)
        for lin in self.views_linear:
            h = F.relu(lin(h))
        return self.rgb_layer(h).reshape(-1, S, 3), raw_density


class _PosEncFn(torch.autograd.Function):
    """helper.py:136-140 (min_deg 0) through aon_pos_enc, adjoint through aon_pos_enc_backward."""

    @staticmethod
    def forward(ctx, x, max_deg):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.max_deg = max_deg
        return L.pos_enc(x, max_deg)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return L.pos_enc_backward(x, g.contiguous(), ctx.max_deg), None


def pos_enc_cuda(x: Tensor, min_deg: int, max_deg: int) -> Tensor:
    if min_deg != 0:
        raise L.AonError("pos_enc: the reference only ever uses min_deg = 0")
    return _PosEncFn.apply(x, max_deg) if x.requires_grad else L.pos_enc(x.contiguous(), max_deg)


class _CompositeFn(torch.autograd.Function):
    """Activations (model.py:186-187 / model_autodecoder.py:321-323) + volumetric_rendering (helper.py:157-195) through
    aon_composite; adjoint w.r.t. the raw MLP outputs through aon_composite_backward.  t_vals and dirs carry no
    gradient (the reference detaches the samples, helper.py:249; rays are data)."""

    @staticmethod
    def forward(ctx, raw_rgb, raw_sigma, t_vals, dirs, white_bkgd, act_mode):
        raw_rgb, raw_sigma = raw_rgb.contiguous(), raw_sigma.contiguous()
        comp, acc, depth, w, trans = L.composite(raw_rgb, raw_sigma, t_vals, dirs, white_bkgd, act_mode)
        ctx.save_for_backward(raw_rgb, raw_sigma, t_vals, dirs, w, trans)
        ctx.cfg = (white_bkgd, act_mode)
        ctx.mark_non_differentiable(w)
        return comp, acc, depth, w

    @staticmethod
    def backward(ctx, g_comp, g_acc, g_depth, _g_w):
        raw_rgb, raw_sigma, t_vals, dirs, w, trans = ctx.saved_tensors
        c = lambda g: None if g is None else g.contiguous()
        g_rgb, g_sigma = L.composite_backward(raw_rgb, raw_sigma, t_vals, dirs, w, trans, c(g_comp), c(g_acc), c(g_depth),
                                              *ctx.cfg)
        return g_rgb, g_sigma, None, None, None, None


def composite_cuda(raw_rgb, raw_sigma, t_vals, dirs, white_bkgd, act_mode):
    """raw_rgb [R,S,3], raw_sigma [R,S,1] -> (comp_rgb, acc, weights, depth) like helper.volumetric_rendering."""
    comp, acc, depth, w = _CompositeFn.apply(raw_rgb, raw_sigma.reshape(raw_sigma.shape[0], -1), t_vals, dirs, bool(white_bkgd), act_mode)
    return comp, acc, w, depth


# ------------------------------------------------------------------------------------------------
# packed-weight cache
# ------------------------------------------------------------------------------------------------


class _PackCache:
    """Packs an MLP's nn.Linear parameters into the kernel layout and re-packs only when a parameter
    was modified in place (optimizer step / load_state_dict), detected through tensor versions."""

    def __init__(self):
        self.key = None
        self.packed = None

    def get(self, mlp: nn.Module, precision: int) -> Tensor:
        lins = mlp.linears()
        key = (precision,) + tuple((p.data_ptr(), p._version) for l in lins for p in (l.weight, l.bias))
        if key != self.key:
            self.packed = L.pack_weights(mlp.KIND, precision, [l.weight for l in lins], [l.bias for l in lins])
            self.key = key
        return self.packed


def _check_rays(rays: Dict[str, Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    out = []
    for k in ("rays_o", "rays_d", "viewdirs"):
        t = rays[k]
        if not t.is_cuda:
            raise L.AonError("rays['%s'] must be a CUDA tensor: the B200 render path has no CPU fallback" % k)
        out.append(t.detach().reshape(-1, 3).float().contiguous())
    return tuple(out)


class _LevelLoop(nn.Module):
    """Shared coarse->fine driver (model.py:147-199 / model_autodecoder.py:278-337)."""

    num_coarse_samples = 64
    num_fine_samples = 128

    def _init_cache(self):
        self._cache = {"coarse": _PackCache(), "fine": _PackCache()}
        self.precision = default_precision()
        # training_step contractions: "tc" = hand-written tcgen05 GEMMs on fp16 hi+lo operand planes (train_tc.py; fp32-grade
        # gradients), "tc16" = the same GEMMs on single fp16 planes (fast mode, ~1e-3 gradient noise), "torch" = library GEMMs
        self.train_gemm = os.environ.get("AON_TRAIN_GEMM", "tc")
        # training forward of a level: "fused" = the whole MLP chain in one launch of the fused render kernel, every layer
        # output written to HBM once (aon_forward_train); "layers" = one tcgen05 GEMM launch per nn.Linear
        self.train_fwd = os.environ.get("AON_TRAIN_FWD", "fused")

    def _render(self, rays, randomized, white_bkgd, near, far, latents=None, t_rand=None, u=None):
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        o, d, v = _check_rays(rays)
        R, dev = o.shape[0], o.device
        nc = self.num_coarse_samples + 1
        if R == 0:      # empty ray batch: nothing to launch (the C ABI rejects null pointers)
            e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            return [(e(0, 3), e(0), e(0)), (e(0, 3), e(0), e(0))]
        rng = None
        if randomized:
            # helper.py:126 / :227 draw with torch.rand; here the draws are generated INSIDE the sampling kernels (Philox,
            # sampling.cuh) unless the caller injects t_rand / u tensors (tests): no [R,65] / [R,128] uniforms in HBM
            if t_rand is None or u is None:
                rng = self.rng(dev)
            t0 = L.sample_along_rays(near, far, nc, R, dev, t_rand=None if t_rand is None else t_rand.contiguous(),
                                     rng=rng if t_rand is None else None)
        if need_grad:
            if not randomized:
                t0 = L.sample_along_rays(near, far, nc, R, dev)
            ret = self._render_autograd(o, d, v, t0, u, white_bkgd, latents, rng if u is None else None)
            if rng is not None:
                rng.advance()
            return ret
        if rng is not None:                 # randomized render without gradients (not a reference code path): the fused
            if u is None:                   # kernel takes the inverse-cdf draws as a tensor
                u = rng.uniform(1, R, self.num_fine_samples)
            rng.advance()
        kind = self.coarse_mlp.KIND
        pc = self._cache["coarse"].get(self.coarse_mlp, self.precision)
        pf = self._cache["fine"].get(self.fine_mlp, self.precision)
        fc = ff = None
        if latents is not None:
            args = (latents["density"].detach().float().contiguous(), latents["color"].detach().float().contiguous(),
                    latents["articulation"].detach().float().contiguous())
            fc = L.fold_latents(kind, self.precision, pc, *args)
            ff = L.fold_latents(kind, self.precision, pf, *args)
        # the whole level loop (coarse level, hierarchical sampling, fine level) is ONE fused kernel launch
        fine, coarse = L.render_rays(kind, self.precision, pc, pf, fc, ff, o, d, v, near, far, white_bkgd,
                                     t_coarse=t0 if randomized else None, u=None if u is None else u.contiguous())
        return [(coarse[:, :3], coarse[:, 3], coarse[:, 4]), (fine[:, :3], fine[:, 3], fine[:, 4])]

    def rng(self, device) -> "L.Rng":
        """Philox state of the randomized sampling steps: seed = torch.initial_seed() + rank at first use (or ``rng_seed``)."""
        r = getattr(self, "_rng", None)
        if r is None or r.offset_dev.device != torch.device(device):
            from . import dist as D
            seed = getattr(self, "rng_seed", None)
            r = self._rng = L.Rng((torch.initial_seed() + D.world()[0]) if seed is None else seed, device)
            r.offset_dev.fill_(int(getattr(self, "rng_offset0", 0)))     # a resumed run continues the draw sequence (one offset per step)
        return r

    def _render_autograd(self, o, d, v, t0, u, white_bkgd, latents, rng=None):
        R = o.shape[0]
        ret = []
        t_vals = t0 if t0.dim() == 2 else t0[None, :].expand(R, -1).contiguous()
        view_enc = pos_enc_cuda(v, 0, 4)
        weights = None
        for level, mlp in enumerate((self.coarse_mlp, self.fine_mlp)):
            if level == 1:
                t_vals = L.sample_pdf(t_vals, weights.detach().contiguous(), self.num_fine_samples,
                                      u=None if u is None else u.contiguous(), rng=rng)
            tc = self.train_gemm in ("tc", "tc16")           # tcgen05 GEMMs: fp16 hi+lo planes ("tc") or the hi plane only ("tc16")
            if tc and self.train_fwd == "fused":
                if latents is None:
                    raw_rgb, raw_sigma = train_tc.vanilla_fused(o, d, v, t_vals, view_enc, mlp, x3=self.train_gemm == "tc")
                else:
                    raw_rgb, raw_sigma = train_tc.autodecoder_fused(o, d, v, t_vals, view_enc, latents, mlp, x3=self.train_gemm == "tc")
                comp, acc, weights, depth = composite_cuda(raw_rgb, raw_sigma, t_vals, d, white_bkgd, 0 if latents is None else 1)
                ret.append((comp, acc, depth))
                continue
            samples = o[:, None, :] + t_vals[..., None] * d[:, None, :]
            if latents is None and tc:
                raw_rgb, raw_sigma = train_tc.vanilla_mlp(pos_enc_cuda(samples, 0, 10), view_enc, samples.shape[1], mlp,
                                                          x3=self.train_gemm == "tc")
            elif latents is None:
                raw_rgb, raw_sigma = mlp(pos_enc_cuda(samples, 0, 10), view_enc)
            elif tc:
                raw_rgb, raw_sigma = train_tc.autodecoder_mlp(samples.contiguous(), view_enc, latents, mlp, x3=self.train_gemm == "tc")
            else:
                raw_rgb, raw_sigma = mlp(samples, view_enc, latents)
            comp, acc, weights, depth = composite_cuda(raw_rgb, raw_sigma, t_vals, d, white_bkgd, 0 if latents is None else 1)
            ret.append((comp, acc, depth))
        return ret


class NeRF(_LevelLoop):
    """models/vanilla_nerf/model.py:123-199."""

    def __init__(self, num_levels: int = 2, min_deg_point: int = 0, max_deg_point: int = 10, deg_view: int = 4,
                 num_coarse_samples: int = 64, num_fine_samples: int = 128, use_viewdirs: bool = True,
                 noise_std: float = 0.0, lindisp: bool = False):
        super().__init__()
        if (num_levels, num_coarse_samples, num_fine_samples, use_viewdirs, lindisp) != (2, 64, 128, True, False):
            raise L.AonError("libaon_b200 implements the reference's hard-coded NeRF configuration only")
        if noise_std != 0.0:
            raise L.AonError("noise_std != 0 is never active in the reference (model.py:133) and is not implemented")
        self.coarse_mlp = NeRFMLP(min_deg_point, max_deg_point, deg_view)
        self.fine_mlp = NeRFMLP(min_deg_point, max_deg_point, deg_view)
        self._init_cache()

    def forward(self, rays, randomized, white_bkgd, near, far, t_rand=None, u=None):
        return self._render(rays, randomized, white_bkgd, near, far, None, t_rand, u)


class NeRF_AE_Art(_LevelLoop):
    """models/vanilla_nerf/model_autodecoder.py:242-337."""

    def __init__(self, num_levels: int = 2, min_deg_point: int = 0, max_deg_point: int = 10, deg_view: int = 4,
                 num_coarse_samples: int = 64, num_fine_samples: int = 128, use_viewdirs: bool = True,
                 noise_std: float = 0.0, lindisp: bool = False, rgb_padding: float = 0.001,
                 density_bias: float = -1.0, enc_after=True, embed_deg=False):
        super().__init__()
        if (num_levels, num_coarse_samples, num_fine_samples, use_viewdirs, lindisp, rgb_padding, density_bias,
                enc_after, embed_deg, noise_std) != (2, 64, 128, True, False, 0.001, -1.0, True, False, 0.0):
            raise L.AonError("libaon_b200 implements the reference's hard-coded NeRF_AE_Art configuration only")
        self.coarse_mlp = NeRFMLP_AE(min_deg_point, max_deg_point, deg_view)
        self.fine_mlp = NeRFMLP_AE(min_deg_point, max_deg_point, deg_view)
        self._init_cache()

    def forward(self, rays, randomized, white_bkgd, near, far, latents, train=True, t_rand=None, u=None):
        return self._render(rays, randomized, white_bkgd, near, far, latents, t_rand, u)


class CodeLibraryArticulated(nn.Module):
    """models/code_library.py:12-71 (host-side: three tiny embedding tables)."""

    def __init__(self, hparams):
        super().__init__()
        self.embedding_instance_shape = nn.Embedding(hparams.N_max_objs, hparams.N_obj_code_length)
        self.embedding_instance_appearance = nn.Embedding(hparams.N_max_objs, hparams.N_obj_code_length)
        self.embedding_instance_articulation = nn.Embedding(10, 32)
        for e in (self.embedding_instance_shape, self.embedding_instance_appearance, self.embedding_instance_articulation):
            init.xavier_uniform_(e.weight)

    def get_interpolated_articulations(self, max_interpolations=2, device=None):
        w = self.embedding_instance_articulation.weight
        out = torch.zeros(2 * w.shape[0] - 1, w.shape[1], dtype=w.dtype, device=w.device)
        out[0::2] = w
        out[1::2] = (w[:-1] + w[1:]) / 2
        return out

    def forward(self, batch, is_test=False):
        ret = {"density": self.embedding_instance_shape(batch["instance_id"]),
               "color": self.embedding_instance_appearance(batch["instance_id"])}
        if is_test:
            ret["articulation"] = self.get_interpolated_articulations()[batch["articulation_id"]]
        else:
            ret["articulation"] = self.embedding_instance_articulation(batch["articulation_id"])
        return ret


# ------------------------------------------------------------------------------------------------
# ray generation with the reference's function names (datasets/ray_utils.py:71-159)
# ------------------------------------------------------------------------------------------------


def get_rays_from_pose(H: int, W: int, focal: float, c2w, device="cuda") -> Tuple[Tensor, Tensor, Tensor]:
    """get_ray_directions + get_rays(output_view_dirs=True) in one kernel.
    Returns (rays_o, viewdirs, rays_d) like the reference's get_rays; viewdirs is rays_d (the
    reference returns the same normalised values for both, ray_utils.py:146-147)."""
    o, d = L.raygen(H, W, focal, c2w, device)
    return o, d, d


def img2mse(x, y):
    return torch.mean((x - y) ** 2)


def mse2psnr(x):
    return -10.0 * torch.log(x) / np.log(10)
