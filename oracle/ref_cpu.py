"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Plain torch-CPU (ATen) restatement of the volume-rendering path of
zubair-irshad/articulated-object-nerf, written from the behaviour of the
reference, one function per reference function, each citing the file:line it
follows.  It exists so that parity can be checked on a machine where
``/root/reference`` is absent (the GPU box).  It is pinned against the
reference itself: ``oracle/gen_golden.py`` imports the unmodified reference in
the build container, runs both on the same seeded inputs and asserts
bit-equality before writing ``tests/golden/*.npz``; ``tests/test_oracle.py``
re-checks this file against those committed vectors on every run.

The reference ships no tests / golden vectors of its own,
so "pinned" here means "pinned to outputs of the reference run in the build
container", not to reference-held fixtures.

All functions take/return ``torch.Tensor`` on CPU.  ``dtype`` follows the
inputs (fp32 for parity; fp64 can be passed to get a higher-precision "truth"
when judging which of two fp32 results is closer).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------
# A1 / A2  ray generation                       datasets/ray_utils.py:71-159
# --------------------------------------------------------------------------


def pixel_grid(H: int, W: int) -> Tensor:
    """kornia==0.6.1 ``create_meshgrid(H, W, normalized_coordinates=False)[0]``
    (third-party, un-vendored; call site datasets/ray_utils.py:83): ``[H,W,2]``
    with ``[...,0] = x = column in [0,W-1]`` and ``[...,1] = y = row``."""
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    gx = xs[None, :].expand(H, W)
    gy = ys[:, None].expand(H, W)
    return torch.stack([gx, gy], -1)


def get_ray_directions(H: int, W: int, focal: float) -> Tensor:
    """datasets/ray_utils.py:71-90 -- camera-frame directions, no +0.5 centring."""
    g = pixel_grid(H, W)
    i, j = g[..., 0], g[..., 1]
    return torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)


def get_rays(directions: Tensor, c2w: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """datasets/ray_utils.py:118-159 with output_view_dirs=True.

    Returns ``(rays_o, viewdirs, rays_d)``.  The reference normalises
    ``viewdirs`` in place on the same storage as ``rays_d`` (ray_utils.py:146-147),
    so the returned ``rays_d`` is unit-norm and equal to ``viewdirs``; the
    restatement reproduces that aliasing outcome.  ``radii`` (ray_utils.py:139-144)
    is never consumed on the hot path and is not produced."""
    rays_d = directions @ c2w[:, :3].T
    rays_o = c2w[:, 3].expand(rays_d.shape)
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    return rays_o.reshape(-1, 3), viewdirs.reshape(-1, 3), viewdirs.reshape(-1, 3).clone()


# --------------------------------------------------------------------------
# A3  coarse sampling                  models/vanilla_nerf/helper.py:25-26,106-133
# --------------------------------------------------------------------------


def cast_rays(t_vals: Tensor, origins: Tensor, directions: Tensor) -> Tensor:
    """helper.py:25-26."""
    return origins[..., None, :] + t_vals[..., None] * directions[..., None, :]


def coarse_t_table(num_samples: int, near: float, far: float, dtype=torch.float32) -> Tensor:
    """helper.py:116-120 (lindisp=False): the ``num_samples+1`` deterministic t values."""
    s = torch.linspace(0.0, 1.0, num_samples + 1, dtype=dtype)
    return near * (1.0 - s) + far * s


def sample_along_rays(rays_o: Tensor, rays_d: Tensor, num_samples: int, near: float, far: float,
                      randomized: bool, t_rand: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """helper.py:106-133.  ``t_rand`` ([R, num_samples+1] uniform) may be injected so that a
    randomized run is reproducible across implementations (reference draws it with torch.rand)."""
    R = rays_o.shape[0]
    t = coarse_t_table(num_samples, near, far, rays_o.dtype)
    if randomized:
        mids = 0.5 * (t[1:] + t[:-1])
        upper = torch.cat([mids, t[-1:]], -1)
        lower = torch.cat([t[:1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand((R, num_samples + 1), dtype=rays_o.dtype)
        t = lower + (upper - lower) * t_rand
    else:
        t = torch.broadcast_to(t, (R, num_samples + 1))
    return t, cast_rays(t, rays_o, rays_d)


# --------------------------------------------------------------------------
# A4  positional encoding                                  helper.py:136-140
# --------------------------------------------------------------------------


def pos_enc(x: Tensor, min_deg: int, max_deg: int) -> Tensor:
    """helper.py:136-140 -- ``[x, sin(2^k x) (k-major, xyz-minor), sin(2^k x + pi/2)]``.
    The cosine half is a *shifted sine*: the fp32 sum ``2^k x + fl32(pi/2)`` is formed first."""
    scales = torch.tensor([2 ** i for i in range(min_deg, max_deg)]).type_as(x)
    xb = (x[..., None, :] * scales[:, None]).reshape(list(x.shape[:-1]) + [-1])
    feat = torch.sin(torch.cat([xb, xb + 0.5 * np.pi], dim=-1))
    return torch.cat([x, feat], dim=-1)


# --------------------------------------------------------------------------
# A6  compositing                                          helper.py:157-195
# --------------------------------------------------------------------------


def volumetric_rendering(rgb: Tensor, density: Tensor, t_vals: Tensor, dirs: Tensor,
                         white_bkgd: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """helper.py:157-195 -> (comp_rgb, acc, weights, depth)."""
    eps = 1e-10
    last = torch.full_like(t_vals[..., :1], 1e10)
    dists = torch.cat([t_vals[..., 1:] - t_vals[..., :-1], last], -1)
    dists = dists * torch.norm(dirs[..., None, :], dim=-1)
    alpha = 1.0 - torch.exp(-density[..., 0] * dists)
    trans = torch.cat([torch.ones_like(alpha[..., :1]),
                       torch.cumprod(1.0 - alpha[..., :-1] + eps, dim=-1)], -1)
    weights = alpha * trans
    comp_rgb = (weights[..., None] * rgb).sum(dim=-2)
    depth = (weights * t_vals).sum(dim=-1)
    depth = torch.nan_to_num(depth, float("inf"))
    # helper.py:180 -- chunk-global clamp to [min, max]; numerically the identity.
    depth = torch.clamp(depth, torch.min(depth), torch.max(depth))
    acc = weights.sum(dim=-1)
    if white_bkgd:
        comp_rgb = comp_rgb + (1.0 - acc[..., None])
    return comp_rgb, acc, weights, depth


# --------------------------------------------------------------------------
# A7  hierarchical sampling                                helper.py:203-252
# --------------------------------------------------------------------------


def fine_u_table(num_samples: int, dtype=torch.float32, float_min_eps: float = 2 ** -32) -> Tensor:
    """helper.py:229 -- deterministic u; the last entry rounds to exactly 1.0 in fp32."""
    return torch.linspace(0.0, 1.0 - float_min_eps, num_samples, dtype=dtype)


def pdf_to_cdf(weights: Tensor) -> Tensor:
    """helper.py:206-222 -- padded pdf and clamped cdf with 0/1 end caps: [R, n_bins]."""
    eps = 1e-5
    wsum = weights.sum(dim=-1, keepdim=True)
    padding = torch.fmax(torch.zeros_like(wsum), eps - wsum)
    w = weights + padding / weights.shape[-1]
    wsum = wsum + padding
    pdf = w / wsum
    cdf = torch.fmin(torch.ones_like(pdf[..., :-1]), torch.cumsum(pdf[..., :-1], dim=-1))
    z = torch.zeros(list(cdf.shape[:-1]) + [1], dtype=cdf.dtype)
    return torch.cat([z, cdf, z + 1.0], -1)


def sorted_piecewise_constant_pdf(bins: Tensor, weights: Tensor, num_samples: int, randomized: bool,
                                  u: Optional[Tensor] = None) -> Tensor:
    """helper.py:203-243, mask-max/min bracketing exactly as the reference does it
    (materialises [R, n_bins, num_samples] temporaries).  ``u`` may be injected."""
    cdf = pdf_to_cdf(weights)
    if u is None:
        if randomized:
            u = torch.rand(list(cdf.shape[:-1]) + [num_samples], dtype=cdf.dtype)
        else:
            u = fine_u_table(num_samples, cdf.dtype)
    u = torch.broadcast_to(u, list(cdf.shape[:-1]) + [num_samples])
    mask = u[..., None, :] >= cdf[..., :, None]

    def lower(x):
        return (mask * x[..., None] + ~mask * x[..., :1, None]).max(dim=-2)[0]

    def upper(x):
        return (~mask * x[..., None] + mask * x[..., -1:, None]).min(dim=-2)[0]

    b0, b1, c0, c1 = lower(bins), upper(bins), lower(cdf), upper(cdf)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return b0 + t * (b1 - b0)


def sorted_piecewise_constant_pdf_bracket(bins: Tensor, weights: Tensor, num_samples: int,
                                          u: Optional[Tensor] = None) -> Tensor:
    """Same result as :func:`sorted_piecewise_constant_pdf` via an explicit bracket search
    (``idx = #(cdf <= u)``) -- the formulation the CUDA kernel uses.  Kept to document and test
    the equivalence (bit-identical on every case in tests/test_oracle.py)."""
    cdf = pdf_to_cdf(weights)
    if u is None:
        u = fine_u_table(num_samples, cdf.dtype)
    u = torch.broadcast_to(u, list(cdf.shape[:-1]) + [num_samples]).contiguous()
    n = cdf.shape[-1]
    idx = torch.searchsorted(cdf.contiguous(), u, right=True)
    i0 = torch.clamp(idx - 1, 0, n - 1)
    i1 = torch.clamp(idx, 0, n - 1)
    c0, c1 = torch.gather(cdf, -1, i0), torch.gather(cdf, -1, i1)
    bb = torch.broadcast_to(bins, cdf.shape)
    b0, b1 = torch.gather(bb, -1, i0), torch.gather(bb, -1, i1)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return b0 + t * (b1 - b0)


def sample_pdf(bins: Tensor, weights: Tensor, origins: Tensor, directions: Tensor, t_vals: Tensor,
               num_samples: int, randomized: bool, u: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """helper.py:246-252 -- draw, merge with the coarse t values by a full sort, cast."""
    t_new = sorted_piecewise_constant_pdf(bins, weights, num_samples, randomized, u=u).detach()
    t_all = torch.sort(torch.cat([t_vals, t_new], dim=-1), dim=-1).values
    return t_all, cast_rays(t_all, origins, directions)


# --------------------------------------------------------------------------
# A5  vanilla MLP                         models/vanilla_nerf/model.py:39-120
# --------------------------------------------------------------------------

VANILLA_LAYERS: List[Tuple[str, int, int]] = (
    [("pts_linears.0", 256, 63)]
    + [("pts_linears.%d" % i, 256, 256) for i in range(1, 5)]
    + [("pts_linears.5", 256, 319), ("pts_linears.6", 256, 256), ("pts_linears.7", 256, 256),
       ("views_linear.0", 128, 283), ("bottleneck_layer", 256, 256), ("density_layer", 1, 256),
       ("rgb_layer", 3, 128)]
)
"""(name, out, in) in state_dict order (model.py:65-93)."""

AUTODECODER_LAYERS: List[Tuple[str, int, int]] = (
    [("deformations_linear.0", 128, 163)]
    + [("deformations_linear.%d" % i, 128, 128) for i in range(1, 4)]
    + [("deformation_layer", 3, 128), ("pts_linears.0", 256, 191)]
    + [("pts_linears.%d" % i, 256, 256) for i in range(1, 5)]
    + [("pts_linears.5", 256, 447), ("pts_linears.6", 256, 256), ("pts_linears.7", 256, 256),
       ("views_linear.0", 128, 411)]
    + [("views_linear.%d" % i, 128, 128) for i in range(1, 4)]
    + [("bottleneck_layer", 256, 256), ("density_layer", 1, 256), ("rgb_layer", 3, 128)]
)
"""(name, out, in) in state_dict order (model_autodecoder.py:95-169)."""


def _lin(p: Dict[str, Tensor], prefix: str, name: str, x: Tensor) -> Tensor:
    return F.linear(x, p[prefix + name + ".weight"], p[prefix + name + ".bias"])


def mlp_vanilla(p: Dict[str, Tensor], prefix: str, x: Tensor, condition: Tensor) -> Tuple[Tensor, Tensor]:
    """model.py:95-120.  ``x`` [R,S,63], ``condition`` [R,27] -> raw_rgb [R,S,3], raw_density [R,S,1]."""
    S, feat = x.shape[1:]
    h = x.reshape(-1, feat)
    inputs = h
    for i in range(8):
        h = torch.relu(_lin(p, prefix, "pts_linears.%d" % i, h))
        if i % 4 == 0 and i > 0:
            h = torch.cat([h, inputs], -1)
    raw_density = _lin(p, prefix, "density_layer", h).reshape(-1, S, 1)
    bott = _lin(p, prefix, "bottleneck_layer", h)
    cond = torch.tile(condition[:, None, :], (1, S, 1)).reshape(-1, condition.shape[-1])
    h = torch.relu(_lin(p, prefix, "views_linear.0", torch.cat([bott, cond], -1)))
    raw_rgb = _lin(p, prefix, "rgb_layer", h).reshape(-1, S, 3)
    return raw_rgb, raw_density


# --------------------------------------------------------------------------
# A9  articulated auto-decoder MLP   models/vanilla_nerf/model_autodecoder.py:171-239
# --------------------------------------------------------------------------


def mlp_autodecoder(p: Dict[str, Tensor], prefix: str, pos: Tensor, condition: Tensor,
                    latents: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
    """model_autodecoder.py:171-239 (deformation_mlp=True, enc_after=True, embed_deg=False).
    ``pos`` [R,S,3] raw xyz; latents density/color [1,128], articulation [1,32]."""
    R, S, feat = pos.shape
    x0 = pos.reshape(-1, feat)
    n = R * S
    shape = latents["density"].expand(n, -1)
    app = latents["color"].expand(n, -1)
    art = latents["articulation"].expand(n, -1)
    h = torch.cat([x0, shape, art], -1)
    for i in range(4):
        h = torch.relu(_lin(p, prefix, "deformations_linear.%d" % i, h))
    warped = _lin(p, prefix, "deformation_layer", h) + x0
    h = torch.cat([pos_enc(warped, 0, 10), shape], -1)
    inputs = h
    for i in range(8):
        h = torch.relu(_lin(p, prefix, "pts_linears.%d" % i, h))
        if i % 4 == 0 and i > 0:
            h = torch.cat([h, inputs], -1)
    raw_density = _lin(p, prefix, "density_layer", h).reshape(-1, S, 1)
    bott = _lin(p, prefix, "bottleneck_layer", h)
    cond = torch.tile(condition[:, None, :], (1, S, 1)).reshape(-1, condition.shape[-1])
    h = torch.cat([bott, cond, app], -1)
    for i in range(4):
        h = torch.relu(_lin(p, prefix, "views_linear.%d" % i, h))
    raw_rgb = _lin(p, prefix, "rgb_layer", h).reshape(-1, S, 3)
    return raw_rgb, raw_density


# --------------------------------------------------------------------------
# A8  level loop                                           model.py:147-199
#                                             model_autodecoder.py:278-337
# --------------------------------------------------------------------------


def nerf_forward(p: Dict[str, Tensor], rays: Dict[str, Tensor], randomized: bool, white_bkgd: bool,
                 near: float, far: float, latents: Optional[Dict[str, Tensor]] = None,
                 t_rand: Optional[Tensor] = None, u: Optional[Tensor] = None,
                 stages: Optional[dict] = None, num_coarse: int = 64, num_fine: int = 128):
    """Both level loops.  ``latents is None`` -> vanilla (model.py:147-199: sigmoid rgb, relu sigma);
    otherwise the articulated auto-decoder (model_autodecoder.py:278-337: padded sigmoid,
    softplus(raw-1)).  Returns ``[(comp_rgb, acc, depth)] * 2``.  If ``stages`` is a dict the
    per-stage tensors are stored in it (used to build golden vectors)."""
    ret = []
    t_vals = weights = None
    for level in range(2):
        if level == 0:
            t_vals, samples = sample_along_rays(rays["rays_o"], rays["rays_d"], num_coarse, near, far,
                                                randomized, t_rand=t_rand)
            prefix = "coarse_mlp."
        else:
            t_mids = 0.5 * (t_vals[..., 1:] + t_vals[..., :-1])
            t_vals, samples = sample_pdf(t_mids, weights[..., 1:-1], rays["rays_o"], rays["rays_d"],
                                         t_vals, num_fine, randomized, u=u)
            prefix = "fine_mlp."
        view_enc = pos_enc(rays["viewdirs"], 0, 4)
        if latents is None:
            raw_rgb, raw_sigma = mlp_vanilla(p, prefix, pos_enc(samples, 0, 10), view_enc)
            rgb = torch.sigmoid(raw_rgb)
            sigma = torch.relu(raw_sigma)
        else:
            raw_rgb, raw_sigma = mlp_autodecoder(p, prefix, samples, view_enc, latents)
            rgb = torch.sigmoid(raw_rgb) * (1 + 2 * 0.001) - 0.001
            sigma = F.softplus(raw_sigma + (-1.0))
        comp_rgb, acc, weights, depth = volumetric_rendering(rgb, sigma, t_vals, rays["rays_d"], white_bkgd)
        if stages is not None:
            stages["t%d" % level] = t_vals
            stages["raw_rgb%d" % level] = raw_rgb
            stages["raw_sigma%d" % level] = raw_sigma
            stages["weights%d" % level] = weights
        ret.append((comp_rgb, acc, depth))
    return ret


def render_chunked(p, rays, white_bkgd, near, far, chunk=3840, latents=None):
    """model.py:295-348 / model_autodecoder.py:479-541 -- Python chunk loop, fine level kept."""
    R = rays["rays_o"].shape[0]
    out = {"comp_rgb": [], "acc": [], "depth": []}
    for i in range(0, R, chunk):
        sub = {k: v[i:i + chunk] for k, v in rays.items()}
        fine = nerf_forward(p, sub, False, white_bkgd, near, far, latents=latents)[1]
        for k, v in zip(("comp_rgb", "acc", "depth"), fine):
            out[k].append(v)
    return {k: torch.cat(v, 0) for k, v in out.items()}


# --------------------------------------------------------------------------
# A10  code library                                 models/code_library.py:36-71
# --------------------------------------------------------------------------


def interpolated_articulations(table: Tensor) -> Tensor:
    """code_library.py:55-71 -- 10 learnt rows at even indices, midpoints at odd: [19,32]."""
    n = table.shape[0]
    out = torch.zeros(2 * n - 1, table.shape[1], dtype=table.dtype)
    out[0::2] = table
    out[1::2] = (table[:-1] + table[1:]) / 2
    return out


def code_library(p: Dict[str, Tensor], instance_id: Tensor, articulation_id: Tensor, is_test: bool = False):
    """code_library.py:36-53 -- keys density / color / articulation."""
    pre = "code_library.embedding_instance_"
    art = p[pre + "articulation.weight"]
    if is_test:
        art = interpolated_articulations(art)
    return {"density": p[pre + "shape.weight"][instance_id],
            "color": p[pre + "appearance.weight"][instance_id],
            "articulation": art[articulation_id]}


# --------------------------------------------------------------------------
# A12  loss                                                 helper.py:17-22
# --------------------------------------------------------------------------


def img2mse(x: Tensor, y: Tensor) -> Tensor:
    return torch.mean((x - y) ** 2)


def mse2psnr(x: Tensor) -> Tensor:
    return -10.0 * torch.log(x) / np.log(10)


# --------------------------------------------------------------------------
# deterministic synthetic weights / cameras shared by goldens, tests, bench
# (not a restatement of anything: just reproducible inputs)
# --------------------------------------------------------------------------


def make_state_dict(kind: str = "vanilla", seed: int = 0, sharp: bool = False) -> Dict[str, Tensor]:
    """Deterministic synthetic parameters with the reference's state_dict names and shapes
    (SURVEY.md section 5 "weight ABI").  Uniform(+-sqrt(6/(in+out))) weights like the reference's
    xavier init, small uniform biases.  ``sharp=True`` scales the density head (x60, bias -3 for
    relu density) to mimic a trained, peaky field."""
    g = torch.Generator().manual_seed(seed)
    layers = VANILLA_LAYERS if kind == "vanilla" else AUTODECODER_LAYERS
    sd: Dict[str, Tensor] = {}
    for mlp in ("coarse_mlp.", "fine_mlp."):
        for name, o, i in layers:
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


def sapien_camera(seed: int = 0, radius: Optional[float] = None) -> Tensor:
    """A SAPIEN-shaped OpenGL c2w [3,4]: camera on a sphere r~U(3.5,4.5) looking at the origin
    (distribution of datagen/data_utils.py:66-80)."""
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
    c2w = np.stack([right, upv, back, pos], 1)
    return torch.tensor(c2w, dtype=torch.float32)


def sapien_rays(H: int, W: int, seed: int = 0) -> Dict[str, Tensor]:
    """Rays of one synthetic SAPIEN-shaped view through A1+A2 (fovy 35 deg, datagen/data_gen.py:64)."""
    focal = 0.5 * H / math.tan(math.radians(17.5))
    d = get_ray_directions(H, W, focal)
    o, v, dd = get_rays(d, sapien_camera(seed))
    return {"rays_o": o.contiguous(), "rays_d": dd.contiguous(), "viewdirs": v.contiguous()}
