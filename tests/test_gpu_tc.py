"""GPU (B200): the tcgen05 tensor-core render kernel (AON_PREC_TC_*) through the C ABI.

* unit dump: the pre-activation output of every 128-wide GEMM unit of the fused kernel (ray tile 0,
  sample 0) against a plain torch fp32 evaluation of the same layers -- localises any operand-layout or
  descriptor error to a layer.  Tolerance: f16x3 2e-5 abs (+1e-5 rel; 2e-4 abs downstream of the
  auto-decoder's warped-position encoding), single-pass modes 3e-2.
* stage-wise (kernel fed the REFERENCE's t values): f16x3 within 1e-4 relative of the reference
  (north_star bar); f16 is the fast mode and is only required to stay within 3e-2; bf16 (8-bit significand:
  it cannot represent the 2^9-frequency encoding inputs well) within 0.3 -- both are reported, not parity.
* end to end f16x3: within max(1e-4, 5 x fp32 noise floor of the reference), as in test_gpu_parity.py (the two MMA issuer
  threads interleave their accumulations in a timing-dependent order, so tensor-core results vary by ~1 ulp from run to
  run; on the chaotic sharp cases that moves the end-to-end error between ~0.8e-4 and ~1.1e-4).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_cpu as O
from tests.test_gpu_parity import _load_case, _make_net, _t, noise_floor, relerr

pytestmark = pytest.mark.gpu

MODES = ["f16x3", "f16", "bf16"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def aon(built_lib):
    from aon_b200 import lib, nerf
    return lib, nerf


def _preacts(net_mlp, kind, pts, view, lat):
    """torch fp32 pre-activation outputs of every GEMM layer, in the kernel's unit order."""
    lin = {n: m for n, m in net_mlp.named_modules() if isinstance(m, torch.nn.Linear)}
    enc = lambda x, L: torch.cat([x] + [torch.sin(torch.cat([x[..., None, :] * (2.0 ** torch.arange(L, device=x.device))[:, None],
                                                             x[..., None, :] * (2.0 ** torch.arange(L, device=x.device))[:, None] + 0.5 * np.pi], -2)).reshape(x.shape[0], -1)], -1)
    out = []
    venc = enc(view, 4)
    if kind == "vanilla":
        h = e = enc(pts, 10)
    else:
        shp, art, app = (lat[k].expand(pts.shape[0], -1) for k in ("density", "articulation", "color"))
        h = torch.cat([pts, shp, art], -1)
        for i in range(4):
            z = lin["deformations_linear.%d" % i](h); out.append(z); h = F.relu(z)
        warped = lin["deformation_layer"](h) + pts
        e = torch.cat([enc(warped, 10), shp], -1)
        h = e
    for i in range(8):
        z = lin["pts_linears.%d" % i](h); out.append(z); h = F.relu(z)
        if i == 4:
            h = torch.cat([h, e], -1)
    z = lin["bottleneck_layer"](h); out.append(z)
    h = torch.cat([z, venc] + ([app] if kind != "vanilla" else []), -1)
    nv = 1 if kind == "vanilla" else 4
    for i in range(nv):
        z = lin["views_linear.%d" % i](h); out.append(z); h = F.relu(z)
    return [z for z in out]          # one unit per GEMM layer; z is [128, N] with N = 128 or 256


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind", ["vanilla", "autodecoder"])
def test_tc_unit_dump(aon, dev, kind, mode):
    lib, nerf = aon
    prec = lib.PRECISIONS[mode]
    sd = O.make_state_dict(kind, 0, sharp=True)
    net = _make_net(nerf, kind, sd, dev)
    k = net.coarse_mlp.KIND
    info = lib.debug_program_info(k, prec)
    assert info["smem_bytes"] <= 232448
    rays = O.sapien_rays(10, 16, seed=3)      # 160 rays: one full tile + a ragged one
    o, d, v = (rays[x].to(dev) for x in ("rays_o", "rays_d", "viewdirs"))
    lat = None
    if kind != "vanilla":
        lat = {kk: vv.to(dev) for kk, vv in O.code_library(sd, torch.tensor([0]), torch.tensor([5]), is_test=True).items()}
    lins = net.coarse_mlp.linears()
    packed = lib.pack_weights(k, prec, [l.weight for l in lins], [l.bias for l in lins])
    folded = None if lat is None else lib.fold_latents(k, prec, packed, lat["density"], lat["color"], lat["articulation"])
    dbg = torch.zeros(info["n_units"], 128, 256, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.debug_set_buffers(dbg, err)
    try:
        t = lib.sample_along_rays(2.0, 6.0, 65, o.shape[0], dev)
        lib.render_level(k, prec, packed, folded, o, d, v, t, True)
        torch.cuda.synchronize()
    finally:
        lib.debug_set_buffers(None, None)
    assert err.item() == 0, "pipeline barrier timed out, code %d" % err.item()
    with torch.no_grad():
        pts = o[:128] + t[0] * d[:128]
        want = _preacts(net.coarse_mlp, kind, pts, v[:128], lat)
    assert len(want) == dbg.shape[0]
    atol, rtol = (2e-5, 1e-5) if mode == "f16x3" else ((3e-2, 3e-2) if mode == "f16" else (0.3, 0.1))
    for ui in range(len(want)):
        if kind != "vanilla" and ui == 4 and mode == "f16x3":
            # everything downstream of the warped position x' = x + deformation(x) sees the 2^9-frequency
            # encoding amplify x''s ~1e-7 rounding differences to ~5e-5 in the sin arguments
            atol = 2e-4
        got = dbg[ui][:, :want[ui].shape[1]]
        diff = (got - want[ui]).abs()
        bound = atol + rtol * want[ui].abs()
        assert (diff <= bound).all(), "unit %d: max abs err %g (max |ref| %g)" % (ui, diff.max().item(), want[ui].abs().max().item())


CASES = ["vanilla_sharp_R33_wb1.npz", "vanilla_smooth_R33_wb0.npz", "autodecoder_sharp_R33_wb1_art7.npz",
         "autodecoder_smooth_R33_wb0_art3.npz", "vanilla_sharp_R3840_wb1.npz"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", CASES)
def test_tc_stagewise(aon, dev, golden_dir, name, mode):
    lib, nerf = aon
    prec = lib.PRECISIONS[mode]
    g, kind, sd, rays, lat = _load_case(os.path.join(golden_dir, name))
    net = _make_net(nerf, kind, sd, dev)
    k = net.coarse_mlp.KIND
    o, d, v = (rays[x].to(dev) for x in ("rays_o", "rays_d", "viewdirs"))
    tol = 1e-4 if mode == "f16x3" else (3e-2 if mode == "f16" else None)
    n = _t(g["t0"]).shape[0]                 # the R=3840 goldens keep per-sample tensors for the first rows only
    o, d, v = o[:n].contiguous(), d[:n].contiguous(), v[:n].contiguous()
    for lv, mlp in enumerate((net.coarse_mlp, net.fine_mlp)):
        lins = mlp.linears()
        packed = lib.pack_weights(k, prec, [l.weight for l in lins], [l.bias for l in lins])
        folded = None
        if lat is not None:
            folded = lib.fold_latents(k, prec, packed, lat["density"].to(dev), lat["color"].to(dev), lat["articulation"].to(dev))
        t = _t(g["t%d" % lv]).to(dev).contiguous()
        rgb, acc, depth, w = lib.render_level(k, prec, packed, folded, o, d, v, t, bool(g["white_bkgd"]))
        torch.cuda.synchronize()
        if tol is None:
            # bf16 (8-bit significand) is a range-safe throughput mode, not a parity mode: the 2^9-frequency
            # encoding inputs lose most of their phase in bf16; require sane output only
            assert all(torch.isfinite(x).all() for x in (rgb, acc, depth, w))
            assert (acc >= -1e-3).all() and (acc <= 1 + 1e-3).all()
            continue
        wref = _t(g["weights%d" % lv])
        werr = (w.cpu() - wref).abs().max().item()
        assert werr < (5e-5 if mode == "f16x3" else 3e-2)   # per-sample weights in [0,1], absolute; the official 1e-4 relative bar is on rgb/acc/depth below, "%s level %d weights abs err %g" % (name, lv, werr)
        for a, nm in ((rgb, "rgb"), (acc, "acc"), (depth, "depth")):
            e = relerr(a.cpu(), _t(g["%s%d" % (nm, lv)])[:n])
            assert e < tol, "%s %s level %d %s rel err %g" % (name, mode, lv, nm, e)


@pytest.mark.parametrize("name", CASES)
def test_tc_level_loop_f16x3(aon, dev, golden_dir, name):
    lib, nerf = aon
    g, kind, sd, rays, lat = _load_case(os.path.join(golden_dir, name))
    net = _make_net(nerf, kind, sd, dev)
    net.precision = lib.PREC_TC_F16X3
    rd = {k: v.to(dev) for k, v in rays.items()}
    wb = bool(g["white_bkgd"])
    with torch.no_grad():
        out = net(rd, False, wb, 2.0, 6.0) if lat is None else net(rd, False, wb, 2.0, 6.0, {k: v.to(dev) for k, v in lat.items()})
    ref32 = [[_t(g["%s%d" % (nm, lv)]) for nm in ("rgb", "acc", "depth")] for lv in range(2)]
    floor = noise_floor(name, sd, rays, lat, wb, ref32)
    for lv in range(2):
        for j, nm in enumerate(("rgb", "acc", "depth")):
            e = relerr(out[lv][j].cpu(), ref32[lv][j])
            tol = max(1e-4, 5 * floor[lv][j])
            assert e < tol, "%s level %d %s rel err %g (tol %g, fp32 noise floor %g)" % (name, lv, nm, e, tol, floor[lv][j])


def test_tc_fast_modes_psnr(aon, dev):
    """fast modes: rendered image within 0.1 dB-class distance of the fp32 render (PSNR(fast, fp32) > 45 dB)."""
    lib, nerf = aon
    sd = O.make_state_dict("vanilla", 0, sharp=True)
    net = _make_net(nerf, "vanilla", sd, dev)
    rays = O.sapien_rays(24, 32, seed=2)
    rd = {k: v.to(dev) for k, v in rays.items()}
    with torch.no_grad():
        net.precision = lib.PREC_FP32
        ref = net(rd, False, True, 2.0, 6.0)[1][0]
        for mode in ("f16", "bf16", "f16x3"):
            net.precision = lib.PRECISIONS[mode]
            img = net(rd, False, True, 2.0, 6.0)[1][0]
            mse = ((img - ref) ** 2).mean().item()
            psnr = -10 * np.log10(max(mse, 1e-20))
            print("PSNR(%s vs fp32) = %.1f dB" % (mode, psnr))
            assert psnr > {"f16": 45, "bf16": 25, "f16x3": 90}[mode], (mode, psnr)
