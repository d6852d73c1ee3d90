"""GPU (B200): the CUDA path, called through the C ABI, against the CPU oracle and the committed
golden vectors of the reference.  Tolerances (written here, per stage):

* A1/A2 raygen, A3 sampling tables: <= 2e-6 abs (same fp32 formulae; matmul summation order differs)
* A7 sample_pdf: compared in CDF space -- |F(t_ours) - F(t_ref)| <= 2e-6 where F is the reference's own
  piecewise-linear cdf (fp64 interpolation of the oracle's fp32 cdf).  Comparing t directly is ill-posed:
  the kernel's warp-tree weight sum differs from ATen's vectorised sum by 1 ulp, and a 1-ulp cdf change
  moves a sample by ulp/(pdf density), i.e. by up to ~1e-4 in t inside bins that carry almost no mass
  (measured 8.5e-5) -- such samples carry no weight in the render.  t itself must still agree to 1e-3.
* one level given the REFERENCE's t values (stage-wise): rgb/acc/depth <= 1e-4 relative (north_star),
  weights <= 2e-5 abs
* full coarse+fine loop (end to end): <= max(1e-4, 5 x fp32 noise floor), where the noise floor is the
  distance of the fp32 reference itself from an fp64 evaluation of the same network on the same rays
  (oracle run in float64).  The reference's fine level is chaotic at the 1e-3 level on some scenes
  (importance samples re-order under 1-ulp perturbations of the coarse weights), so a
  bar below its own rounding noise would not be testing the kernel.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_cpu as O

pytestmark = pytest.mark.gpu

REL = 1e-4


def _t(x):
    return torch.from_numpy(np.asarray(x))


def relerr(a, b, floor=1e-2):
    """max |a - b| / max(|b|, floor).  DECLARED: the denominator is clamped at 1e-2, i.e. values below 0.01 (dark pixels, empty
    rays' accumulation) are held to an ABSOLUTE bar of 1e-6 = 1e-4 x 1e-2 instead of a relative one -- BASELINE.md 4.4 writes a
    1e-6 floor, which would turn fp32 rounding noise of near-zero outputs into O(1) "relative" errors.  bench.py's `parity`
    block reports the e2e error with both floors."""
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item()


def cdf_space_err(t_ours, t_ref, t_coarse, weights):
    """max |F(t_ours) - F(t_ref)| with F the oracle's piecewise-linear cdf over the 64 bin edges."""
    t_c = torch.broadcast_to(t_coarse, weights.shape)
    bins = (0.5 * (t_c[..., 1:] + t_c[..., :-1])).double()
    cdf = O.pdf_to_cdf(weights[..., 1:-1]).double()

    def F(t):
        t = t.double().contiguous()
        idx = torch.searchsorted(bins.contiguous(), t, right=True).clamp(1, bins.shape[-1] - 1)
        b0, b1 = torch.gather(bins, -1, idx - 1), torch.gather(bins, -1, idx)
        c0, c1 = torch.gather(cdf, -1, idx - 1), torch.gather(cdf, -1, idx)
        return c0 + ((t - b0) / (b1 - b0)).clamp(0, 1) * (c1 - c0)

    return (F(t_ours) - F(t_ref)).abs().max().item()


_TRUTH = {}


def noise_floor(name, sd, rays, lat, wb, ref32):
    """rel. distance of the fp32 reference outputs from an fp64 evaluation (per level / output)."""
    if name not in _TRUTH:
        sd64 = {k: v.double() for k, v in sd.items()}
        r64 = {k: v.double() for k, v in rays.items()}
        l64 = None if lat is None else {k: v.double() for k, v in lat.items()}
        with torch.no_grad():
            _TRUTH[name] = O.nerf_forward(sd64, r64, False, wb, 2.0, 6.0, latents=l64)
    t = _TRUTH[name]
    return [[relerr(ref32[lv][j], t[lv][j]) for j in range(3)] for lv in range(2)]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def aon(built_lib):
    from aon_b200 import lib, nerf
    return lib, nerf


def test_raygen(aon, dev, golden_dir):
    lib, nerf = aon
    g = np.load(os.path.join(golden_dir, "raygen_24x32.npz"))
    o, v, d = nerf.get_rays_from_pose(int(g["H"]), int(g["W"]), float(g["focal"]), g["c2w"], dev)
    assert (o.cpu() - _t(g["rays_o"])).abs().max() == 0
    assert (d.cpu() - _t(g["rays_d"])).abs().max() < 2e-6
    assert v.data_ptr() == d.data_ptr()
    # full-size image against the oracle
    H, W = 480, 640
    focal = 0.5 * H / np.tan(np.radians(17.5))
    c2w = O.sapien_camera(11)
    o, _, d = nerf.get_rays_from_pose(H, W, focal, c2w, dev)
    oo, ov, od = O.get_rays(O.get_ray_directions(H, W, focal), c2w)
    assert (d.cpu() - od).abs().max() < 2e-6 and torch.equal(o.cpu(), oo)


def test_sample_along_rays(aon, dev):
    lib, _ = aon
    t = lib.sample_along_rays(2.0, 6.0, 65, 0, dev)
    assert torch.equal(t.cpu(), O.coarse_t_table(64, 2.0, 6.0))
    g = torch.Generator().manual_seed(3)
    tr = torch.rand(77, 65, generator=g)
    t = lib.sample_along_rays(2.0, 6.0, 65, 77, dev, t_rand=tr.to(dev))
    z = torch.zeros(77, 3)
    want, _ = O.sample_along_rays(z, z, 64, 2.0, 6.0, True, t_rand=tr)
    assert torch.equal(t.cpu(), want)


def test_sample_pdf_golden(aon, dev, golden_dir):
    lib, _ = aon
    g = np.load(os.path.join(golden_dir, "sample_pdf.npz"))
    t_c, w = _t(g["t_coarse"]).to(dev), _t(g["weights"]).to(dev)
    tf = lib.sample_pdf(t_c, w, 128).cpu()
    want = _t(g["t_fine"])
    assert (tf[:, 1:] >= tf[:, :-1]).all()
    assert cdf_space_err(tf, want, t_c.cpu(), w.cpu()) < 2e-6
    assert (tf - want).abs().max() < 1e-3
    assert ((tf - want).abs() < 2e-6).float().mean() > 0.98
    # shared-table form: must be bit-identical to the per-ray form on rays that use the table
    tab = O.coarse_t_table(64, 2.0, 6.0)
    same = (t_c.cpu() == tab).all(-1).nonzero().flatten()
    if len(same):
        tf2 = lib.sample_pdf(tab.to(dev), w[same.to(dev)].contiguous(), 128).cpu()
        assert torch.equal(tf2, tf[same])


def test_sample_pdf_random_u(aon, dev):
    lib, _ = aon
    g = torch.Generator().manual_seed(5)
    R = 300
    t_c = O.coarse_t_table(64, 2.0, 6.0).expand(R, 65).contiguous()
    w = torch.rand(R, 65, generator=g) ** 6
    u = torch.rand(R, 128, generator=g)
    bins = 0.5 * (t_c[..., 1:] + t_c[..., :-1])
    z = torch.zeros(R, 3)
    want, _ = O.sample_pdf(bins, w[..., 1:-1], z, z, t_c, 128, True, u=u)
    got = lib.sample_pdf(t_c.to(dev), w.to(dev), 128, u=u.to(dev)).cpu()
    assert (got[:, 1:] >= got[:, :-1]).all()
    assert cdf_space_err(got, want, t_c, w) < 2e-6
    assert (got - want).abs().max() < 1e-3


def _load_case(path):
    g = np.load(path)
    kind, sharp = str(g["kind"]), bool(g["sharp"])
    sd = O.make_state_dict(kind, seed=0, sharp=sharp)
    rays = {k: _t(g[k]) for k in ("rays_o", "rays_d", "viewdirs")}
    lat = None
    if kind != "vanilla":
        lat = {k: _t(g["lat_" + k]) for k in ("density", "color", "articulation")}
    return g, kind, sd, rays, lat


def _make_net(nerf, kind, sd, dev):
    if kind == "vanilla":
        net = nerf.NeRF()
        net.load_state_dict(sd)
    else:
        net = nerf.NeRF_AE_Art()
        net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
    return net.to(dev).eval()


def _cases():
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(root, "*_R*_wb*.npz")))


@pytest.mark.parametrize("name", _cases())
def test_level_loop_fp32_vs_golden(aon, dev, golden_dir, name):
    lib, nerf = aon
    g, kind, sd, rays, lat = _load_case(os.path.join(golden_dir, name))
    net = _make_net(nerf, kind, sd, dev)
    net.precision = lib.PREC_FP32
    rd = {k: v.to(dev) for k, v in rays.items()}
    wb = bool(g["white_bkgd"])
    with torch.no_grad():
        out = net(rd, False, wb, 2.0, 6.0) if lat is None else net(rd, False, wb, 2.0, 6.0, {k: v.to(dev) for k, v in lat.items()})
    ref32 = [[_t(g["%s%d" % (nm, lv)]) for nm in ("rgb", "acc", "depth")] for lv in range(2)]
    floor = noise_floor(name, sd, rays, lat, wb, ref32)
    for lv in range(2):
        for j, nm in enumerate(("rgb", "acc", "depth")):
            e = relerr(out[lv][j].cpu(), ref32[lv][j])
            tol = max(REL, 5 * floor[lv][j])
            assert e < tol, "%s level %d %s rel err %g (tol %g, fp32 noise floor %g)" % (name, lv, nm, e, tol, floor[lv][j])


@pytest.mark.parametrize("name", ["vanilla_sharp_R33_wb1.npz", "autodecoder_sharp_R33_wb1_art7.npz",
                                  "vanilla_smooth_R33_wb0.npz"])
def test_stagewise_fp32(aon, dev, golden_dir, name):
    """Stage-wise: our level kernel fed the REFERENCE's t values; weights and outputs compared."""
    lib, nerf = aon
    g, kind, sd, rays, lat = _load_case(os.path.join(golden_dir, name))
    net = _make_net(nerf, kind, sd, dev)
    k = net.coarse_mlp.KIND
    o, d, v = (rays[x].to(dev) for x in ("rays_o", "rays_d", "viewdirs"))
    for lv, mlp in enumerate((net.coarse_mlp, net.fine_mlp)):
        lins = mlp.linears()
        packed = lib.pack_weights(k, lib.PREC_FP32, [l.weight for l in lins], [l.bias for l in lins])
        folded = None
        if lat is not None:
            folded = lib.fold_latents(k, lib.PREC_FP32, packed, lat["density"].to(dev), lat["color"].to(dev),
                                      lat["articulation"].to(dev))
        t = _t(g["t%d" % lv]).to(dev).contiguous()
        rgb, acc, depth, w = lib.render_level(k, lib.PREC_FP32, packed, folded, o, d, v, t, bool(g["white_bkgd"]))
        assert (w.cpu() - _t(g["weights%d" % lv])).abs().max() < 2e-5
        for a, nm in ((rgb, "rgb"), (acc, "acc"), (depth, "depth")):
            e = relerr(a.cpu(), _t(g["%s%d" % (nm, lv)]))
            assert e < REL, "%s level %d %s rel err %g" % (name, lv, nm, e)


def test_ragged_and_properties(aon, dev):
    """R not a multiple of 128, R=1; white-background range; ray-order invariance (sharding property)."""
    lib, nerf = aon
    sd = O.make_state_dict("vanilla", 0, sharp=True)
    net = _make_net(nerf, "vanilla", sd, dev)
    rays = O.sapien_rays(20, 27, seed=2)           # 540 rays
    rd = {k: v.to(dev) for k, v in rays.items()}
    with torch.no_grad():
        full = net(rd, False, True, 2.0, 6.0)
        a = net({k: v[:129].contiguous() for k, v in rd.items()}, False, True, 2.0, 6.0)
        b = net({k: v[129:].contiguous() for k, v in rd.items()}, False, True, 2.0, 6.0)
    for lv in range(2):
        for j in range(3):
            cat = torch.cat([a[lv][j], b[lv][j]], 0)
            assert torch.equal(cat, full[lv][j]), "ray-sharded result must equal the single-call result bit-for-bit"
    rgb, acc, depth = full[1]
    assert (acc <= 1 + 1e-5).all() and (acc >= 0).all()
    assert (rgb >= -1e-5).all() and (rgb <= 1 + 1e-5).all()
    assert (depth >= 0).all() and (depth <= 6.0 + 1e-4).all()


def test_render_image_host_matches_device_path(aon, dev):
    lib, nerf = aon
    sd = O.make_state_dict("vanilla", 0, sharp=True)
    net = _make_net(nerf, "vanilla", sd, dev)
    rays = O.sapien_rays(12, 16, seed=4)
    rd = {k: v.to(dev) for k, v in rays.items()}
    with torch.no_grad():
        want = net(rd, False, True, 2.0, 6.0)[1]
    pc = net._cache["coarse"].get(net.coarse_mlp, net.precision)
    pf = net._cache["fine"].get(net.fine_mlp, net.precision)
    out = lib.render_image_host(0, net.precision, pc, pf, None, None, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                2.0, 6.0, True)
    assert torch.equal(out[:, :3], want[0].cpu()) and torch.equal(out[:, 3], want[1].cpu()) and torch.equal(out[:, 4], want[2].cpu())


def test_missing_cuda_inputs_fail_loudly(aon):
    lib, nerf = aon
    net = nerf.NeRF()
    rays = O.sapien_rays(4, 4, 0)
    with pytest.raises(lib.AonError):
        with torch.no_grad():
            net(rays, False, True, 2.0, 6.0)
