"""GPU (B200): the fused image kernel (aon_render_rays / aon_render_image: coarse level, in-kernel hierarchical sampling,
fine level and -- for aon_render_image -- ray generation in ONE launch) through the C ABI.

* fused == three-launch path (aon_render_level, aon_sample_pdf, aon_render_level) BIT for bit, every tensor-core mode,
  both model kinds, ragged / single-ray / multi-wave batches (the multi-wave batch exercises the split into fused full
  waves + sample-segmented tail), deterministic and randomized draws;
* aon_render_image (camera in, pixels out) == raygen + aon_render_rays, and any split of the pixel range into blocks
  (what N ranks render) concatenates to the full image exactly -- SURVEY 8e "sharded == 1-GPU exactly";
* against the oracle / golden vectors: auto-decoder edge cases (R = 1 / 127 / 129, empty and saturated density), the
  auto-decoder sharp R = 3840 golden, the product CodeLibraryArticulated (19-row test-time interpolation) against the
  reference-generated latents in the goldens;
* A7 alone: t_fine of aon_sample_pdf against the golden t_fine within 1e-6 abs (the 63-term weight sum follows ATen's
  reduction order) and identical merge ranks.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_cpu as O
from tests.test_gpu_parity import _load_case, _make_net, _t, noise_floor, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def aon(built_lib):
    from aon_b200 import lib, nerf
    return lib, nerf


def _setup(lib, nerf, kind, prec, dev, art=6):
    sd = O.make_state_dict(kind, 0, sharp=True)
    net = _make_net(nerf, kind, sd, dev)
    net.precision = prec
    k = net.coarse_mlp.KIND
    pc = net._cache["coarse"].get(net.coarse_mlp, prec)
    pf = net._cache["fine"].get(net.fine_mlp, prec)
    fc = ff = lat = None
    if kind != "vanilla":
        lat = {n: v.to(dev) for n, v in O.code_library(sd, torch.tensor([0]), torch.tensor([art]), is_test=True).items()}
        a = (lat["density"].contiguous(), lat["color"].contiguous(), lat["articulation"].contiguous())
        fc, ff = lib.fold_latents(k, prec, pc, *a), lib.fold_latents(k, prec, pf, *a)
    return sd, net, k, pc, pf, fc, ff, lat


def _three_launch(lib, k, prec, pc, pf, fc, ff, o, d, t0, u=None):
    rgb0, acc0, dep0, w0 = lib.render_level(k, prec, pc, fc, o, d, d, t0, True, True)
    t1 = lib.sample_pdf(t0, w0, 128, u=u)
    rgb1, acc1, dep1, _ = lib.render_level(k, prec, pf, ff, o, d, d, t1, True, False)
    return (torch.cat([rgb0, acc0[:, None], dep0[:, None]], 1), torch.cat([rgb1, acc1[:, None], dep1[:, None]], 1))


@pytest.mark.parametrize("kind,prec_name", [("vanilla", "f16x3"), ("autodecoder", "f16x3"), ("vanilla", "f16"), ("autodecoder", "bf16")])
def test_fused_equals_three_launch(aon, dev, kind, prec_name):
    lib, nerf = aon
    prec = lib.PRECISIONS[prec_name]
    sd, net, k, pc, pf, fc, ff, lat = _setup(lib, nerf, kind, prec, dev)
    slots = torch.cuda.get_device_properties(dev).multi_processor_count // 2
    rays = O.sapien_rays(120, 200, seed=9)
    for R in (1, 129, 3840, slots * 256 + 300):
        o, d = rays["rays_o"][:R].contiguous().to(dev), rays["rays_d"][:R].contiguous().to(dev)
        t0 = lib.sample_along_rays(2.0, 6.0, 65, R, dev)
        c_ref, f_ref = _three_launch(lib, k, prec, pc, pf, fc, ff, o, d, t0)
        fine, coarse = lib.render_rays(k, prec, pc, pf, fc, ff, o, d, d, 2.0, 6.0, True)
        assert torch.equal(coarse, c_ref), (R, (coarse - c_ref).abs().max().item())
        assert torch.equal(fine, f_ref), (R, (fine - f_ref).abs().max().item())
        lib.debug_no_tail_split(True)          # everything in the fused launch, ragged last wave included
        try:
            fine2, coarse2 = lib.render_rays(k, prec, pc, pf, fc, ff, o, d, d, 2.0, 6.0, True)
        finally:
            lib.debug_no_tail_split(False)
        assert torch.equal(fine2, f_ref) and torch.equal(coarse2, c_ref), R
        lib.debug_no_fuse(True)
        try:
            fine3, coarse3 = lib.render_rays(k, prec, pc, pf, fc, ff, o, d, d, 2.0, 6.0, True)
        finally:
            lib.debug_no_fuse(False)
        assert torch.equal(fine3, f_ref) and torch.equal(coarse3, c_ref), R


def test_fused_randomized_draws(aon, dev):
    """training-style draws: per-ray jittered coarse positions [R,65] and random inverse-cdf draws [R,128] (unsorted: the
    in-kernel sampler takes its rank-sort path)."""
    lib, nerf = aon
    prec = lib.PREC_TC_F16X3
    sd, net, k, pc, pf, fc, ff, lat = _setup(lib, nerf, "vanilla", prec, dev)
    slots = torch.cuda.get_device_properties(dev).multi_processor_count // 2
    R = slots * 256 + 77
    rays = O.sapien_rays(120, 200, seed=3)
    o, d = rays["rays_o"][:R].contiguous().to(dev), rays["rays_d"][:R].contiguous().to(dev)
    g = torch.Generator().manual_seed(1)
    t_rand, u = torch.rand(R, 65, generator=g).to(dev), torch.rand(R, 128, generator=g).to(dev)
    t0 = lib.sample_along_rays(2.0, 6.0, 65, R, dev, t_rand=t_rand)
    c_ref, f_ref = _three_launch(lib, k, prec, pc, pf, fc, ff, o, d, t0, u=u)
    fine, coarse = lib.render_rays(k, prec, pc, pf, fc, ff, o, d, d, 2.0, 6.0, True, t_coarse=t0, u=u)
    assert torch.equal(coarse, c_ref) and torch.equal(fine, f_ref)
    with torch.no_grad():                                   # and through the module surface
        out = net({"rays_o": o, "rays_d": d, "viewdirs": d}, True, True, 2.0, 6.0, t_rand=t_rand, u=u)
    assert torch.equal(out[1][0], f_ref[:, :3]) and torch.equal(out[0][2], c_ref[:, 4])
    # against the oracle with the same draws (first rays only: CPU time)
    n = 64
    want = O.nerf_forward(sd, {kk: v[:n] for kk, v in rays.items()}, True, True, 2.0, 6.0, t_rand=t_rand[:n].cpu(), u=u[:n].cpu())
    for lv, got in ((0, coarse), (1, fine)):
        assert relerr(got[:n, :3].cpu(), want[lv][0]) < 1e-4 and relerr(got[:n, 3].cpu(), want[lv][1]) < 1e-4


@pytest.mark.parametrize("kind", ["vanilla", "autodecoder"])
def test_render_image_camera_and_pixel_blocks(aon, dev, kind):
    """aon_render_image: ray generation fused in.  Equal to raygen + aon_render_rays, and any partition of the pixel range
    (what the ranks of a sharded render compute) concatenates to the full image bit for bit."""
    lib, nerf = aon
    from aon_b200 import synth
    prec = lib.PREC_TC_F16X3
    sd, net, k, pc, pf, fc, ff, lat = _setup(lib, nerf, kind, prec, dev)
    H, W = 120, 160                                        # 19 200 rays = 75 CTA pairs: one full wave + a tail
    focal, c2w = synth.sapien_focal(H), synth.sapien_camera(4)
    o, d = lib.raygen(H, W, focal, c2w, dev)
    f_ref, c_ref = lib.render_rays(k, prec, pc, pf, fc, ff, o, d, d, 2.0, 6.0, True)
    fine, coarse = lib.render_image(k, prec, pc, pf, fc, ff, c2w, focal, H, W, 2.0, 6.0, True, want_coarse=True)
    assert torch.equal(fine, f_ref) and torch.equal(coarse, c_ref)
    for world in (2, 3, 8):
        from aon_b200.dist import shard_bounds
        parts = [lib.render_image(k, prec, pc, pf, fc, ff, c2w, focal, H, W, 2.0, 6.0, True, ray0=lo, R=hi - lo)[0]
                 for lo, hi in shard_bounds(H * W, world)]
        assert torch.equal(torch.cat(parts, 0), f_ref), world
    # unaligned block boundaries too
    parts = [lib.render_image(k, prec, pc, pf, fc, ff, c2w, focal, H, W, 2.0, 6.0, True, ray0=lo, R=hi - lo)[0]
             for lo, hi in ((0, 1), (1, 4097), (4097, H * W))]
    assert torch.equal(torch.cat(parts, 0), f_ref)


def test_full_size_image_is_deterministic_and_traffic_free(aon, dev):
    """640x480 (BASELINE configs[1]) through aon_render_image: five renders are bit-identical; properties hold."""
    lib, nerf = aon
    from aon_b200 import synth
    prec = lib.PREC_TC_F16X3
    sd, net, k, pc, pf, fc, ff, lat = _setup(lib, nerf, "vanilla", prec, dev)
    H, W = 480, 640
    focal, c2w = synth.sapien_focal(H), synth.sapien_camera(7)
    first, _ = lib.render_image(k, prec, pc, pf, fc, ff, c2w, focal, H, W, 2.0, 6.0, True)
    first = first.clone()
    for _ in range(4):
        again, _ = lib.render_image(k, prec, pc, pf, fc, ff, c2w, focal, H, W, 2.0, 6.0, True)
        assert torch.equal(again, first)
    assert torch.isfinite(first).all()
    assert (first[:, 3] >= 0).all() and (first[:, 3] <= 1 + 1e-5).all()
    assert (first[:, :3] >= -1e-5).all() and (first[:, :3] <= 1 + 1e-5).all()
    assert (first[:, 4] >= 0).all() and (first[:, 4] <= 6.0 + 1e-3).all()
    # a 3840-ray slice of the image against the oracle (the reference's own chunk size)
    o, d = lib.raygen(H, W, focal, c2w, dev)
    lo = 150 * W
    rays = {"rays_o": o[lo:lo + 3840].cpu(), "rays_d": d[lo:lo + 3840].cpu(), "viewdirs": d[lo:lo + 3840].cpu()}
    want = O.nerf_forward(sd, rays, False, True, 2.0, 6.0)
    floor = noise_floor("full_size_slice", sd, rays, None, True, want)
    got = first[lo:lo + 3840].cpu()
    for j, (a, nm) in enumerate(((got[:, :3], "rgb"), (got[:, 3], "acc"), (got[:, 4], "depth"))):
        e = relerr(a, want[1][j])
        assert e < max(1e-4, 5 * floor[1][j]), (nm, e, floor[1][j])


@pytest.mark.parametrize("R", [1, 127, 129])
def test_autodecoder_edge_rays_vs_oracle(aon, dev, R):
    lib, nerf = aon
    prec = lib.PREC_TC_F16X3
    sd, net, k, pc, pf, fc, ff, lat = _setup(lib, nerf, "autodecoder", prec, dev, art=7)
    from tests.test_gpu_edge import _axis_rays
    rays = {kk: v.repeat((R + 11) // 12, 1)[:R] + 0.0 for kk, v in _axis_rays().items()}
    rd = {kk: v.to(dev) for kk, v in rays.items()}
    with torch.no_grad():
        got = net(rd, False, True, 2.0, 6.0, lat)
    want = O.nerf_forward(sd, rays, False, True, 2.0, 6.0, latents={n: v.cpu() for n, v in lat.items()})
    for lv in range(2):
        for j, nm in enumerate(("rgb", "acc", "depth")):
            g = got[lv][j].cpu()
            assert torch.isfinite(g).all()
            assert relerr(g, want[lv][j]) < 1e-4, (R, lv, nm, relerr(g, want[lv][j]))


def test_autodecoder_saturated_and_empty_density(aon, dev):
    lib, nerf = aon
    sd = O.make_state_dict("autodecoder", 0, sharp=True)
    rays = O.sapien_rays(9, 16, seed=5)
    rd = {k: v.to(dev) for k, v in rays.items()}
    lat = O.code_library(sd, torch.tensor([0]), torch.tensor([3]), is_test=False)
    latd = {n: v.to(dev) for n, v in lat.items()}
    for bias, acc_want in ((1e4, 1.0), (-1e4, 0.0)):
        sd2 = {k: v.clone() for k, v in sd.items()}
        for m in ("coarse_mlp", "fine_mlp"):
            sd2[m + ".density_layer.weight"].zero_()
            sd2[m + ".density_layer.bias"].fill_(bias)
        n2 = _make_net(nerf, "autodecoder", sd2, dev)
        n2.precision = lib.PREC_TC_F16X3
        with torch.no_grad():
            got = n2(rd, False, True, 2.0, 6.0, latd)
        want = O.nerf_forward(sd2, rays, False, True, 2.0, 6.0, latents=lat)
        for lv in range(2):
            assert torch.allclose(got[lv][1].cpu(), torch.full_like(want[lv][1], acc_want), atol=1e-6)
            for j in range(3):
                assert torch.isfinite(got[lv][j]).all()
                assert relerr(got[lv][j].cpu(), want[lv][j]) < 1e-4
        if acc_want == 0.0:
            assert (got[1][0] == 1.0).all()          # white background only


def test_autodecoder_sharp_R3840_golden(aon, dev, golden_dir):
    """the case round 1 skipped: sharp densities, 3840 rays (the reference's chunk), auto-decoder, f16x3 -- end to end
    against the reference-generated golden, bar max(1e-4, 5 x the reference's own fp32 noise floor)."""
    lib, nerf = aon
    name = "autodecoder_sharp_R3840_wb1_art3.npz"
    g, kind, sd, rays, lat = _load_case(os.path.join(golden_dir, name))
    net = _make_net(nerf, kind, sd, dev)
    net.precision = lib.PREC_TC_F16X3
    with torch.no_grad():
        out = net({k: v.to(dev) for k, v in rays.items()}, False, True, 2.0, 6.0, {k: v.to(dev) for k, v in lat.items()})
    ref32 = [[_t(g["%s%d" % (nm, lv)]) for nm in ("rgb", "acc", "depth")] for lv in range(2)]
    floor = noise_floor(name, sd, rays, lat, True, ref32)
    for lv in range(2):
        for j, nm in enumerate(("rgb", "acc", "depth")):
            e = relerr(out[lv][j].cpu(), ref32[lv][j])
            assert e < max(1e-4, 5 * floor[lv][j]), (lv, nm, e, floor[lv][j])


def test_code_library_product_class_vs_golden_latents(aon, dev, golden_dir):
    """A10: nerf.CodeLibraryArticulated (the product class, incl. the 19-row test-time interpolation of
    models/code_library.py:55-71) against the latents the REFERENCE's CodeLibraryArticulated produced (stored in the
    auto-decoder goldens: art3 = training lookup, art7 = is_test interpolation, an odd row = mean of two learnt codes)."""
    lib, nerf = aon
    from types import SimpleNamespace
    sd = O.make_state_dict("autodecoder", 0, sharp=True)
    codes = nerf.CodeLibraryArticulated(SimpleNamespace(N_max_objs=1, N_obj_code_length=128))
    codes.load_state_dict({k[len("code_library."):]: v for k, v in sd.items() if k.startswith("code_library.")})
    codes = codes.to(dev)
    seen = set()
    for path in sorted(glob.glob(os.path.join(golden_dir, "autodecoder_sharp_R33_wb1_art*.npz"))):
        g = np.load(path)
        art, is_test = int(g["articulation_id"]), bool(g["is_test"])
        seen.add((art, is_test))
        batch = {"instance_id": torch.tensor([0], device=dev), "articulation_id": torch.tensor([art], device=dev)}
        with torch.no_grad():
            lat = codes(batch, is_test=is_test)
        for k in ("density", "color", "articulation"):
            assert torch.equal(lat[k].cpu(), _t(g["lat_" + k])), (path, k)
    assert (3, False) in seen and (7, True) in seen
    with torch.no_grad():
        tab = codes.get_interpolated_articulations()
    assert tab.shape == (19, 32)
    assert torch.equal(tab.cpu(), O.interpolated_articulations(sd["code_library.embedding_instance_articulation.weight"]))


def test_sample_pdf_golden_exact(aon, dev, golden_dir):
    """A7 against the reference's t_fine: <= 1e-6 abs (BASELINE bar; the kernel's weight sum follows ATen's AVX2 reduction
    order, its cumsum is sequential like torch.cumsum's) and identical merge ranks of the 128 drawn samples."""
    lib, _ = aon
    g = np.load(os.path.join(golden_dir, "sample_pdf.npz"))
    t_c, w = _t(g["t_coarse"]).to(dev), _t(g["weights"]).to(dev)
    tf = lib.sample_pdf(t_c, w, 128).cpu()
    want = _t(g["t_fine"])
    assert (tf - want).abs().max().item() <= 1e-6, (tf - want).abs().max().item()
    # merge ranks: position of every coarse t inside the sorted 193 (ties: first occurrence), ours == reference's
    tc = t_c.cpu()
    rank = lambda t: torch.searchsorted(t.contiguous(), tc.contiguous(), right=False)
    assert torch.equal(rank(tf), rank(want))
    assert torch.equal(tf, want), "t_fine differs in %d of %d entries" % ((tf != want).sum().item(), tf.numel())
