"""GPU: edge cases of SURVEY.md 8c(3) through the tensor-core parity mode (f16x3), and BASELINE-size properties.

* axis-parallel rays, rays that miss everything (zero density -> acc = 0, white background), rays with saturated alpha,
  R = 0 / 1 / 127 / 129 (ragged CTA pair: the peer CTA of the last pair has no valid rays) against the CPU oracle;
* 640x480 (BASELINE.json configs[1]) through size-independent properties: finite outputs, acc in [0,1], rgb in [0,1]
  for a white background, depth in [0, far], fine samples sorted within [near, far], BIT equality of a full-image render
  with the concatenation of two part-image renders, and bit equality of five repeated renders (the two MMA issuer threads
  hand an issue ticket back and forth, so every accumulator sees its K steps in program order)."""
import numpy as np
import pytest
import torch

from oracle import ref_cpu as O
from tests.test_gpu_parity import _make_net, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_lib):
    from aon_b200 import lib, nerf
    dev = torch.device("cuda:0")
    sd = O.make_state_dict("vanilla", 0, sharp=True)
    net = _make_net(nerf, "vanilla", sd, dev)
    net.precision = lib.PREC_TC_F16X3
    return lib, net, sd, dev


def _axis_rays():
    o, d = [], []
    for ax in range(3):
        for sgn in (1.0, -1.0):
            v = torch.zeros(3); v[ax] = -sgn
            p = torch.zeros(3); p[ax] = 4.0 * sgn
            o.append(p); d.append(v)                       # through the origin along +-x, +-y, +-z
            o.append(p + torch.tensor([0.0, 0.0, 5.0]) * (ax != 2) + torch.tensor([5.0, 0.0, 0.0]) * (ax == 2)); d.append(v)  # misses
    o, d = torch.stack(o), torch.stack(d)
    return {"rays_o": o, "rays_d": d, "viewdirs": d.clone()}


@pytest.mark.parametrize("R", [0, 1, 12, 127, 129])
def test_edge_rays_vs_oracle(ctx, R):
    lib, net, sd, dev = ctx
    rays = _axis_rays()
    reps = (R + 11) // 12 if R else 0
    rays = {k: (v.repeat(max(reps, 1), 1)[:R] + 0.0) for k, v in rays.items()}
    rd = {k: v.to(dev) for k, v in rays.items()}
    with torch.no_grad():
        got = net(rd, False, True, 2.0, 6.0)
    if R == 0:
        assert all(x.shape[0] == 0 for lv in got for x in lv)
        return
    want = O.nerf_forward(sd, rays, False, True, 2.0, 6.0)
    for lv in range(2):
        for j, nm in enumerate(("rgb", "acc", "depth")):
            g = got[lv][j].cpu()
            assert torch.isfinite(g).all()
            assert relerr(g, want[lv][j]) < 1e-4, (R, lv, nm, relerr(g, want[lv][j]))


def test_saturated_and_empty_density(ctx):
    """density head forced to +/- large values: alpha = 1 at the first sample (acc = 1, depth = t_0) / alpha = 0 (acc = 0)."""
    lib, net, sd, dev = ctx
    from aon_b200 import nerf
    rays = O.sapien_rays(9, 16, seed=5)
    rd = {k: v.to(dev) for k, v in rays.items()}
    for bias, acc_want in ((1e4, 1.0), (-1e4, 0.0)):
        sd2 = {k: v.clone() for k, v in sd.items()}
        for m in ("coarse_mlp", "fine_mlp"):
            sd2[m + ".density_layer.weight"].zero_()
            sd2[m + ".density_layer.bias"].fill_(bias)
        n2 = _make_net(nerf, "vanilla", sd2, dev)
        n2.precision = lib.PREC_TC_F16X3
        with torch.no_grad():
            got = n2(rd, False, True, 2.0, 6.0)
        want = O.nerf_forward(sd2, rays, False, True, 2.0, 6.0)
        for lv in range(2):
            assert torch.allclose(got[lv][1].cpu(), torch.full_like(want[lv][1], acc_want), atol=1e-6)
            for j in range(3):
                assert torch.isfinite(got[lv][j]).all()
                assert relerr(got[lv][j].cpu(), want[lv][j]) < 1e-4
        if acc_want == 0.0:
            assert (got[1][0] == 1.0).all()          # white background only


def test_full_size_properties(ctx):
    lib, net, sd, dev = ctx
    H, W = 480, 640
    from aon_b200 import synth
    o, d = lib.raygen(H, W, synth.sapien_focal(H), synth.sapien_camera(7), dev)
    rays = {"rays_o": o, "rays_d": d, "viewdirs": d}
    with torch.no_grad():
        full = net(rays, False, True, 2.0, 6.0)
        R = H * W
        cut = 150016                                      # tile aligned, not pair aligned (1172 tiles)
        a = net({k: v[:cut].contiguous() for k, v in rays.items()}, False, True, 2.0, 6.0)
        b = net({k: v[cut:].contiguous() for k, v in rays.items()}, False, True, 2.0, 6.0)
    for lv in range(2):
        rgb, acc, depth = full[lv]
        assert torch.isfinite(rgb).all() and torch.isfinite(acc).all() and torch.isfinite(depth).all()
        assert (acc >= 0).all() and (acc <= 1 + 1e-5).all()
        assert (rgb >= -1e-5).all() and (rgb <= 1 + 1e-5).all()
        assert (depth >= 0).all() and (depth <= 6.0 + 1e-3).all()
        for j in range(3):
            # same rays, different CTA pairing / different split into fused waves and sample-segmented tail: identical bits
            # (a ray-sharded render must equal the 1-GPU render exactly)
            assert torch.equal(torch.cat([a[lv][j], b[lv][j]], 0), full[lv][j]), (lv, j)
    with torch.no_grad():
        for rep in range(4):                              # run-to-run determinism of the parity mode
            again = net(rays, False, True, 2.0, 6.0)
            for lv in range(2):
                for j in range(3):
                    assert torch.equal(again[lv][j], full[lv][j]), (rep, lv, j)
    # fine samples of the full image: sorted, inside [near, far]
    t0 = lib.sample_along_rays(2.0, 6.0, 65, R, dev)
    pc = net._cache["coarse"].get(net.coarse_mlp, net.precision)
    _, _, _, w0 = lib.render_level(0, net.precision, pc, None, o, d, d, t0, True, True)
    t1 = lib.sample_pdf(t0, w0, 128)
    assert (t1[:, 1:] >= t1[:, :-1]).all() and t1.min() >= 2.0 and t1.max() <= 6.0
    assert (w0 >= 0).all() and torch.allclose(w0.sum(-1), full[0][1], atol=1e-5)


@pytest.mark.parametrize("nseg", [0, 1, 3, 7])
def test_sample_segments_match_oracle(ctx, nseg):
    """Small ray batches are spread over the SMs by cutting every ray's sample range into segments (one CTA pair per
    tile and segment; the per-sample (alpha, rgb) are composited in order by a second kernel).  0 = automatic choice.  Both
    levels, incl. the per-sample weights, against the oracle; the segment count must not change a single bit."""
    lib, net, sd, dev = ctx
    rays = O.sapien_rays(15, 20, seed=6)                  # 300 rays: 2 CTA pairs, ragged
    rd = {k: v.to(dev) for k, v in rays.items()}
    want = O.nerf_forward(sd, rays, False, True, 2.0, 6.0)
    pc = net._cache["coarse"].get(net.coarse_mlp, net.precision)
    t0 = lib.sample_along_rays(2.0, 6.0, 65, 300, dev)
    lib.debug_force_segments(1)
    ref = lib.render_level(0, net.precision, pc, None, rd["rays_o"], rd["rays_d"], rd["viewdirs"], t0, True, True)
    lib.debug_force_segments(nseg)
    try:
        with torch.no_grad():
            got = net(rd, False, True, 2.0, 6.0)
        seg = lib.render_level(0, net.precision, pc, None, rd["rays_o"], rd["rays_d"], rd["viewdirs"], t0, True, True)
    finally:
        lib.debug_force_segments(0)
    for lv in range(2):
        for j in range(3):
            assert relerr(got[lv][j].cpu(), want[lv][j]) < 1e-4, (nseg, lv, j)
    for a, b in zip(seg, ref):                              # rgb, acc, depth, weights [R,65] of the coarse level
        assert torch.equal(a, b)


def test_tail_wave_split_matches_unsplit(ctx):
    """A batch of more ray tiles than CTA-pair slots renders its last, partly filled wave in a second launch with
    sample segments.  Same per-ray arithmetic in the same order: bit-equal to the single unsplit launch (both levels,
    per-sample coarse weights included)."""
    lib, net, sd, dev = ctx
    slots = torch.cuda.get_device_properties(dev).multi_processor_count // 2
    R = slots * 256 + 300                                   # one full wave + 2 ragged remainder tiles
    rays = O.sapien_rays(120, 200, seed=9)
    rd = {k: v[:R].contiguous().to(dev) for k, v in rays.items()}
    assert rd["rays_o"].shape[0] == R
    pc = net._cache["coarse"].get(net.coarse_mlp, net.precision)
    t0 = lib.sample_along_rays(2.0, 6.0, 65, R, dev)
    lib.debug_no_tail_split(True)
    try:
        with torch.no_grad():
            ref_full = net(rd, False, True, 2.0, 6.0)
        ref = lib.render_level(0, net.precision, pc, None, rd["rays_o"], rd["rays_d"], rd["viewdirs"], t0, True, True)
    finally:
        lib.debug_no_tail_split(False)
    with torch.no_grad():
        got_full = net(rd, False, True, 2.0, 6.0)
    got = lib.render_level(0, net.precision, pc, None, rd["rays_o"], rd["rays_d"], rd["viewdirs"], t0, True, True)
    for j, (a, b) in enumerate(zip(got, ref)):             # rgb, acc, depth, weights [R,65]
        assert torch.equal(a, b), (j, (a - b).abs().max().item())
    for lv in range(2):
        for j in range(3):
            assert torch.equal(got_full[lv][j], ref_full[lv][j]), (lv, j)
