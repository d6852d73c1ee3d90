"""CPU: the oracle restatement against the committed golden vectors (outputs of the UNMODIFIED
reference, written by oracle/gen_golden.py).  Bit-exact: same ATen CPU ops, same image."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_cpu as O


def _t(x):
    return torch.from_numpy(np.asarray(x))


def test_raygen_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "raygen_24x32.npz"))
    d = O.get_ray_directions(int(g["H"]), int(g["W"]), float(g["focal"]))
    o, v, rd = O.get_rays(d, _t(g["c2w"]))
    assert torch.equal(o, _t(g["rays_o"])) and torch.equal(v, _t(g["viewdirs"])) and torch.equal(rd, _t(g["rays_d"]))
    # the reference's aliasing: returned rays_d IS the normalised viewdirs
    assert torch.equal(v, rd)
    assert torch.allclose(rd.norm(dim=-1), torch.ones(rd.shape[0]), atol=1e-6)


def test_pos_enc_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pos_enc.npz"))
    x = _t(g["x"])
    assert torch.equal(O.pos_enc(x, 0, 10), _t(g["enc10"]))
    assert torch.equal(O.pos_enc(x, 0, 4), _t(g["enc4"]))
    # channel order: identity, k-major sines, then shifted sines
    e = O.pos_enc(x, 0, 10)
    assert torch.equal(e[:, :3], x)
    assert torch.equal(e[:, 3 + 3 * 2 + 1], torch.sin(x[:, 1] * 4))
    assert torch.equal(e[:, 33 + 3 * 9 + 2], torch.sin(x[:, 2] * 512 + np.float32(0.5 * np.pi)))


def test_sample_pdf_golden_and_bracket_equivalence(golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_pdf.npz"))
    t_c, w = _t(g["t_coarse"]), _t(g["weights"])
    bins = 0.5 * (t_c[..., 1:] + t_c[..., :-1])
    s = O.sorted_piecewise_constant_pdf(bins, w[..., 1:-1], 128, False)
    assert torch.equal(s, _t(g["samples"]))
    assert torch.equal(O.sorted_piecewise_constant_pdf_bracket(bins, w[..., 1:-1], 128), _t(g["samples"]))
    z = torch.zeros(t_c.shape[0], 3)
    tf, _ = O.sample_pdf(bins, w[..., 1:-1], z, z, t_c, 128, False)
    assert torch.equal(tf, _t(g["t_fine"]))
    # properties: sorted, inside [near, far], contains the coarse points
    assert (tf[:, 1:] >= tf[:, :-1]).all()
    assert tf.min() >= t_c.min() and tf.max() <= t_c.max()
    # all-zero weights -> uniform over [bins_0, bins_63]
    assert torch.allclose(s[0], torch.linspace(float(bins[0, 0]), float(bins[0, -1]), 128), atol=1e-5)
    # deterministic u ends at exactly 1.0 in fp32 (helper.py:229)
    assert O.fine_u_table(128)[-1].item() == 1.0


@pytest.mark.parametrize("wb", [0, 1])
def test_volrend_golden(golden_dir, wb):
    g = np.load(os.path.join(golden_dir, "volrend_wb%d.npz" % wb))
    out = O.volumetric_rendering(_t(g["rgb"]), _t(g["sigma"]), _t(g["t"]), _t(g["dirs"]), bool(wb))
    for a, nm in zip(out, ("comp_rgb", "acc", "weights", "depth")):
        assert torch.equal(a, _t(g[nm])), nm
    comp, acc, w, depth = out
    assert (w >= 0).all() and (acc <= 1 + 1e-5).all()
    assert acc[0] == 0 and (comp[0] == float(wb)).all()      # empty ray: background only
    assert abs(acc[2].item() - 1.0) < 1e-6                     # last interval absorbs everything


def _golden_levels():
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(root, "*_R*_wb*.npz")))


@pytest.mark.parametrize("name", [n for n in _golden_levels() if "_R3840_" not in n])
def test_level_loop_golden_small(golden_dir, name):
    _check_level_file(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["vanilla_sharp_R3840_wb1.npz"])
def test_level_loop_golden_chunk(golden_dir, name):
    _check_level_file(os.path.join(golden_dir, name))


def _check_level_file(path):
    g = np.load(path)
    kind, sharp = str(g["kind"]), bool(g["sharp"])
    sd = O.make_state_dict(kind, seed=0, sharp=sharp)
    assert abs(sum(v.double().abs().sum().item() for v in sd.values()) - float(g["sd_checksum"])) < 1e-6, \
        "synthetic weights differ from the ones the goldens were made with"
    rays = {k: _t(g[k]) for k in ("rays_o", "rays_d", "viewdirs")}
    lat = None
    if kind != "vanilla":
        lat = O.code_library(sd, torch.tensor([0]), torch.tensor([int(g["articulation_id"])]), bool(g["is_test"]))
        for k in lat:
            assert torch.equal(lat[k], _t(g["lat_" + k]))
    st = {}
    with torch.no_grad():
        out = O.nerf_forward(sd, rays, False, bool(g["white_bkgd"]), float(g["near"]), float(g["far"]), latents=lat, stages=st)
    for lv in range(2):
        for j, nm in enumerate(("rgb", "acc", "depth")):
            assert torch.equal(out[lv][j], _t(g["%s%d" % (nm, lv)])), (nm, lv)
    for k, v in st.items():
        ref = _t(g[k])
        assert torch.equal(v[: ref.shape[0]], ref), k


def test_code_library_interpolation():
    sd = O.make_state_dict("autodecoder", 0)
    tab = sd["code_library.embedding_instance_articulation.weight"]
    it = O.interpolated_articulations(tab)
    assert it.shape == (19, 32)
    assert torch.equal(it[0::2], tab) and torch.equal(it[5], (tab[2] + tab[3]) / 2)


def test_layer_tables_match_param_counts():
    # 595 844 / 798 215 params per MLP
    assert sum(o * i + o for _, o, i in O.VANILLA_LAYERS) == 595844
    assert sum(o * i + o for _, o, i in O.AUTODECODER_LAYERS) == 798215


def test_product_synth_generators_match_oracle_copies(built_lib):
    """bench.py / smoke() build their scenes with aon_b200.synth; the parity tests with the oracle's own
    generators -- both must describe the same synthetic scene."""
    from aon_b200 import synth
    for kind in ("vanilla", "autodecoder"):
        a, b = synth.make_state_dict(kind, 0, True), O.make_state_dict(kind, 0, True)
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(synth.sapien_camera(3), O.sapien_camera(3))


@pytest.mark.parametrize("kind", ["vanilla", "autodecoder"])
def test_training_gradients_golden(kind, golden_dir):
    """The oracle's TRAINING path (randomized sampling with the stored draws, loss0 + loss1, torch autograd) against the
    reference's own autograd (oracle/gen_golden_train.py asserted bit equality of the loss and all 48 / 83 gradients in the
    build container): the GPU gradient tests check our kernels against exactly this autograd."""
    g = np.load(os.path.join(golden_dir, "train_%s_sharp_R33.npz" % kind))
    sd = O.make_state_dict(kind, 0, sharp=True)
    assert abs(sum(v.double().abs().sum().item() for v in sd.values()) - float(g["sd_checksum"])) < 1e-6 * float(g["sd_checksum"])
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rays = {k: _t(g[k]) for k in ("rays_o", "rays_d", "viewdirs")}
    lat = O.code_library(p, torch.tensor([0]), torch.tensor([3])) if kind == "autodecoder" else None
    out = O.nerf_forward(p, rays, True, True, 2.0, 6.0, latents=lat, t_rand=_t(g["t_rand"]), u=_t(g["u"]))
    loss = O.img2mse(out[0][0], _t(g["target"])) + O.img2mse(out[1][0], _t(g["target"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-7
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == {k for k, v in p.items() if v.grad is not None}
    for n, want in zip(names, g["grad_abs_sums"]):               # every gradient, by its sum of magnitudes
        got = p[n].grad.double().abs().sum().item()
        assert abs(got - float(want)) <= 1e-5 * max(float(want), 1e-12), n
    for key in g.files:                                          # a few whole gradients
        if key.startswith("grad/"):
            ref = _t(g[key])
            got = p[key[5:]].grad
            assert (got - ref).abs().max() <= 1e-5 * ref.abs().max().clamp_min(1e-12), key


def test_philox_known_answers():
    """oracle/philox.py against the known-answer vectors of the Philox authors' Random123 distribution
    (kat_vectors: philox4x32, 10 rounds) -- the generator csrc/sampling.cuh evaluates inside the sampling kernels."""
    from oracle import philox as P

    def run(ctr, key):
        return [int(x) for x in P.philox4x32_10([np.uint32(v) for v in ctr], key)]

    assert run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = P.uniform(seed=7, offset=3, stream=1, rows=257, cols=128)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.01 and abs(float(u.var()) - 1.0 / 12.0) < 0.005
    # streams, offsets and seeds are independent coordinates of the counter / key
    assert not np.array_equal(u, P.uniform(7, 3, 0, 257, 128)) and not np.array_equal(u, P.uniform(7, 4, 1, 257, 128))
    assert not np.array_equal(u, P.uniform(8, 3, 1, 257, 128))
    # a draw is a function of (row, column) only: sub-blocks agree
    assert np.array_equal(u[:33, :65], P.uniform(7, 3, 1, 33, 65))
