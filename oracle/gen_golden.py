"""TEST INFRASTRUCTURE ONLY -- pins the oracle to the reference and writes tests/golden/*.npz.

Runs ONLY in the build container (needs /root/reference, read-only).  It imports the
UNMODIFIED reference modules behind six stub modules (pytorch_lightning,
piqa, kornia.create_meshgrid, matplotlib, imageio, torch_optimizer are absent from the image and
touch nothing on the arithmetic path except create_meshgrid), loads the deterministic synthetic
weights of ``oracle.ref_cpu.make_state_dict`` into the reference modules, runs the reference and
the restatement on the same inputs, asserts they agree BIT-FOR-BIT, and stores the reference's
outputs as golden vectors.

    python -m oracle.gen_golden            # from the repo root

Weights are not stored (they are regenerated from the seed by make_state_dict; torch's CPU
generator is deterministic for a given torch build -- the GPU box runs the same image); a
checksum of every state_dict is stored so a silent generator change is detected.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

from oracle import ref_cpu as O

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    """the UNMODIFIED reference of /root/reference behind the stubs of oracle/ref_import.py"""
    from oracle import ref_import
    return ref_import.import_reference(REF)


def checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def beq(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), "%s: oracle != reference, max abs diff %g" % (what, (a - b).abs().max().item())


def run_reference_staged(ref, net, rays, white_bkgd, near, far, latents=None):
    """Drive the reference's own functions level by level to expose per-stage tensors
    (the sequence of calls is that of model.py:147-199 / model_autodecoder.py:278-337)."""
    h = ref.helper
    st = {}
    out = []
    t_vals = weights = None
    for level in range(2):
        if level == 0:
            t_vals, samples = h.sample_along_rays(rays["rays_o"], rays["rays_d"], 64, near, far, False, False)
            mlp = net.coarse_mlp
        else:
            t_mids = 0.5 * (t_vals[..., 1:] + t_vals[..., :-1])
            t_vals, samples = h.sample_pdf(t_mids, weights[..., 1:-1], rays["rays_o"], rays["rays_d"],
                                           t_vals, 128, False)
            mlp = net.fine_mlp
        venc = h.pos_enc(rays["viewdirs"], 0, 4)
        if latents is None:
            raw_rgb, raw_sigma = mlp(h.pos_enc(samples, 0, 10), venc)
            rgb, sigma = net.rgb_activation(raw_rgb), net.sigma_activation(raw_sigma)
        else:
            raw_rgb, raw_sigma = mlp(samples, venc, latents)
            rgb = net.rgb_activation(raw_rgb) * (1 + 2 * net.rgb_padding) - net.rgb_padding
            sigma = net.sigma_activation(raw_sigma + net.density_bias)
        comp_rgb, acc, weights, depth = h.volumetric_rendering(rgb, sigma, t_vals, rays["rays_d"], white_bkgd)
        st["t%d" % level], st["raw_rgb%d" % level] = t_vals, raw_rgb
        st["raw_sigma%d" % level], st["weights%d" % level] = raw_sigma, weights
        out.append((comp_rgb, acc, depth))
    return out, st


def pick_rays(H, W, n, seed):
    rays = O.sapien_rays(H, W, seed)
    g = torch.Generator().manual_seed(seed + 100)
    idx = torch.randperm(H * W, generator=g)[:n].sort().values
    return {k: v[idx].contiguous() for k, v in rays.items()}


@torch.no_grad()
def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()

    # ---- A1/A2 ray generation ------------------------------------------------------------
    H, W = 24, 32
    focal = 0.5 * H / np.tan(np.radians(17.5))
    c2w = O.sapien_camera(3)
    d_ref = ref.get_ray_directions(H, W, focal)
    beq(O.get_ray_directions(H, W, focal), d_ref, "get_ray_directions")
    ro, vd, rd, _ = ref.get_rays(d_ref.clone(), c2w, output_view_dirs=True, output_radii=True)
    o_o, o_v, o_d = O.get_rays(O.get_ray_directions(H, W, focal), c2w)
    beq(o_o, ro, "rays_o"); beq(o_v, vd, "viewdirs"); beq(o_d, rd, "rays_d")
    np.savez_compressed(os.path.join(OUT, "raygen_24x32.npz"), H=H, W=W, focal=np.float64(focal),
                        c2w=c2w.numpy(), rays_o=ro.numpy(), viewdirs=vd.numpy(), rays_d=rd.numpy())

    # ---- A4 pos_enc / A7 sample_pdf adversarial cases -------------------------------------
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(257, 3, generator=g) * 2 - 1) * 6.5
    e10, e4 = ref.helper.pos_enc(x, 0, 10), ref.helper.pos_enc(x, 0, 4)
    beq(O.pos_enc(x, 0, 10), e10, "pos_enc10"); beq(O.pos_enc(x, 0, 4), e4, "pos_enc4")
    np.savez_compressed(os.path.join(OUT, "pos_enc.npz"), x=x.numpy(), enc10=e10.numpy(), enc4=e4.numpy())

    Rp = 96
    t_c = O.coarse_t_table(64, 2.0, 6.0).expand(Rp, 65).contiguous()
    w = torch.rand(Rp, 65, generator=g) ** 4
    w[0] = 0.0                                   # all-zero weights -> padding branch
    w[1] = 0.0; w[1, 30] = 1.0                   # delta pdf
    w[2] = 1e-9                                  # sub-eps weights
    w[3] = 0.0; w[3, 10:12] = 0.5; w[3, 40] = 0.7  # flat cdf segments
    w[4] = 0.0; w[4, 1] = 1.0                    # mass in first interior bin
    w[5] = 0.0; w[5, 63] = 1.0                   # mass in last interior bin
    w[6] = 0.0; w[6, 0] = 1.0; w[6, 64] = 1.0    # mass only in the dropped end weights
    t_j = t_c + (torch.rand(Rp, 65, generator=g) - 0.5) * 0.05   # jittered (still sorted) t for half the rays
    t_c[48:] = t_j[48:]
    bins = 0.5 * (t_c[..., 1:] + t_c[..., :-1])
    zero3 = torch.zeros(Rp, 3)
    s_ref = ref.helper.sorted_piecewise_constant_pdf(bins, w[..., 1:-1], 128, False)
    beq(O.sorted_piecewise_constant_pdf(bins, w[..., 1:-1], 128, False), s_ref, "pdf(mask)")
    beq(O.sorted_piecewise_constant_pdf_bracket(bins, w[..., 1:-1], 128), s_ref, "pdf(bracket)")
    tf_ref, _ = ref.helper.sample_pdf(bins, w[..., 1:-1], zero3, zero3, t_c, 128, False)
    beq(O.sample_pdf(bins, w[..., 1:-1], zero3, zero3, t_c, 128, False)[0], tf_ref, "sample_pdf")
    np.savez_compressed(os.path.join(OUT, "sample_pdf.npz"), t_coarse=t_c.numpy(), weights=w.numpy(),
                        samples=s_ref.numpy(), t_fine=tf_ref.numpy())

    # ---- A6 compositing adversarial cases -------------------------------------------------
    Rc, S = 64, 65
    rgb = torch.rand(Rc, S, 3, generator=g)
    sig = torch.relu(torch.randn(Rc, S, 1, generator=g) * 3.0)
    sig[0] = 0.0                                 # empty ray
    sig[1] = 0.0; sig[1, 20] = 1e4               # saturated alpha
    sig[2] = 0.0; sig[2, -1] = 1e-3              # only the 1e10-wide last interval
    tt = O.coarse_t_table(64, 2.0, 6.0).expand(Rc, S).contiguous()
    dirs = torch.nn.functional.normalize(torch.randn(Rc, 3, generator=g), dim=-1)
    for wb in (0, 1):
        c_ref = ref.helper.volumetric_rendering(rgb, sig, tt, dirs, bool(wb))
        c_o = O.volumetric_rendering(rgb, sig, tt, dirs, bool(wb))
        for a, b, nm in zip(c_o, c_ref, ("rgb", "acc", "weights", "depth")):
            beq(a, b, "volrend." + nm)
        np.savez_compressed(os.path.join(OUT, "volrend_wb%d.npz" % wb), rgb=rgb.numpy(), sigma=sig.numpy(),
                            t=tt.numpy(), dirs=dirs.numpy(), comp_rgb=c_ref[0].numpy(), acc=c_ref[1].numpy(),
                            weights=c_ref[2].numpy(), depth=c_ref[3].numpy())

    # ---- A5/A8/A9 full level loops ---------------------------------------------------------
    for kind in ("vanilla", "autodecoder"):
        for sharp in (False, True):
            sd = O.make_state_dict(kind, seed=0, sharp=sharp)
            if kind == "vanilla":
                net = ref.M.NeRF().eval()
                net.load_state_dict({k: v for k, v in sd.items()}, strict=True)
                lat_sets = [None]
            else:
                net = ref.MA.NeRF_AE_Art().eval()
                net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")}, strict=True)
                lib = ref.CodeLibraryArticulated(SimpleNamespace(N_max_objs=1, N_obj_code_length=128))
                lib.load_state_dict({k[len("code_library."):]: v for k, v in sd.items()
                                     if k.startswith("code_library.")}, strict=True)
                lat_sets = []
                for art_id, is_test in ((3, False), (7, True)):
                    batch = {"instance_id": torch.tensor([0]), "articulation_id": torch.tensor([art_id])}
                    import io, contextlib
                    with contextlib.redirect_stdout(io.StringIO()):
                        lat = lib(batch, is_test=is_test)
                    lo = O.code_library(sd, batch["instance_id"], batch["articulation_id"], is_test)
                    for k in lat:
                        beq(lo[k], lat[k], "code_library." + k)
                    lat_sets.append((art_id, is_test, lat))
            for R in (1, 33, 3840):
                rays = pick_rays(240, 320, R, seed=R)
                for wb in (1, 0):
                    if wb == 0 and R != 33:
                        continue
                    for li, lat_item in enumerate(lat_sets):
                        if li > 0 and R == 3840:
                            continue
                        lat = None if lat_item is None else lat_item[2]
                        if lat is None:
                            full = net(rays, False, bool(wb), 2.0, 6.0)
                        else:
                            full = net(rays, False, bool(wb), 2.0, 6.0, lat)
                        staged, st = run_reference_staged(ref, net, rays, bool(wb), 2.0, 6.0, lat)
                        ost = {}
                        mine = O.nerf_forward(sd, rays, False, bool(wb), 2.0, 6.0, latents=lat, stages=ost)
                        for lv in range(2):
                            for j, nm in enumerate(("rgb", "acc", "depth")):
                                beq(staged[lv][j], full[lv][j], "staged-vs-forward %s%d" % (nm, lv))
                                beq(mine[lv][j], full[lv][j], "%s R=%d %s%d" % (kind, R, nm, lv))
                        for k in st:
                            beq(ost[k], st[k], "%s R=%d stage %s" % (kind, R, k))
                        rec = {"kind": kind, "sharp": sharp, "R": R, "white_bkgd": wb, "near": 2.0, "far": 6.0,
                               "sd_checksum": np.float64(checksum(sd)),
                               "rays_o": rays["rays_o"].numpy(), "rays_d": rays["rays_d"].numpy(),
                               "viewdirs": rays["viewdirs"].numpy()}
                        for lv in range(2):
                            for j, nm in enumerate(("rgb", "acc", "depth")):
                                rec["%s%d" % (nm, lv)] = full[lv][j].numpy()
                        n_keep = R if R <= 33 else 64   # per-stage tensors for the first rays only
                        for k, v in st.items():
                            rec[k] = v[:n_keep].numpy()
                        if lat_item is not None:
                            rec["articulation_id"], rec["is_test"] = lat_item[0], lat_item[1]
                            for k, v in lat.items():
                                rec["lat_" + k] = v.numpy()
                        name = "%s_%s_R%d_wb%d%s.npz" % (kind, "sharp" if sharp else "smooth", R, wb,
                                                         "" if lat_item is None else "_art%d" % lat_item[0])
                        np.savez_compressed(os.path.join(OUT, name), **rec)
                        print("wrote", name, "acc mean %.3f" % full[1][1].mean().item())
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden dir: %.2f MB" % (tot / 1e6))


if __name__ == "__main__":
    main()
