"""Parity table: every golden case x precision mode, stage-wise (kernel fed the reference's t values) and end to end,
max relative error of rgb / acc / depth against the reference's fp32 outputs, next to the reference's own fp32
noise floor (distance from an fp64 evaluation).  Test infrastructure (imports oracle/).
    python tools/parity_report.py [--modes f16x3,f16] > profiles/rN_parity.md
"""
import argparse
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.test_gpu_parity import _load_case, _make_net, _t, noise_floor, relerr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="fp32,f16x3,f16")
    a = ap.parse_args()
    from aon_b200 import lib, nerf
    dev = torch.device("cuda:0")
    names = sorted(glob.glob(os.path.join("tests", "golden", "*_R*_wb*.npz")))
    print("| case | mode | stage-wise rgb/acc/depth (coarse ; fine) | end-to-end rgb/acc/depth (coarse ; fine) | fp32 noise floor fine rgb/depth |")
    print("|---|---|---|---|---|")
    for path in names:
        name = os.path.basename(path)
        g, kind, sd, rays, lat = _load_case(path)
        net = _make_net(nerf, kind, sd, dev)
        wb = bool(g["white_bkgd"])
        rd = {k: v.to(dev) for k, v in rays.items()}
        latd = None if lat is None else {k: v.to(dev) for k, v in lat.items()}
        ref32 = [[_t(g["%s%d" % (nm, lv)]) for nm in ("rgb", "acc", "depth")] for lv in range(2)]
        floor = noise_floor(name, sd, rays, lat, wb, ref32)
        n = _t(g["t0"]).shape[0]
        for mode in a.modes.split(","):
            prec = lib.PRECISIONS[mode]
            net.precision = prec
            k = net.coarse_mlp.KIND
            sw = []
            for lv, mlp in enumerate((net.coarse_mlp, net.fine_mlp)):
                lins = mlp.linears()
                packed = lib.pack_weights(k, prec, [l.weight for l in lins], [l.bias for l in lins])
                folded = None
                if latd is not None:
                    folded = lib.fold_latents(k, prec, packed, latd["density"], latd["color"], latd["articulation"])
                t = _t(g["t%d" % lv]).to(dev).contiguous()
                o, d, v = (rd[x][:n].contiguous() for x in ("rays_o", "rays_d", "viewdirs"))
                out = lib.render_level(k, prec, packed, folded, o, d, v, t, wb)
                sw.append([relerr(out[j].cpu(), ref32[lv][j][:n]) for j in range(3)])
            with torch.no_grad():
                out = net(rd, False, wb, 2.0, 6.0) if latd is None else net(rd, False, wb, 2.0, 6.0, latd)
            ee = [[relerr(out[lv][j].cpu(), ref32[lv][j]) for j in range(3)] for lv in range(2)]
            f = lambda rows: " ; ".join("/".join("%.1e" % x for x in r) for r in rows)
            print("| %s | %s | %s | %s | %.1e/%.1e |" % (name, mode, f(sw), f(ee), floor[1][0], floor[1][2]))


if __name__ == "__main__":
    main()
