"""Run one render level of the fused kernel on synthetic rays (for ncu captures / quick timing).
    python tools/prof_level.py [--precision f16x3] [--kind vanilla] [--rays 18944] [--samples 65] [--iters 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aon_b200 import lib as L, nerf, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="f16x3")
    ap.add_argument("--kind", default="vanilla")
    ap.add_argument("--rays", type=int, default=148 * 128)
    ap.add_argument("--samples", type=int, default=65)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--errflag", action="store_true", help="pinned host int that receives the barrier time-out code (readable after a trap)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    prec = L.PRECISIONS[a.precision]
    sd = synth.make_state_dict(a.kind, 0, True)
    net = (nerf.NeRF() if a.kind == "vanilla" else nerf.NeRF_AE_Art())
    net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("code_library.")})
    net = net.to(dev).eval()
    kind = net.coarse_mlp.KIND
    o, d = L.raygen(480, 640, synth.sapien_focal(480), synth.sapien_camera(0), dev)
    o, d = o[:a.rays].contiguous(), d[:a.rays].contiguous()
    packed = net._cache["coarse"].get(net.coarse_mlp, prec)
    folded = None
    if a.kind != "vanilla":
        z = torch.zeros(128, device=dev)
        folded = L.fold_latents(kind, prec, packed, z + 0.01, z - 0.01, torch.zeros(32, device=dev) + 0.02)
    t = torch.linspace(2.0, 6.0, a.samples, device=dev)
    err = None
    if a.errflag:
        err = torch.zeros(4, dtype=torch.int32).pin_memory()
        L._dbg["err"] = err          # pinned memory is device-addressable under UVA
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    ev[0].record()
    for i in range(a.iters):
        L.render_level(kind, prec, packed, folded, o, d, d, t, True, False)
        ev[i + 1].record()
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED: %s; barrier time-out code = %s" % (str(e).splitlines()[0], None if err is None else err.tolist()))
        return
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
    flop = {"vanilla": 1186816, "autodecoder": 1589760}[a.kind] * a.samples * a.rays
    if a.timeline:
        tlbuf = torch.zeros(3 * 4 * 28 * 4 + 64, dtype=torch.int64, device=dev)
        tl = tlbuf
        L.debug_set_timeline(tl)
        L.render_level(kind, prec, packed, folded, o, d, d, t, True, False)
        torch.cuda.synchronize()
        L.debug_set_timeline(None)
        stats = tlbuf[3 * 4 * 28 * 4:].cpu().tolist()
        tl = tlbuf[:3 * 4 * 28 * 4].view(3, 4, 28, 4).cpu()
        for me in range(2):
            f, c, d, t, tot, n = stats[me * 8:me * 8 + 6]
            print("issuer %d: total %d cycles, %d own stages (%.0f cycles / own stage); waits: weights-full %d (%.0f%%) chunk %d (%.0f%%) acc-empty %d (%.0f%%) ticket %d (%.0f%%)"
                  % (me, tot, n, tot / max(n, 1), f, 100.0 * f / max(tot, 1), c, 100.0 * c / max(tot, 1), d, 100.0 * d / max(tot, 1), t, 100.0 * t / max(tot, 1)))
        print("producer: total %d cycles, waiting for a free ring slot %d (%.0f%%)" % (stats[17], stats[16], 100.0 * stats[16] / max(stats[17], 1)))
        for s in range(1, 3):
            print("--- sample %d (clocks rel. to kernel start)" % s)
            print("enc: efree-wait-done %d  E published %d" % (tl[2, s, 0, 0], tl[2, s, 0, 1]))
            for ui in range(28):
                if tl[0, s, ui, 3] == 0:
                    break
                m, e = tl[0, s, ui], tl[1, s, ui]
                print("unit %2d  MMA: start %8d chunk0 %8d lastchunk %8d committed %8d | EPI: dfull %8d pub0 %8d done %8d"
                      % (ui, m[0], m[1], m[2], m[3], e[0], e[1], e[2]))
    print("precision %s kind %s rays %d samples %d: ms %s  -> %.1f algorithmic TFLOP/s" %
          (a.precision, a.kind, a.rays, a.samples, ["%.2f" % m for m in ms], flop / (min(ms) * 1e-3) / 1e12))


if __name__ == "__main__":
    main()
