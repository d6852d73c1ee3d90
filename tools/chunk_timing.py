import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aon_b200 import lib as L, nerf, synth
dev = torch.device("cuda:0")
sd = synth.make_state_dict("vanilla", 0, True)
net = nerf.NeRF(); net.load_state_dict(sd); net = net.to(dev).eval(); net.precision = L.PREC_TC_F16X3
o, d = L.raygen(480, 640, synth.sapien_focal(480), synth.sapien_camera(0), dev)
for R in (2048, 3840, 76800):
    rays = {"rays_o": o[:R].contiguous(), "rays_d": d[:R].contiguous(), "viewdirs": d[:R].contiguous()}
    for force in (1, 0):
        L.debug_force_segments(force)
        with torch.no_grad():
            net(rays, False, True, 2.0, 6.0)
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(5):
                net(rays, False, True, 2.0, 6.0)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 5
        print("R %6d  segments %s : %.2f ms per full coarse+fine render -> %.0f rays/s" % (R, "1 (off)" if force else "auto", dt * 1e3, R / dt))
L.debug_force_segments(0)
# whole image: last partly filled wave of ray tiles as its own sample-segmented launch (on) vs one unsplit launch (off)
rays = {"rays_o": o, "rays_d": d, "viewdirs": d}
for prec in (L.PREC_TC_F16X3, L.PREC_TC_F16):
    net.precision = prec
    for off in (True, False, True, False):
        L.debug_no_tail_split(off)
        with torch.no_grad():
            for _ in range(2):
                net(rays, False, True, 2.0, 6.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(6):
                net(rays, False, True, 2.0, 6.0)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        print("640x480 prec %d  tail split %s : %.2f ms -> %.0f rays/s" % (prec, "off" if off else "on ", ms, 307200 / ms * 1e3))
L.debug_no_tail_split(False)
