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
