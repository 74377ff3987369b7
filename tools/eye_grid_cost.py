"""What the per-camera screen grid of blob scenes costs on the host: config-4 scene, 3840x2160, depth 1, 32 frames through
rfx_render_frames_device with one camera repeated (the grid is built once) against 32 different cameras (built per frame);
wall clock around the call + synchronize.  usage: eye_grid_cost.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from reflaxman_b200 import capi, scenes as S  # noqa: E402

W, H, N = 3840, 2160, 32
c = capi.Context(0)
c.load_scene(S.synthetic_scene(32, floor=S.synthetic_texture(1024, 1024, 11), skybox=S.synthetic_texture(2048, 1536, 7)))
c.set_seeds(12345, 12345); c.set_image_size(W, H)
out = torch.empty((N, H, W), dtype=torch.int32, device="cuda")
same = capi.pack_cameras([S.default_camera()] * N)
orbit = capi.pack_cameras(S.orbit_cameras(240)[:N])
res = {}
for depth in (1, 8):
    for name, cams in (("same_camera", same), ("orbit_cameras", orbit)):
        for on in (1, 0):
            c.set_option("eye_grid", on)
            c.render_frames_device(cams, depth, 1, out.data_ptr(), 0); c.synchronize()
            t0 = time.perf_counter()
            c.render_frames_device(cams, depth, 1, out.data_ptr(), 0); c.synchronize()
            res["d%d_%s_eye%d_ms_per_frame" % (depth, name, on)] = round(1e3 * (time.perf_counter() - t0) / N, 4)
print(json.dumps(res))
