"""A/B: does the CTA barrier + atomic of the tile-class filing cost anything?  K2 time per 1080p frame with the cost classes recorded on
every launch (period 1) against every 8th / never again after the first (the launches in between replay the last recording and have no
barrier and no atomic at their end).  usage: python tools/tile_period.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflaxman_b200 import capi, scenes as S

for period in (1, 8, 1 << 20, 1, 8, 1 << 20):
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(1920, 1080)
    c.set_option("tile_order_period", period)
    n = 32
    out = torch.empty((n, 1080, 1920), dtype=torch.int32, device="cuda")
    cams = capi.pack_cameras([S.default_camera()] * n)
    for _ in range(2):
        c.render_frames_device(cams, 20, 1, out.data_ptr(), 0)
    c.synchronize()
    c.enable_profiling(True); c.stats_reset()
    c.render_frames_device(cams, 20, 1, out.data_ptr(), 0); c.synchronize()
    st = c.stats()
    c.enable_profiling(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        c.render_frames_device(cams, 20, 1, out.data_ptr(), 0)
    c.synchronize()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"tile_order_period": period, "k2_us": 1e3 * st["trace_kernel_ms"] / st["trace_kernels"], "step_us_per_frame": 1e3 * e0.elapsed_time(e1) / (10 * n),
                      "checksum": int(out[5].to(torch.int64).sum().item())}))
    c.close()
