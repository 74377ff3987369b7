"""K2 time per Scene::trace call on the Render-API path (general kernel: float image, any mode) next to the batch path (fast kernel)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflaxman_b200 import capi, scenes as S

W, H = 1920, 1080
c = capi.Context(0)
c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(W, H)
cam = S.default_camera()
res = {}
for samples in (1, 2, 4):
    c.render(cam, 20, samples, False)
    c.enable_profiling(True); c.stats_reset()
    c.render(cam, 20, samples, False); c.synchronize()
    st = c.stats(); c.enable_profiling(False)
    res["render_api_samples_%d" % samples] = {"k2_ms": st["trace_kernel_ms"], "launches": st["trace_kernels"], "ns_per_call": 1e6 * st["trace_kernel_ms"] / (W * H * samples * samples)}
out = torch.empty((4, H, W), dtype=torch.int32, device="cuda")
cams = capi.pack_cameras([cam] * 4)
for samples in (1, 2):
    c.render_frames_device(cams, 20, samples, out.data_ptr(), 0); c.synchronize()
    c.enable_profiling(True); c.stats_reset()
    c.render_frames_device(cams, 20, samples, out.data_ptr(), 0); c.synchronize()
    st = c.stats(); c.enable_profiling(False)
    res["batch_samples_%d" % samples] = {"k2_ms_per_frame": st["trace_kernel_ms"] / 4, "launches": st["trace_kernels"], "ns_per_call": 1e6 * st["trace_kernel_ms"] / 4 / (W * H * samples * samples)}
print(json.dumps(res))
