"""K2 time of config 2 with real texels (SURVEY 'variant B': seeded synthetic 2048x1536 skybox atlas + 1024x1024 floor texture),
i.e. the F_ALL instantiation of the fast kernel with the bilinear sampler.  usage: textured_quick.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflaxman_b200 import capi, scenes as S

for name, scene in (("checker (variant A)", S.default_scene()),
                    ("texels (variant B)", S.default_scene(skybox=S.synthetic_texture(2048, 1536, 7), floor=S.synthetic_texture(1024, 1024, 11)))):
    c = capi.Context(0)
    c.load_scene(scene); c.set_seeds(12345, 12345); c.set_image_size(1920, 1080)
    n = 8
    out = torch.empty((n, 1080, 1920), dtype=torch.int32, device="cuda")
    cams = capi.pack_cameras([S.default_camera()] * n)
    c.render_frames_device(cams, 20, 1, out.data_ptr(), 0); c.synchronize()
    c.enable_profiling(True); c.stats_reset()
    c.render_frames_device(cams, 20, 1, out.data_ptr(), 0); c.synchronize()
    st = c.stats()
    print(json.dumps({"scene": name, "k2_us": 1e3 * st["trace_kernel_ms"] / st["trace_kernels"], "rays_per_frame": st["rays"] // n}))
    c.close()
