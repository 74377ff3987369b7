"""ncu driver for the general blob kernel: config-4 scene (1024 spheres + textured triangles) at 1920x1080, depth 8."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflaxman_b200 import capi, scenes as S
c = capi.Context(0)
c.load_scene(S.synthetic_scene(32, floor=S.synthetic_texture(1024, 1024, 11), skybox=S.synthetic_texture(2048, 1536, 7)))
c.set_seeds(12345, 12345); c.set_image_size(1920, 1080)
out = torch.empty((2, 1080, 1920), dtype=torch.int32, device="cuda")
c.render_frames_device(capi.pack_cameras([S.default_camera()] * 2), 8, 1, out.data_ptr(), 0)
c.synchronize()
print(c.stats())
