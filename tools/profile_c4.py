"""ncu driver for the blob kernels: config-4 scene (1024 spheres + textured triangles), depth 8.  usage: profile_c4.py [W H]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflaxman_b200 import capi, scenes as S
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
c = capi.Context(0)
c.load_scene(S.synthetic_scene(32, floor=S.synthetic_texture(1024, 1024, 11), skybox=S.synthetic_texture(2048, 1536, 7)))
c.set_seeds(12345, 12345); c.set_image_size(W, H)
for k, v in (("blob_wavefront", "RFX_BLOB_WAVEFRONT"), ("blob_smem_bvh", "RFX_BLOB_SMEM_BVH")):
    if os.environ.get(v):
        c.set_option(k, int(os.environ[v]))
out = torch.empty((2, H, W), dtype=torch.int32, device="cuda")
c.render_frames_device(capi.pack_cameras([S.default_camera()] * 2), 8, 1, out.data_ptr(), 0)
c.synchronize()
print(c.stats())
