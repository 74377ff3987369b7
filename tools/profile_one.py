"""Small driver for ncu: renders a few config-2 frames (1920x1080, depth 20) through the C ABI.  usage: profile_one.py [frames] [W H]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from reflaxman_b200 import capi, scenes as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
c = capi.Context(0)
c.load_scene(S.default_scene())
c.set_seeds(12345, 12345)
c.set_image_size(W, H)
out = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
c.render_frames_device(capi.pack_cameras([S.default_camera()] * n), 20, 1, out.data_ptr(), 0)
c.synchronize()
print(c.stats())
