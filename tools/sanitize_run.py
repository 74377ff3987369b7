"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): fast kernel with tile ordering, general kernel
(SSAA, block preview, additive), blob kernel with BVH, K1 ranking + skips, resolve.  usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reflaxman_b200 import capi, scenes as S

cams = S.orbit_cameras(7)[:3]
c = capi.Context(0)
c.load_scene(S.default_scene(skybox=S.synthetic_texture(64, 48, 5), floor=S.synthetic_texture(32, 32, 9)))
c.set_seeds(11, 22); c.set_image_size(150, 77)
a = c.render_frames(cams, 8)                      # fast kernel, F_ALL, tile ordering from frame 2 on
c.skip_samples(150 * 77)
for samples, additive in ((1, False), (2, False), (-3, False), (1, True), (1, True)):
    c.render(cams[0], 6, samples, additive)       # general kernel: every renderNext mode
img = c.read_argb()
c.close()
d = capi.Context(0)
d.load_scene(S.default_scene()); d.set_seeds(3, 4); d.set_image_size(128, 72)
b = d.render_frames(cams, 20)                     # fast kernel, lean instantiation
d.close()
e = capi.Context(0)
e.load_scene(S.synthetic_scene(12, floor=S.synthetic_texture(32, 32, 3))); e.set_seeds(5, 6); e.set_image_size(96, 54)
g = e.render_frames(cams[:2], 4)                  # blob kernel with BVH (144 spheres)
e.close()
print("ok", int(a.sum() % 1000003), int(img.sum() % 1000003), int(b.sum() % 1000003), int(g.sum() % 1000003))
