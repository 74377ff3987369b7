"""How close is the GPU float image to the reference's?  Renders config 1 / config 2 on the GPU and compares with the
unmodified reference binary (oracle/_ref/ref_render) bit for bit.  Prints one JSON line per config."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reflaxman_b200 import capi, scenes as S
from oracle import pyoracle as O

for (W, H) in ((1024, 768), (1920, 1080)):
    info, imgs = O.run_reference(W, H, refl=20, seed=12345)
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(W, H)
    c.render(S.default_camera(), 20)
    rgbf, argb = c.read_rgbf(), c.read_argb()
    c.close()
    same = np.all(rgbf.view(np.uint32) == imgs[0][0].view(np.uint32), axis=2)
    d = np.abs(((argb[..., None] >> np.array([16, 8, 0], np.uint32)) & 255).astype(int) - ((imgs[0][1][..., None] >> np.array([16, 8, 0], np.uint32)) & 255).astype(int)).max(axis=2)
    print(json.dumps({"config": "%dx%d depth 20 seed 12345 vs unmodified reference" % (W, H), "pixels": W * H,
                      "float_rgb_bit_identical_pixels": int(same.sum()), "float_rgb_bit_identical_fraction": float(same.mean()),
                      "max_abs_float_diff": float(np.abs(rgbf - imgs[0][0]).max()),
                      "argb_identical_fraction": float((d == 0).mean()), "argb_within_1lsb_fraction": float((d <= 1).mean()), "argb_max_lsb_diff": int(d.max())}))
