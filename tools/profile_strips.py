"""ncu driver: one rank's share of a split 8K frame (config 3) on a single GPU: world 8, rank 3, 16-row strips, local gather buffer."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflaxman_b200 import capi, scenes as S, sharding as P
world, rank = int(os.environ.get("RFX_WORLD", "8")), int(os.environ.get("RFX_RANK", "3"))
c = capi.Context(0)
c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(7680, 4320)
buf = c.buffer_alloc(7680 * 4320 * 4)
for _ in range(4):
    P.split_frame(c, S.default_camera(), 20, 1, world, rank, buf, strip_rows=16)
c.synchronize()
print(c.stats())
