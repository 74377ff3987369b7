"""Op census of BASELINE.json configs[3] (1024 spheres + textured floor/wall/sky) with the counting build of the oracle
(oracle/librfx_oracle_census.so: every ADD/SUB/MUL/DIV/SQRT of the REFERENCE's algorithm, i.e. its brute-force list walk over
all 1028 objects per ray), at a reduced resolution (the per-pixel figures depend on the resolution only through
pixel-footprint sampling).  usage: python tools/census_c4.py [W H] > profiles/census_c4_r2.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as O                     # noqa: E402
from reflaxman_b200 import scenes as S               # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 270)
O.build(ref=False)
L = O.lib(os.path.join(O.HERE, "librfx_oracle_census.so"))
assert L.rfxo_census_enabled()
scene = S.synthetic_scene(32, floor=S.synthetic_texture(1024, 1024, 11), skybox=S.synthetic_texture(2048, 1536, 7))
sc = O.OracleScene(scene, L)
out = {"workload": "config4: 1024 spheres + textured floor/wall/sky, default camera", "width": W, "height": H, "seed": 12345, "per_depth": {}}
for depth in range(1, 9):
    r = O.OracleRender(sc, W, H, seed=12345, L=L).render(S.default_camera(), depth)
    k = r.counters
    n = W * H
    flop = k["op_add"] + k["op_mul"] + k["op_div"] + k["op_sqrt"]
    out["per_depth"][depth] = {"flop_per_pixel": flop / n, "add": k["op_add"] / n, "mul": k["op_mul"] / n, "div": k["op_div"] / n, "sqrt": k["op_sqrt"] / n,
                               "powf": k["op_powf"] / n, "rays_per_pixel": k["rays"] / n, "sphere_tests_per_pixel": k["sphere_tests"] / n,
                               "tri_tests_per_pixel": k["tri_tests"] / n}
print(json.dumps(out, indent=1))
