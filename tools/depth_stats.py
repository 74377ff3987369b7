"""How much of the fast kernel's SIMT work is lanes waiting for longer paths in their warp?  Debug build (-DRFX_DEBUG_DEPTH)
writes per-pixel event counts instead of colours.  usage: depth_stats.py build | run"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "gpurun_variants", "depth.so")
if sys.argv[1] == "build":
    from reflaxman_b200 import build as B
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    B.build(force=True, defines=["RFX_DEBUG_DEPTH=1"], out=SO)
else:
    os.environ["RFX_LIB"] = SO
    import numpy as np
    from reflaxman_b200 import capi, scenes as S
    W, H = 1920, 1080
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(W, H)
    ev = c.render_frames([S.default_camera()], 20)[0]
    np.save(os.path.join(ROOT, "gpurun_out", "depth_events.npy"), ev.astype(np.uint32))
    b = (ev & 0xFFFF).astype(np.int64)          # bounce-loop iterations per pixel
    sh = (ev >> 16).astype(np.int64)            # shadow rays per pixel
    q = b + sh                                  # intersection queries per pixel
    # warps own 4x8 tiles; a CTA = 4 warps = 16x8 pixels
    def tiles(a, tw, th):
        return a.reshape(H // th, th, W // tw, tw).swapaxes(1, 2).reshape(H // th, W // tw, tw * th)
    wq = tiles(q, 4, 8)
    warp_queries = wq.max(axis=2).sum()         # the state machine runs max-over-lanes passes per warp
    lane_queries = q.sum() / 32.0
    cq = tiles(q, 16, 8)                        # CTA view: 128 lanes
    # ideal in-CTA compaction: at pass k the CTA needs ceil(active_lanes / 32) warps
    kmax = int(q.max())
    ideal = 0
    for k in range(kmax):
        act = (cq > k).sum(axis=2)
        ideal += np.ceil(act / 32.0).sum()
    # compaction only after pass T (cheap variant): passes <= T as today, later passes compacted
    res = {}
    for T in (2, 4, 6, 8):
        tot = np.minimum(wq.max(axis=2), T).sum()
        for k in range(T, kmax):
            act = (cq > k).sum(axis=2)
            tot += np.ceil(act / 32.0).sum()
        res["compact_after_%d" % T] = float(tot)
    print(json.dumps({"mean_bounces": float(b.mean()), "mean_shadow": float(sh.mean()), "warp_passes_now": float(warp_queries),
                      "lane_passes_div_32": float(lane_queries), "utilisation_now": float(lane_queries / warp_queries),
                      "warp_passes_ideal_cta_compaction": float(ideal), **res,
                      "hist_bounces": np.bincount(b.ravel(), minlength=21).tolist()}))
