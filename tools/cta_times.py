"""Per-CTA timeline of one fast-kernel launch (debug build with -DRFX_CTA_TIMES): where does the tail come from?
usage: cta_times.py build | run"""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "gpurun_variants", "cta_times.so")
if sys.argv[1] == "build":
    from reflaxman_b200 import build as B
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    B.build(force=True, defines=["RFX_CTA_TIMES=1"], out=SO)
else:
    os.environ["RFX_LIB"] = SO
    import numpy as np
    import torch
    from reflaxman_b200 import capi, scenes as S
    W, H = 1920, 1080
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(W, H)
    out = torch.empty((2, H, W), dtype=torch.int32, device="cuda")
    cams = capi.pack_cameras([S.default_camera()] * 2)
    c.render_frames_device(cams, 20, 1, out.data_ptr(), 0); c.synchronize()
    n = (W // 16) * (H // 8)
    buf = np.zeros((n, 3), np.uint64)
    lib = C.CDLL(SO)
    rc = lib.rfx_debug_cta_times(buf.ctypes.data_as(C.c_void_p), n)
    t0 = buf[:, 0].min()
    start = (buf[:, 0] - t0).astype(np.float64) / 1e3
    end = (buf[:, 1] - t0).astype(np.float64) / 1e3
    dur = end - start
    total = end.max()
    sm_end = {}
    for e, s in zip(end, buf[:, 2]):
        sm_end[int(s)] = max(sm_end.get(int(s), 0), e)
    ends = np.array(sorted(sm_end.values()))
    gx = W // 16
    rows = dur.reshape(H // 8, gx)
    late = np.argsort(-end)[:10]
    print(json.dumps({"rc": rc, "ctas": n, "kernel_us": total, "cta_us_mean": dur.mean(), "cta_us_p50": float(np.median(dur)), "cta_us_p99": float(np.percentile(dur, 99)),
                      "cta_us_max": dur.max(), "sm_last_end_us_min": ends.min(), "sm_last_end_us_median": float(np.median(ends)), "sm_last_end_us_max": ends.max(),
                      "mean_sm_idle_tail_us": float((total - ends).mean()),
                      "last_ctas": [{"cta": int(i), "row": int(i // gx), "col": int(i % gx), "start": start[i], "dur": dur[i]} for i in late],
                      "row_mean_dur_us": [round(float(x), 1) for x in rows.mean(axis=1)][::6]}))
