"""A/B of the pooled replay kernel (rfx_set_option "small_pool") against the tile kernel, one context per arm:
   python tools/pool_ab.py > gpurun_out/pool_ab.jsonl"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(size, depth, pool, n=16):
    import torch
    from reflaxman_b200 import capi, scenes as S
    W, H = size
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(W, H)
    c.set_option("small_pool", pool)
    out = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
    cams = capi.pack_cameras([S.default_camera()] * n)
    c.render_frames_device(cams, depth, 1, out.data_ptr(), 0); c.synchronize()
    c.enable_profiling(True); c.stats_reset()
    c.render_frames_device(cams, depth, 1, out.data_ptr(), 0); c.synchronize()
    st = c.stats()
    return {"size": size, "depth": depth, "small_pool": pool, "k2_us": 1e3 * st["trace_kernel_ms"] / st["trace_kernels"],
            "rays": st["rays"] // n, "checksum": int(out.to(torch.int64).sum().item() // n), "frames_equal": bool((out == out[0]).all().item()) if False else None}


if __name__ == "__main__":
    for size, depth in (((1920, 1080), 20), ((1920, 1080), 4), ((1024, 768), 20), ((3840, 2160), 20)):
        for rep in range(2):
            for pool in (1, 0):
                print(json.dumps(run(size, depth, pool)), flush=True)
