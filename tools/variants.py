"""Builds kernel variants (compile-time knobs) into build/variants/*.so and, on a GPU box, times K2 for each.
usage: variants.py build | run"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "gpurun_variants")

VARIANTS = {
    # name: (defines, force_path)
    "base": ([], 1),
    "nocull": (["RFX_PRIMARY_CULL=0"], 1),      # every query of a path walks every object (the kernel before the primary screen bounds)
}


def build():
    from reflaxman_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    built = {}
    for name, (defs, *_) in VARIANTS.items():
        out = os.path.join(VDIR, name + ".so")
        key = tuple(defs)
        if key in built:
            if os.path.lexists(out):
                os.remove(out)
            os.symlink(os.path.basename(built[key]), out)   # same binary, different runtime knobs
        else:
            B.build(force=True, defines=defs, out=out)
            built[key] = out
        print("built", name)


def run_one(name):
    import torch
    from reflaxman_b200 import capi, scenes as S
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(12345, 12345); c.set_image_size(1920, 1080)
    c.force_path(VARIANTS[name][1])
    n = 8
    out = torch.empty((n, 1080, 1920), dtype=torch.int32, device="cuda")
    cams = capi.pack_cameras([S.default_camera()] * n)
    c.render_frames_device(cams, 20, 1, out.data_ptr(), 0); c.synchronize()
    c.enable_profiling(True); c.stats_reset()
    c.render_frames_device(cams, 20, 1, out.data_ptr(), 0); c.synchronize()
    st = c.stats()
    print(json.dumps({"variant": name, "k2_us": 1e3 * st["trace_kernel_ms"] / st["trace_kernels"], "rays": st["rays"] // n,
                      "checksum": int(out[0].to(torch.int64).sum().item())}))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "run":
        for name in VARIANTS:
            so = os.path.join(VDIR, name + ".so")
            if os.path.exists(so):
                subprocess.run([sys.executable, __file__, "one", name], env=dict(os.environ, RFX_LIB=so))
    else:
        run_one(sys.argv[2])
