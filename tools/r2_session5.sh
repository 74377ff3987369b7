#!/bin/bash
# 1-GPU: tests after the blob rewrite (wavefront pair, general kernel, rays kernel), config-4 sweep wavefront vs tile kernel
out=gpurun_out/${1:-r2_s5}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY' 2>&1 | tee $out/c4_wave.txt
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_extra
env = bench_extra.Env()
for wave in (2, 0):
    os.environ["RFX_BLOB_WAVEFRONT"] = str(wave)
    r = bench_extra.config4(env, steps=5)
    print("wavefront pair" if wave else "tile kernel   ", " ".join("d%d=%.3f" % (x["depth"], x["ms_per_frame"]) for x in r["sweep"]))
PY
