#!/bin/bash
# config-4 sweep: queue policies of the wavefront pair (library variants) x segments rendered in tiles
out=gpurun_out/${1:-r2_s6}; mkdir -p $out
for lib in default; do
  if [ "$lib" != default ]; then export RFX_LIB=gpurun_variants/$lib.so; fi
  python - "$lib" <<'PY' 2>&1 | tee -a $out/c4_wave.txt
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_extra
env = bench_extra.Env()
for wave, stages, smem in ((2, 0, 1), (1, 0, 1), (3, 0, 1), (2, 0, 0), (2, 1, 0), (0, 0, 0)):
    os.environ["RFX_BLOB_WAVEFRONT"] = str(wave); os.environ["RFX_BLOB_WAVE_STAGES"] = str(stages); os.environ["RFX_BLOB_SMEM_BVH"] = str(smem)
    r = bench_extra.config4(env, steps=5)
    print(("wavefront, %d segment(s) in tiles, %s, bvh in %s" % (wave, "stage kernels" if stages else "queue kernel", "smem" if smem else "L1")) if wave else "tile kernel", " ".join("d%d=%.3f" % (x["depth"], x["ms_per_frame"]) for x in r["sweep"]))
PY
done
unset RFX_LIB
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "blob or config4 or bvh or edge" 2>&1 | tail -2
