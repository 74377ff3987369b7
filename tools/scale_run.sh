#!/bin/bash
# usage (under gpurun --gpus 8): bash tools/scale_run.sh [tag]   -> bench.py (headline + secondary: configs 3 and 5) on 1, 2, 4 and 8 GPUs
out=gpurun_out/${1:-r2_scale}; mkdir -p $out
run() { n=$1; shift; if [ "$n" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; fi; }
for n in 1 2 4 8; do
  run $n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $out/scale_n$n.json 2> $out/scale_n$n.err; echo "bench n=$n rc=$?"
  grep -h '^{' $out/scale_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; s=d['secondary']; c3=s['config3_split_8k_frame']
print('N=%d value %.0f e2e %.0f Mrays/s (%.0f frames/s); d2h ceiling %.1f GB/s, e2e at %.2f of it; c3 %.4f ms (1 GPU %.4f, speedup %.2f, eff %.3f); c5 %.2f ms' % (d['n_gpus'], d['value'], e['value'], e['frames_per_s'], e['d2h_copy_only_GBps'], e['frac_of_d2h_ceiling'],
  c3['ms_per_frame'], c3.get('single_gpu_ms_per_frame',0), c3.get('speedup_vs_1gpu',0), c3.get('strong_scaling_efficiency',0), s['config5_orbit_240_frames']['ms_per_path']))"
done
