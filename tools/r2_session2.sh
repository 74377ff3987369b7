#!/bin/bash
# round-2 multi-GPU session (gpurun --gpus 8): bench.py (with its secondary object: config 3 peer-store gather, config 5) on 2/4/8 GPUs,
# config 3 A/B with the gather buffer local instead of peer, e2e A/B with two copy streams, FP32 microbenchmark on GPU 0
out=gpurun_out/${1:-r2_s2}; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
tools/microbench_fp32 > $out/microbench.jsonl 2>&1; echo "microbench rc=$?"; cat $out/microbench.jsonl
nvidia-smi topo -m > $out/topo.txt 2>&1
for n in 2 4 8; do
  run $n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench n=$n rc=$?"
  grep -h '^{' $out/bench_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; s=d['secondary']
print('N=%d value %.0f e2e %.0f Mrays/s; d2h ceiling %.1f GB/s frac %.2f; c3 %.3f ms (1gpu %.3f, eff %.2f); c5 %.2f ms' % (d['n_gpus'], d['value'], e['value'], e['d2h_copy_only_GBps'], e['frac_of_d2h_ceiling'],
  s['config3_split_8k_frame']['ms_per_frame'], s['config3_split_8k_frame'].get('single_gpu_ms_per_frame',0), s['config3_split_8k_frame'].get('strong_scaling_efficiency',0), s['config5_orbit_240_frames']['ms_per_path']))"
  RFX_C3_GATHER=local run $n bench.py --gpus $n --workload config3 --steps 10 > $out/c3_local_n$n.json 2> $out/c3_local_n$n.err; echo "c3 local n=$n rc=$?"
  grep -h '^{' $out/c3_local_n$n.json | cut -c1-400
done
RFX_COPY_STREAMS=2 run 8 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/bench_n8_2cs.json 2> $out/bench_n8_2cs.err; echo "bench n=8 two copy streams rc=$?"
grep -h '^{' $out/bench_n8_2cs.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('N=8, 2 copy streams: e2e %.0f Mrays/s frac of ceiling %.2f' % (e['value'], e['frac_of_d2h_ceiling']))"
