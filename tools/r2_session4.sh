#!/bin/bash
# 8-GPU session: config 3 (8K frame split by strips, peer-store gather into GPU 0) with the staged 64-byte row-segment stores
# against the direct 16-byte stores (variant library), and strip heights 8 / 16 / 32
out=gpurun_out/${1:-r2_s4}; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
show() { grep -h '^{' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2: %.4f ms per frame, 1 GPU %.4f, speedup %.2f, eff %.3f' % (d['ms_per_frame'], d.get('single_gpu_ms_per_frame',0), d.get('speedup_vs_1gpu',0), d.get('strong_scaling_efficiency',0)))"; }
for n in ${2:-8}; do
run $n bench.py --gpus $n --workload config3 --steps 20 > $out/c3_staged_n$n.json 2> $out/c3_staged_n$n.err; show $out/c3_staged_n$n.json "n=$n staged peer"
RFX_LIB=gpurun_variants/no_staging.so run $n bench.py --gpus $n --workload config3 --steps 20 > $out/c3_direct_n$n.json 2> $out/c3_direct_n$n.err; show $out/c3_direct_n$n.json "n=$n direct peer"
RFX_C3_GATHER=local run $n bench.py --gpus $n --workload config3 --steps 20 > $out/c3_local_n$n.json 2> $out/c3_local_n$n.err; show $out/c3_local_n$n.json "n=$n local"
done
RFX_C3_STRIP_ROWS=32 run 8 bench.py --gpus 8 --workload config3 --steps 20 > $out/c3_staged_rows32.json 2> $out/c3_staged_rows32.err; show $out/c3_staged_rows32.json "n=8 staged peer, 32-row strips"
RFX_C3_STRIP_ROWS=8 run 8 bench.py --gpus 8 --workload config3 --steps 20 > $out/c3_staged_rows8.json 2> $out/c3_staged_rows8.err; show $out/c3_staged_rows8.json "n=8 staged peer, 8-row strips"
python -m pytest tests/test_sharding.py tests/test_gpu_fullsize.py -m gpu -x -q -k "split or config3" 2>&1 | tail -2
