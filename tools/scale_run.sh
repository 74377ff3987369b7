#!/bin/bash
# usage (under gpurun --gpus N): bash tools/scale_run.sh N [tag]   -> bench.py and configs 3 and 5 on N GPUs
N=$1; out=gpurun_out/${2:-r1_scale2}; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
if [ "$N" = "1" ]; then run() { python "$@"; }; fi
run bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $out/scale_n$N.json 2> $out/scale_n$N.err; echo bench $?
run bench.py --gpus $N --workload config3 --steps 5 > $out/c3_n$N.json 2> $out/c3_n$N.err; echo c3 $?
run bench.py --gpus $N --workload config5 --steps 2 > $out/c5_n$N.json 2> $out/c5_n$N.err; echo c5 $?
grep -h -v NCCL $out/scale_n$N.json | cut -c1-300; grep -h -v NCCL $out/c3_n$N.json $out/c5_n$N.json | cut -c1-400
