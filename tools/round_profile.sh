#!/bin/bash
# Round-end evidence, run under gpurun on ONE B200:  bash tools/round_profile.sh <tag>
#   1. GPU tests   2. bench.py (ours, then the reference arm)   3. ncu launch list of the same bench command
#   4. ncu --set full capture of the dominant kernel (tools/profile_one.py)   5. clocks during the bench
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $out/clocks.csv &
smi=$!
python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
kill $smi
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"
python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline > $out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline > $out/ncu_launches.log 2>&1
python tools/profile_one.py 3 > $out/plain_profile_one.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_small -s 1 -c 1 -o $out/prof_k_trace_small -f python tools/profile_one.py 3 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
