#!/bin/bash
# Round-end evidence, run under gpurun on ONE B200:  bash tools/round_profile.sh <tag>
#   1. GPU tests + smoke   2. bench.py (ours, then the reference arm)   3. ncu launch list of the same bench command
#   4. ncu --set full capture of the dominant kernel (tools/profile_one.py)   5. clocks during the bench
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log; tail -3 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; cat $out/smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $out/clocks.csv &
smi=$!
python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
kill $smi
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"
python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline --no-secondary > $out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline --no-secondary > $out/ncu_launches.log 2>&1
python tools/profile_one.py 3 > $out/plain_profile_one.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_small -s 1 -c 1 -o $out/prof_k_trace_small -f python tools/profile_one.py 3 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); r=d['roofline']; s=d['secondary']
print('value %.0f e2e %.0f Mrays/s; k2 warm %.4f ms cold %.4f ms; frac %.4f; c3 %.3f ms c5 %.2f ms c4 d8 %.3f ms' % (d['value'], d['e2e']['value'], r['kernel_ms_per_launch'], r['kernel_ms_per_launch_cold'], r['frac'],
  s['config3_split_8k_frame']['ms_per_frame'], s['config5_orbit_240_frames']['ms_per_path'], s['config4_1024_spheres_depth_sweep']['sweep'][-1]['ms_per_frame']))"
