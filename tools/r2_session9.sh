#!/bin/bash
out=gpurun_out/${1:-r2_s9}; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f e2e %.0f Mrays/s; ms/step %.4f; k2 warm %.4f ms cold %.4f ms; frac %.4f share %.4f; create %.1f ms' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['kernel_ms_per_launch_cold'], r['frac'], r['kernel_share_of_step'], d['context_create_ms']))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline --no-secondary > $out/ncu_launches.log 2>&1
python tools/ncu_launches.py $out/launches.csv
