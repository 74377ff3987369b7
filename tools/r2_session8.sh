#!/bin/bash
# tile_order_period A/B on the moving-camera path (config 5) and the headline
out=gpurun_out/${1:-r2_s8}; mkdir -p $out
for p in 1 4 8 16 64; do
  RFX_TILE_ORDER_PERIOD=$p python bench.py --workload config5 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config5 tile_order_period $p: %.3f ms per path' % d['ms_per_path'])" | tee -a $out/c5_period.txt
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f e2e %.0f Mrays/s; k2 warm %.4f ms cold %.4f ms; frac %.4f' % (d['value'], d['e2e']['value'], r['kernel_ms_per_launch'], r['kernel_ms_per_launch_cold'], r['frac']))"
