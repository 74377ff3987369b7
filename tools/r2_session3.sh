#!/bin/bash
# quick 1-GPU check: GPU tests (incl. Pulse headless), fast-kernel time with the staged stores
out=gpurun_out/${1:-r2_s3}; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.log
python -m pytest tests/test_shim.py -m gpu -x -q -s -k pulse 2>&1 | grep -E "Pulse|screenshot|passed|failed" | head
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f e2e %.0f Mrays/s; k2 warm %.4f ms cold %.4f ms; frac %.4f' % (d['value'], d['e2e']['value'], r['kernel_ms_per_launch'], r['kernel_ms_per_launch_cold'], r['frac']))"
