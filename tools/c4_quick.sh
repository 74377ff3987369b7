#!/bin/bash
# quick config-4 timing (BVH path, depth 1/2/4/8) for kernel experiments: RFX_LIB selects the library
python bench.py --workload config4 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' '.join('d%d=%.2fms' % (x['depth'], x['ms_per_frame']) for x in d['sweep']))"
