#!/bin/bash
# secondary workloads on one GPU: configs 3, 4 (BVH and brute force), 5
out=gpurun_out/${1:-r1_extra}; mkdir -p $out
python bench.py --workload config3 --steps 5 > $out/c3_n1.json 2> $out/c3_n1.err; echo c3 $?
python bench.py --workload config4 --steps 3 > $out/c4_n1_bvh.json 2> $out/c4_n1_bvh.err; echo c4 $?
python bench.py --workload config4 --steps 1 --frames 0 > $out/c4_n1_bruteforce.json 2> $out/c4_n1_bruteforce.err; echo c4b $?
python bench.py --workload config5 --steps 2 > $out/c5_n1.json 2> $out/c5_n1.err; echo c5 $?
tail -n 2 $out/*.err | tail -20
