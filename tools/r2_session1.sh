#!/bin/bash
# round-2 first GPU session (one B200): tests, bench with the secondary object, reference arm, pow check, blob-compact A/B, sanitizer
out=gpurun_out/r2_s1; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log; tail -5 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log
python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-1500 $out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"; cut -c1-300 $out/bench_reference.json
tools/pow_check > $out/pow_check.jsonl 2>&1; echo "pow_check rc=$?"; cat $out/pow_check.jsonl
echo "== blob compact A/B"
bash tools/c4_quick.sh | tee $out/c4_base.txt
RFX_LIB=gpurun_variants/blob_compact.so bash tools/c4_quick.sh | tee $out/c4_compact.txt
RFX_LIB=gpurun_variants/blob_compact.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blob_batch_kernel or edge_cases" > $out/pytest_compact.log 2>&1; echo "compact pytest rc=$?"; tail -3 $out/pytest_compact.log
echo "== sanitizer"
timeout 600 compute-sanitizer --error-exitcode 9 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "resize or unaligned" > $out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 $out/sanitizer.log
