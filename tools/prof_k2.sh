set -x
python tools/profile_one.py 3 > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace_small -s 1 -c 1 -o gpurun_out/prof_v6 -f python tools/profile_one.py 3 > gpurun_out/ncu_v6.log 2>&1
tail -3 gpurun_out/ncu_v6.log
