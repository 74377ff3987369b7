#!/bin/bash
# ncu --set full of the three blob kernels on config 4 (3840x2160, depth 8): wavefront first / rest, and the single tile kernel
out=gpurun_out/${1:-r2_s7}; mkdir -p $out
python tools/profile_c4.py > $out/plain.log 2>&1 || { cat $out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_blob_wave_first -s 1 -c 1 -o $out/prof_wave_first -f python tools/profile_c4.py > $out/ncu1.log 2>&1; tail -1 $out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:k_blob_wave_rest -s 1 -c 1 -o $out/prof_wave_rest -f python tools/profile_c4.py > $out/ncu2.log 2>&1; tail -1 $out/ncu2.log
RFX_BLOB_WAVEFRONT=0 ncu --set full --clock-control none --import-source on -k regex:k_trace_blob -s 1 -c 1 -o $out/prof_tile -f python tools/profile_c4.py > $out/ncu3.log 2>&1; tail -1 $out/ncu3.log
ls -la $out
