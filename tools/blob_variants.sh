#!/bin/bash
# config-4 timing of the blob batch kernel variants built into gpurun_variants/ (names on the command line)
for v in "$@"; do
  echo -n "$v: "; RFX_LIB=gpurun_variants/$v.so bash tools/c4_quick.sh
done
