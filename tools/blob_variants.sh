#!/bin/bash
# config-4 timing of the blob batch kernel variants built into gpurun_variants/ (names on the command line), e.g. after
#   python -c "from reflaxman_b200 import build as B; B.build(force=True, defines=['RFX_BLOB_COMPACT=1'], out='gpurun_variants/compact.so')"
# run under gpurun:  RFX_LIB=gpurun_variants/compact.so python -m pytest tests -m gpu -q -k blob_batch; bash tools/blob_variants.sh compact
for v in "$@"; do
  echo -n "$v: "; RFX_LIB=gpurun_variants/$v.so bash tools/c4_quick.sh
done
