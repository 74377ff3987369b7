#!/bin/bash
# config-4 depth sweep (1..8) for a list of libraries / settings, one line each:  tools/c4_ab.sh "name:lib:ENV=..;ENV=.." ...
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; lib=${rest%%:*}; envs=${rest#*:}
  (
    [ -n "$lib" ] && [ "$lib" != "-" ] && export RFX_LIB=$lib
    IFS=';' read -ra kv <<< "$envs"; for e in "${kv[@]}"; do [ -n "$e" ] && export "$e"; done
    python - <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_extra
env = bench_extra.Env()
r = bench_extra.config4(env, steps=5)
print(" ".join("d%d=%.3f" % (x["depth"], x["ms_per_frame"]) for x in r["sweep"]))
PY
  ) | sed "s/^/$name  /"
done
