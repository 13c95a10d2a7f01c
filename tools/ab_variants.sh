#!/bin/bash
# A/B of environment-selected kernel variants in ONE GPU-box visit: for every "NAME:ENV=VAL ENV=VAL" argument runs the
# GEMM suite (launch-weighted total of the step's conv_gemm shapes) and, unless SKIP_BENCH is set, bench.py.
#   tools/ab_variants.sh tag "default:" "tma_res:MD_EPI_TMA=2" "m256:MD_M256=1" "cg2:MD_CG2=1" "nopdl:MD_PDL=0"
TAG=${1:-ab}; shift
O=gpurun_out/$TAG
mkdir -p $O
for spec in "$@"; do
  name=${spec%%:*}
  envs=${spec#*:}
  echo "=== $name [$envs]"
  env $envs timeout 300 python tools/gemm_suite.py > $O/suite_$name.log 2>&1
  tail -1 $O/suite_$name.log
  if [ -z "$SKIP_BENCH" ]; then
    env $envs timeout 300 python bench.py --no-cpu > $O/bench_$name.json 2> $O/bench_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$name.json").read())
    print("  bench: %.2f steps/s  %.3f ms/step  e2e %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("  bench failed:", e)
PY
  fi
done
