#!/bin/bash
# round-2 visit 30: verification after the attention work: full GPU suite, memcheck smoke, default bench
O=gpurun_out/r02ah; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/memcheck.log 2>&1; echo "memcheck exit $?" >> $O/memcheck.log; tail -3 $O/memcheck.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read())
print("%.2f steps/s %.3f ms e2e %.2f launches %d frac_burst %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac_of_burst"]))
print(json.dumps(d.get("clocks")))
for k in d["roofline"].get("kernels", []): print(k["kernel"], k["shape"], "%.1f us frac %.3f"%(k["us"], k["frac"]))
PY
