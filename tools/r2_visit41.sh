#!/bin/bash
# round-2 visit 41: verification after the sparse-conv and peer-exchange work: full GPU suite, smoke, memcheck + racecheck of the
# conditioning branch, default bench and the reference arm
O=gpurun_out/r02at; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py --volume > $O/memcheck.log 2>&1; echo "memcheck exit $?" >> $O/memcheck.log; tail -3 $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py --volume-only > $O/racecheck.log 2>&1; echo "racecheck exit $?" >> $O/racecheck.log; tail -4 $O/racecheck.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read())
print("%.2f steps/s %.3f ms e2e %.2f launches %d frac_burst %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac_of_burst"]))
print(json.dumps(d.get("clocks")), json.dumps(d.get("cpu_baseline")))
for k in d["roofline"].get("kernels", []): print(k["kernel"], k["shape"], "%.1f us frac %.3f"%(k["us"], k["frac"]))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
