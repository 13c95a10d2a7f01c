#!/bin/bash
# round-2 visit 22: full GPU suite (incl. training forward loss, latent-64 decode), smoke(), default bench, reference arm
O=gpurun_out/r02w; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ) 2> $O/bench_time.txt; tail -3 $O/bench_time.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read())
print("%.2f steps/s %.3f ms e2e %.2f launches %d frac_burst %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac_of_burst"]))
print(json.dumps(d.get("vae_decode"))[:200]); print(json.dumps(d.get("clocks")))
for k in d["roofline"].get("kernels", []): print(k["kernel"], k["shape"], "%.1f us frac %.3f"%(k["us"], k["frac"]))
r=json.loads(open("$O/bench_ref.json").read()); print("reference arm", r["value"], r["cpu_baseline"]["sample"])
PY
