O=gpurun_out/r01j; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu > $O/bench_cg2.json 2> $O/bench_cg2.err; cut -c1-180 $O/bench_cg2.json
MD_CG2=0 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu > $O/bench_nocg2.json 2> $O/bench_nocg2.err; cut -c1-180 $O/bench_nocg2.json
MD_CG2=1 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu > $O/bench_cg2all.json 2> $O/bench_cg2all.err; cut -c1-180 $O/bench_cg2all.json
