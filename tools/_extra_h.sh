O=gpurun_out/r01h; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench1.json 2> $O/bench1.err; cat $O/bench1.json | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > $O/mgpu.log 2>&1; tail -3 $O/mgpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 > $O/bench2.json 2> $O/bench2.err; cat $O/bench2.json | cut -c1-200; tail -3 $O/bench2.err
