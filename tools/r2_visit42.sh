#!/bin/bash
# round-2 visit 42 (2 GPUs): the multi-GPU test in its three exchange modes (peer, NCCL, peer with a simulated attach failure)
O=gpurun_out/r02au; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k multi_gpu ) > $O/pytest_mgpu.log 2>&1; tail -6 $O/pytest_mgpu.log
