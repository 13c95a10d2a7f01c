#!/bin/bash
# round-2 visit 16: end-to-end sample() test, compute-sanitizer memcheck over the kernel families (incl. round-2 additions)
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -x -q -k "end_to_end or shell" > $O/pytest_e2e.log 2>&1; tail -8 $O/pytest_e2e.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py --vae > $O/memcheck.log 2>&1; echo "memcheck exit $?" >> $O/memcheck.log; tail -8 $O/memcheck.log
