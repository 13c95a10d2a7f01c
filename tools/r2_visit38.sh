#!/bin/bash
# round-2 visit 38: short-K GEMMs of the level-0 transformer blocks, single-CTA tiles vs CTA pairs
O=gpurun_out/r02aq; mkdir -p $O
timeout 300 python tools/shortk_ab.py 2>&1 | tee $O/shortk_ab.log
MD_EPI_TMA=0 timeout 300 python tools/shortk_ab.py 2>&1 | tee $O/shortk_ab_noepitma.log
