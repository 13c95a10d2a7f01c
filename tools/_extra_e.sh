O=gpurun_out/r01e; mkdir -p $O
for tag in "" ew12 ew8; do
  MD_BUILD_TAG=$tag timeout 300 python tools/gemm_suite.py > $O/suite_${tag:-ew16}.log 2>&1
  tail -1 $O/suite_${tag:-ew16}.log
done
