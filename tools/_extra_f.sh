O=gpurun_out/r01f; mkdir -p $O
for b in 24 128; do
  MD_BN_BIAS=$b timeout 300 python tools/gemm_suite.py > $O/suite_bias$b.log 2>&1
  tail -1 $O/suite_bias$b.log
done
