O=gpurun_out/r01d; mkdir -p $O
timeout 300 python tools/dev_attn_check.py > $O/attn.log 2>&1; echo "attn exit $?" >> $O/attn.log
cat $O/attn.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/gemm -f python tools/gemm_cases.py geglu_l0 proj_l0_res qkv_l0 > $O/gemm_cases.log 2>&1
tail -4 $O/gemm_cases.log
ncu -i $O/gemm.ncu-rep --page raw --csv > $O/gemm_raw.csv 2>/dev/null
ncu -i $O/gemm.ncu-rep --page source --csv > $O/gemm_source.csv 2>$O/source.err
python tools/ncu_source_top.py 120 < $O/gemm_source.csv > $O/gemm_source_top.txt 2>&1
head -c 3000 $O/gemm_source.csv > $O/gemm_source_head.txt
gzip -9 $O/gemm_source.csv
ls -la $O
sz=$(stat -c %s $O/gemm.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 20000000 ]; then rm -f $O/gemm.ncu-rep; fi
