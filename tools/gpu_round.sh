#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list of one 16-view step, full captures of the top kernels.
# usage: [SKIP_TESTS=1] [SKIP_BENCH=1] [FULL=regex FULLN=n] [EXTRA="cmd"] tools/gpu_round.sh <tag>
TAG=${1:-x}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
fi
if [ -z "$SKIP_BENCH" ]; then
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
cat $O/bench.json; cat $O/bench_ref.json
fi
if [ -n "$EXTRA" ]; then
bash -c "$EXTRA" > $O/extra.log 2>&1
tail -${EXTRA_TAIL:-40} $O/extra.log
fi
if [ -z "$SKIP_LIST" ]; then
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py 16 > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv 40 > $O/launches_summary.txt 2>&1
head -30 $O/launches_summary.txt
fi
if [ -n "$FULL" ]; then
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$FULL" -c ${FULLN:-6} \
  -o $O/full -f python tools/profile_step.py 16 > $O/full.log 2>&1
ncu -i $O/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>> $O/full.log
ncu -i $O/full.ncu-rep --page details --csv > $O/full_details.csv 2>> $O/full.log
ncu -i $O/full.ncu-rep --page source --csv 2>> $O/full.log | gzip -9 > $O/full_source.csv.gz
sz=$(stat -c %s $O/full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 25000000 ]; then rm -f $O/full.ncu-rep; echo "full.ncu-rep dropped ($sz bytes)" >> $O/full.log; fi
fi
du -sh gpurun_out
