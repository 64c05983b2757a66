#!/bin/bash
# One GPU session: tests, probes, bench lines, optionally (NCU=1) the ncu launch list + full capture of the conv kernel.
# usage: [NCU=1] [WL="cfg2 cfg4a"] [TAG=name] bash tools/gpu_round.sh
mkdir -p gpurun_out
TAG=${TAG:-run}
WL=${WL:-"cfg2 cfg4a"}
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
if [ "${RATE:-0}" = "1" ]; then python tools/gpu_diag.py rate > gpurun_out/rate_$TAG.log 2>&1; fi
for w in $WL; do
  EXTRA="--no-cpu-baseline"; [ "$w" = "cfg2" ] && EXTRA=""
  python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 $EXTRA > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
  cat gpurun_out/bench_${w}_$TAG.json
done
if [ "${REF:-0}" = "1" ]; then python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cat gpurun_out/bench_ref_$TAG.json; fi
if [ "${NCU:-0}" = "1" ]; then
  NW=${NCU_WL:-cfg4a}
  CMD="python bench.py --workload $NW --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
  $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-249} -c ${NCU_COUNT:-83} --csv --log-file gpurun_out/launches_${NW}_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
  echo "ncu launches rc=$?"
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s ${NCU_KSKIP:-170} -c 2 -o gpurun_out/prof_conv_${NW}_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
  echo "ncu full rc=$?"
fi
