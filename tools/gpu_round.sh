#!/bin/bash
# One GPU session: tests, probes, bench lines, ncu launch list + full capture of the conv kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/gpu_diag.py rate > gpurun_out/rate.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --workload cfg4a --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4a.json 2> gpurun_out/bench_cfg4a.err; echo "bench cfg4a rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
CMD="python bench.py --workload cfg4a --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 249 -c 83 --csv --log-file gpurun_out/launches_cfg4a.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 170 -c 2 -o gpurun_out/prof_conv_cfg4a $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/bench_cfg2.json gpurun_out/bench_cfg4a.json gpurun_out/bench_ref.json
