mkdir -p gpurun_out
for NW in cfg2 cfg4a; do
  if [ $NW = cfg2 ]; then PER=43; else PER=83; fi
  CMD="python bench.py --workload $NW --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
  $CMD > gpurun_out/plain_${NW}_r13.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*PER)) -c $PER --csv --log-file gpurun_out/launches_${NW}_r13.csv $CMD > gpurun_out/ncu1_${NW}_r13.log 2>&1
  echo "ncu launches $NW rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((2*PER-3+20)) -c 2 -f -o gpurun_out/prof_conv_${NW}_r13 $CMD > gpurun_out/ncu2_${NW}_r13.log 2>&1
  echo "ncu full $NW rc=$?"
done
ncu --set full --clock-control none -k regex:"bicubic_kernel|stem_kernel" -c 4 -f -o gpurun_out/prof_small_r13 python tools/time_small.py 48 2 540 960 16 > gpurun_out/ncu3_r13.log 2>&1; echo "ncu small rc=$?"
