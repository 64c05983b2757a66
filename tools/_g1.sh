mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export MZ_NO_PDL=1; else unset MZ_NO_PDL; fi
  for w in cfg2 cfg3 cfg4a; do python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-also --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('no_pdl=$v', d['config']['workload'][:30], round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; done
done
