timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
MZ_VERBOSE=1 SWEEP_CFGS="[dict(), dict(pair=1,a_stages=3,b_stages=5), dict(pair=1,a_stages=3,b_stages=6)]" timeout 300 python tools/sweep.py 96 540 960 1 2>&1 | awk '!seen[$0]++' | sed -E 's/\[mz conv\] mode//'
MZ_VERBOSE=1 SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 54 720 1280 1 2>&1 | awk '!seen[$0]++' | sed -E 's/\[mz conv\] mode//'
MZ_VERBOSE=1 SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 48 540 960 4 2>&1 | awk '!seen[$0]++' | sed -E 's/\[mz conv\] mode//'
for w in cfg3 cfg4a cfg4b; do python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['config']['workload'][:30], round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), round(d['e2e']['value'],1))"; done
