timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
ONLY=conv2 SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 96 540 960 1 2>&1
ONLY=conv2 SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 54 720 1280 1 2>&1
for w in cfg2 cfg3 cfg4a; do python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['config']['workload'][:30], round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), round(d['e2e']['value'],1))"; done
