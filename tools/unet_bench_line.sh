python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --unet-ops --no-also 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(os.environ.get('TAG',''), [(k[:14], round(v['ms'],4), round(v['frac_of_hbm_peak'],3)) for k,v in d['also']['unet_ops']['ops'].items() if k.endswith('tf32')])
"
