for W in openfwi_b1 marmousi_b1; do for o in "" "--opt cluster_size=16 --opt cluster_rows=7" "--opt cluster_size=8 --opt cluster_rows=7" "--opt cluster_size=16 --opt cluster_rows=4"; do python bench.py --only-headline --workload $W --steps 20 --warmup 5 $o 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(d['config']['workload'], '$o', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in r['kernels'].items()}, d['config']['engine']['forward'])
" ; done; done
