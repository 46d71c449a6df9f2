for cfg in "32768 500" "65536 500" "65536 1000" "32768 250"; do set -- $cfg
HERRO_B200_CHUNK_POS=$1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --launch-targets $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('chunk $cfg', round(d['value']/1e6), round(d['e2e']['value']/1e6), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernels_ms_per_step'].items()}, 'allocs', c['host_allocs_in_e2e_region'], round(c['host_alloc_ms_in_e2e_region'],1))"
done
