timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for cfg in "4 12" "8 12" "4 6" "2 12"; do set -- $cfg
timeout 200 python bench.py --steps $2 --warmup 3 --no-cpu-baseline --feature-threads $1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('$cfg', round(d['value']/1e6), round(d['e2e']['value']/1e6), 'busy', round(c['host_worker_busy_ms_per_launch'],2), 'wait', round(c['host_worker_gpu_wait_ms_per_launch'],2), 'launches', c['launches_in_e2e_region'], 'harness_s', round(c['harness_seconds'],4), 'submit_sum', round(c['submit_seconds_sum'],3), 'bp', round(c['submit_backpressure_ms_sum']), 'allocs', c['host_allocs_in_e2e_region'], round(c['host_alloc_ms_in_e2e_region'],1), 'phase', c['worker_phase_ms_per_launch'])"
done
