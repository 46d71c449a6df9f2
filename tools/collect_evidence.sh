#!/bin/bash
# GPU-box evidence run: full GPU suite, N=1 bench line, ncu launch list + full-set capture (CSV only: reports stay on the box), sanitizer logs.
# Run as: gpurun --timeout 3000 -- tools/collect_evidence.sh ; then summarise with tools/ncu_summary.py into profiles/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_r02f.log
timeout 400 python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
tail -c 400 gpurun_out/bench_r02f.json
B="python bench.py --reads 10000 --read-len 15000 --batch-size 64 --step-targets 500 --feature-threads 1 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02f.csv $B > gpurun_out/bench_under_ncu_f.log 2>&1
timeout 700 ncu --set full --clock-control none -s 160 -c 60 -f -o /tmp/prof_f $B > gpurun_out/ncu_full_f.log 2>&1
ncu -i /tmp/prof_f.ncu-rep --page raw --csv > gpurun_out/prof_r02f_raw.csv 2>/dev/null; rm -f /tmp/prof_f.ncu-rep
ls -la gpurun_out/prof_r02f_raw.csv gpurun_out/launches_r02f.csv
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "submit_target_equals or launch_batching or host_harness" > gpurun_out/sanitizer_memcheck_f.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_f.log
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "submit_target_equals" > gpurun_out/sanitizer_racecheck_f.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_f.log
du -sh gpurun_out
