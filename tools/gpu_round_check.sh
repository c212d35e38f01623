#!/bin/bash
# One gpurun call that produces every piece of GPU evidence quoted in profiles/: GPU test suite, smoke(), bench (both arms), the ncu
# launch list and one full capture of the step kernel, compute-sanitizer memcheck / racecheck on a small workload.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh r2b'
# Everything lands in gpurun_out/<tag>_*; numbers printed under ncu or the sanitizer are never bench values.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $out/${tag}_pytest_gpu.log 2>&1
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 600 python bench.py --steps 100 --warmup 25 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; head -c 600 $out/${tag}_bench_n1.json; echo
if [ "$2" = "full" ]; then timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; head -c 300 $out/${tag}_bench_reference.json; echo; fi
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_probe.py > $out/${tag}_memcheck.log 2>&1; grep -E "ERROR SUMMARY|SANITIZE_PROBE_DONE" $out/${tag}_memcheck.log
# ncu: launch list of the bench command, then one full capture of a steady-state launch
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $out/${tag}_launches_bench.csv \
  python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-side-results > $out/${tag}_ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rsb_step -s 40 -c 1 -f -o $out/${tag}_step_kernel \
  python bench.py --steps 4 --warmup 6 --no-cpu-baseline --no-side-results > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_step_kernel.ncu-rep 2>/dev/null
# compute-sanitizer on every kernel path (small batch)
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_probe.py > $out/${tag}_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|SANITIZE_PROBE_DONE" $out/${tag}_racecheck.log
# stage timers of the three bench workloads (exploratory)
for c in c3 $( [ "$2" = "full" ] && echo c2 c4 ); do timeout 200 python tools/stage_probe.py $c > $out/${tag}_stage_probe_$c.txt 2>&1; done
tail -12 $out/${tag}_stage_probe_c3.txt
