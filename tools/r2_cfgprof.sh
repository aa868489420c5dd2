# launch lists (per-kernel device time) of BASELINE configs #2, #4, #5 and of the single-blob call: ncu --metrics gpu__time_duration.sum
set -x
O=gpurun_out/r2cfg
mkdir -p $O
for c in eip4844 recover verify latency; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_$c.csv python tools/bench_configs.py --only $c --reps 1 --parity 2 > $O/$c.log 2>&1
  tail -1 $O/$c.log | cut -c1-200
done
ls -la $O
