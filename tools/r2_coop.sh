set -x
O=gpurun_out/r2coop
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fk20.py -x -q -k "radix4_equals_radix2" > $O/pytest.log 2>&1; tail -15 $O/pytest.log
timeout 300 python tools/k5_sweep.py 1,8,16,32,48,64,96,128 > $O/k5_sweep.jsonl 2> $O/k5_sweep.err; tail -3 $O/k5_sweep.err; cat $O/k5_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['blobs'], d['k5'], round(d['step_ms'],2), d['stages_ms']['K5'], d['same_as_radix2'])
"
