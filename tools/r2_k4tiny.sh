set -x
O=gpurun_out/r2k4
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fk20.py tests/test_gpu_4844.py tests/test_gpu_recover.py -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python tools/k5_sweep.py 1,8,16,32 > $O/k5_sweep.jsonl 2> $O/k5_sweep.err; tail -3 $O/k5_sweep.err; cat $O/k5_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['blobs'], d['k5'], round(d['step_ms'],2), d['stages_ms'], d['same_as_radix2'])
"
