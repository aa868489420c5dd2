set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fk20.py tests/test_gpu_threads.py tests/test_gpu_recover.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 600 python tools/k5_sweep.py > $O/k5_sweep.jsonl 2> $O/k5_sweep.err; cat $O/k5_sweep.jsonl; tail -3 $O/k5_sweep.err
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()}, d.get("latency_1blob_ms"), d.get("latency_32blob_ms"), d.get("abi_single_blob"))
PY
