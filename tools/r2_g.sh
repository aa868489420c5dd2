# round 2, run g: full GPU suite on both table layouts, the new bench line
set -x
O=gpurun_out/r2g
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err; python - <<PY
import json
try:
    d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "latency_1blob_ms", "latency_32blob_ms") if k in d}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["peak"], d["strong"]["value"])
    print(json.dumps(d.get("configs"))[:3000]); print(d.get("abi_single_blob")); print(d.get("cpu_baseline"))
except Exception as e:
    print("bench failed", e)
PY
