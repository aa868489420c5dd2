# round 2, run f: full GPU suite on both table layouts, the new bench line, native load generator sweeps
set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err; python - <<PY
import json
try:
    d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "latency_1blob_ms", "latency_32blob_ms") if k in d}, d["e2e"]["value"], d["roofline"]["frac"], d["strong"]["value"])
    print(json.dumps(d.get("configs"))[:1500]); print(d.get("abi_single_blob")); print(d.get("cpu_baseline"))
except Exception as e:
    print("bench failed", e)
PY
L=rust-eth-kzg_b200/lib/abi_load
for cap in 256 512; do
  EKZG_COALESCE_MAX=$cap EKZG_TRACE_COALESCE=1 timeout 300 $L --threads 1024 --calls 8 > $O/abi_c${cap}_t1024.json 2> $O/abi_c${cap}_t1024.err; cat $O/abi_c${cap}_t1024.json; tail -12 $O/abi_c${cap}_t1024.err
done
for t in 64 256; do timeout 300 $L --threads $t --calls 8 > $O/abi_t$t.json 2>$O/abi_t$t.err; cat $O/abi_t$t.json; done
timeout 300 $L --threads 256 --calls 4 --mode recover > $O/abi_recover_t256.json 2>$O/abi_recover.err; cat $O/abi_recover_t256.json
