set -x
O=gpurun_out/r2final
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 4), d["latency_1blob_ms"], d["latency_32blob_ms"], d["abi_single_blob"].get("blobs_per_s"), d["cpu_baseline"]["value"])
print({k: round(v["gpu_ms"], 2) for k, v in d["configs"]["1"]["per_item_symbols_ms"].items()})
for k in ("2", "4", "5"): print(k, round(d["configs"][k]["value"]), round(d["configs"][k]["ms"], 2))
PY
# launch list of a 16-blob call (latency mode: cooperative G1-NTT kernel) and abi_load at 64 / 256 callers
cat > /tmp/small_batch.py <<PY
import sys; sys.path.insert(0, ".")
import __graft_entry__, importlib
pkg = __graft_entry__.load_package(); syn = importlib.import_module("eth_kzg_b200.synthetic")
ctx = pkg.DASContext(use_precomp=True); flat = syn.blobs(16)
for _ in range(3): ctx.compute_cells_and_kzg_proofs_batch(flat, 16)
ctx.close()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_16blobs.csv python /tmp/small_batch.py > /dev/null 2>&1; grep -c "" $O/launches_16blobs.csv
for t in 16 64 256; do timeout 300 rust-eth-kzg_b200/lib/abi_load --threads $t --calls 16 2>/dev/null | tail -1 >> $O/abi_load.jsonl; done; cat $O/abi_load.jsonl
