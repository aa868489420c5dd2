set -x
O=gpurun_out/r2pairs
mkdir -p $O
export EKZG_LIB=$PWD/rust-eth-kzg_b200/lib/ab_pairs.so
timeout 600 python -m pytest tests/test_gpu_fk20.py -m gpu -x -q -k "consensus_vectors or radix4 or batch_matches" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python tools/k5_sweep.py 1,32,128,256,512 > $O/k5_sweep.jsonl 2> $O/err; cat $O/k5_sweep.jsonl | cut -c1-200
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("pairs", round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
