set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fk20.py -m gpu -x -q -k "alternative_forms" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for occ in 3 4; do
  EKZG_NTT_OCC=$occ timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_occ$occ.json 2> $O/bench_occ$occ.err
  python - <<PY
import json
d = json.loads(open("$O/bench_occ$occ.json").read().strip().splitlines()[-1])
print("K5 occ $occ", round(d["value"]), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
done
