set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fk20.py tests/test_gpu_verify.py tests/test_gpu_4844.py tests/test_gpu_recover.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
