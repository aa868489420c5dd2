# round 2, first GPU pass: parity of the shared-memory-operand K4/K5, then A/B against the register forms
set -x
O=gpurun_out/r2a
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for cfg in vm:vm vm:reg; do
  k4=${cfg%%:*}; k5=${cfg##*:}
  EKZG_K4=$k4 EKZG_K5=$k5 timeout 600 python bench.py --no-cpu-baseline > $O/bench_k4${k4}_k5${k5}.json 2> $O/bench_k4${k4}_k5${k5}.err
  python - <<E
import json
try:
    d = json.loads(open("$O/bench_k4${k4}_k5${k5}.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("$cfg failed", e); print(open("$O/bench_k4${k4}_k5${k5}.err").read()[-1500:])
E
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fk20_msm_vm|k_fk20_g1_ntts_vm' -s 2 -c 2 -o $O/prof_k4_k5_vm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/prof.log 2>&1
tail -2 $O/prof.log
