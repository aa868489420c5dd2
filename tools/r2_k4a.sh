# batched-affine K4: parity with the kernel forced on for every launch size, then timings
set -x
O=gpurun_out/r2e
mkdir -p $O
EKZG_K4A_MIN=0 timeout 900 python -m pytest tests -m gpu -x -q -k "fk20 or 4844 or chunks or recover" > $O/pytest_k4a_all.log 2>&1; tail -3 $O/pytest_k4a_all.log
for cfg in "1:0" "2:0" "4:0"; do
  sl=${cfg%%:*}; occ=${cfg##*:}
  EKZG_K4A_SLICES=$sl EKZG_K4A_OCC=$occ timeout 600 python bench.py --no-cpu-baseline > $O/bench_s${sl}_o${occ}.json 2> $O/bench_s${sl}_o${occ}.err
  python - <<EOF
import json
try:
    d = json.loads(open("$O/bench_s${sl}_o${occ}.json").read().strip().splitlines()[-1])
    print("slices $sl occ $occ", round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("$cfg failed", e); print(open("$O/bench_s${sl}_o${occ}.err").read()[-1500:])
EOF
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fk20_msm_affine' -s 2 -c 1 -o $O/prof_k4a python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/prof.log 2>&1
tail -2 $O/prof.log
