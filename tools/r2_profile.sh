# round 2 profiles: launch list of the bench command, full ncu captures of K4 (shared-memory-operand form), K5 (radix-2, wide
# batch) and K5 (radix-4, 32-blob batch)
set -x
O=gpurun_out/r2p
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fk20_msm_vm|k_fk20_g1_ntts' -s 2 -c 2 -o $O/prof_k4_k5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/prof.log 2>&1
tail -2 $O/prof.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fk20_g1_ntts_r4' -s 1 -c 1 -o $O/prof_k5_r4 python tools/k5_sweep.py 32 > $O/prof_r4.log 2>&1
tail -2 $O/prof_r4.log
ls -la $O
