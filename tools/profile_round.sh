set -x
mkdir -p gpurun_out/v4
python bench.py > gpurun_out/v4/bench.json 2> gpurun_out/v4/bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/v4/bench_reference.json 2> gpurun_out/v4/bench_reference.err
python tools/bench_configs.py > gpurun_out/v4/configs.jsonl 2> gpurun_out/v4/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v4/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/v4/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fk20_msm|k_fk20_g1_ntts' -s 2 -c 2 -o gpurun_out/v4/prof_k4_k5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/v4/prof.log 2>&1
tail -2 gpurun_out/v4/prof.log
cat gpurun_out/v4/configs.jsonl | cut -c1-400
