set -x
mkdir -p gpurun_out/v4b
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/v4b/bench.json 2> gpurun_out/v4b/bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/v4b/bench_reference.json 2> gpurun_out/v4b/bench_reference.err
python tools/bench_configs.py > gpurun_out/v4b/configs.jsonl 2> gpurun_out/v4b/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v4b/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/v4b/launches.log 2>&1
cut -c1-260 gpurun_out/v4b/configs.jsonl
