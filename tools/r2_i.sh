set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fk20.py tests/test_gpu_threads.py tests/test_gpu_recover.py tests/test_gpu_chunks.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python tools/k5_sweep.py 1,32,64,128,192,224,256,288,320 > $O/k5_sweep.jsonl 2> $O/k5_sweep.err; cat $O/k5_sweep.jsonl | cut -c1-200; tail -3 $O/k5_sweep.err
timeout 600 python tools/bench_configs.py --only latency 2>&1 | tail -1 | cut -c1-700
