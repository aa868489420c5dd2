set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_verify.py tests/test_gpu_4844.py tests/test_gpu_threads.py -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 python tools/bench_configs.py --only latency 2>&1 | tail -2
