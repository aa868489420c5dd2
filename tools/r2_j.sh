set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_verify.py tests/test_gpu_4844.py -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python tools/verify_trace.py 2>&1 | grep -v "compute_cells\|blob_to\|timeline" | tail -12
