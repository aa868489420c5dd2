set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fk20.py tests/test_gpu_threads.py tests/test_gpu_recover.py tests/test_gpu_chunks.py tests/test_gpu_multi.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 600 python tools/bench_configs.py --only latency 2>&1 | tail -1 | cut -c1-420
EKZG_TRACE=1 timeout 300 python - <<'PY' 2>&1 | grep "timeline" | tail -8
import __graft_entry__ as g, importlib
pkg = g.load_package(); syn = importlib.import_module("eth_kzg_b200.synthetic")
ctx = pkg.DASContext(use_precomp=True); b = syn.blob(1)
for _ in range(3): ctx.compute_cells_and_kzg_proofs(b)
ctx.close()
PY
