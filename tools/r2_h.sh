# round 2, run h (2 GPUs): multi-device tests over NCCL / two real devices, torchrun bench with the strong-scaling block
set -x
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 1500 $O/bench_n2.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(d["value"], d["e2e"]["value"], json.dumps(d["strong"]))
except Exception as e:
    print("bench failed", e)
PY
