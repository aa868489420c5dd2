# round 2 (8 GPUs): the torchrun bench line with the strong-scaling block and the one-process multi-device context
set -x
O=gpurun_out/r2n8
mkdir -p $O
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 1500 $O/bench_n$N.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_n$N.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(d["value"], d["e2e"]["value"], json.dumps(d["strong"]))
except Exception as e:
    print("bench failed", e)
PY
