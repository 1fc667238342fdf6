#!/bin/bash
# Multi-GPU bench line only (peer-memory exchange), preceded by the in-process exchange tests.
TAG=${1:-r2mq}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_exchange" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N exit $?"
if [ "$3" == "nccl" ]; then
TSKB_BENCH_NCCL=1 TSKB_BENCH_SYNC_COLL=1 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n${N}_sync.json 2> $OUT/bench_n${N}_sync.err; echo "bench nccl sync exit $?"
fi
grep -v "^\[W\|OMP_NUM\|^\*\*\*" $OUT/bench_n$N.err | tail -8
python - $OUT $N <<'EOF2'
import json, sys
for f in (f"bench_n{sys.argv[2]}.json", f"bench_n{sys.argv[2]}_sync.json"):
    try:
        txt = open(sys.argv[1] + "/" + f).read()
        d = json.loads([l for l in txt.split("\n") if l.startswith("{")][0])
        print(f, "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "coll %.3f" % d["config"]["collective_ms_per_step"],
              {k: round(v, 3) for k, v in d["config"]["phase_ms_per_step"].items()}, "e2e %.4g" % d["e2e"]["value"], d["parity"])
        print(d["config"]["sharding"])
    except Exception as e:
        print(f, "failed", e)
EOF2
