#!/bin/bash
# Multi-GPU visit, measurements only: the genome-sharded bench line and C3 (strong scaling).
# Usage: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi_final.sh TAG N'
TAG=${1:-r2m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N exit $?"
timeout 1200 $TR bench.py --config c3 --steps 6 --warmup 2 > $OUT/c3_n$N.json 2> $OUT/c3_n$N.err; echo "c3 exit $?"
grep -v "^\[W\|OMP_NUM\|^\*\*\*\|^$" $OUT/c3_n$N.err | tail -5
python - $OUT $N <<'EOF2'
import json, sys
def load(p):
    l = [x for x in open(p).read().split("\n") if x.startswith("{")]
    return json.loads(l[0]) if l else None
out, n = sys.argv[1], sys.argv[2]
d = load(f"{out}/bench_n{n}.json")
if d:
    print("bench", round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in d["config"]["phase_ms_per_step"].items()},
          "coll", round(d["config"]["collective_ms_per_step"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 2), d.get("parity"))
d = load(f"{out}/c3_n{n}.json")
if d:
    c = d["config"]
    print("c3", round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in c["ms_per_step_by_statistic"].items()},
          "coll", round(c["collective_ms_per_step"], 2), c["per_rank_plan_bytes_edge_diffs_engine_ms"], d["parity"])
    print(c["phases_ms_by_statistic_rank0"]); print(c["slowest_call_ms_by_statistic_rank0"])
EOF2
