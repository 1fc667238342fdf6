"""Strong-scaling check of the genome-sharded path on real GPUs (BASELINE configs[2] shape, scaled):
f2 / f3 / f4 / divergence over 8 sample sets on the C2 ARG, windows sharded over the ranks, per-window
partials summed with one NCCL all_reduce.  Rank 0 also computes the whole genome alone and compares.
  torchrun --nproc-per-node N tools/shard_check.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from tskit_b200.lowlevel import LLTreeSequence
from tskit_b200.sharding import ShardedTreeSequence

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
barrier = dist.barrier if world > 1 else (lambda: None)
t, W, _ = bench.load_workload("c2", rank, barrier)
windows = np.linspace(0, t.sequence_length, W + 1)
s = t.samples
sets = np.array_split(s, 8)
sizes = np.array([len(x) for x in sets], dtype=np.uint64)
flat = np.concatenate(sets).astype(np.int32)
calls = [("divergence", np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], dtype=np.int32)),
         ("f2", np.array([(0, 1), (2, 3), (4, 5), (6, 7)], dtype=np.int32)),
         ("f3", np.array([(0, 1, 2), (3, 4, 5)], dtype=np.int32)),
         ("f4", np.array([(0, 1, 2, 3), (4, 5, 6, 7), (0, 2, 4, 6)], dtype=np.int32))]
sh = ShardedTreeSequence(t, windows, rank, world, device=local)
out, times = {}, {}
for name, idx in calls:
    for rep in range(3):
        barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        r = sh.stat(name, sizes, flat, idx, windows=windows, mode="branch")
        torch.cuda.synchronize(); barrier(); dt = time.perf_counter() - t0
    out[name], times[name] = r, dt
# relatedness vector (GRM x vector), 10 windows, 2 weight columns: the ranks' rows are summed the same way
w10 = np.linspace(0, t.sequence_length, 11)
wt = np.random.default_rng(1).normal(size=(len(s), 2))
for rep in range(3):
    barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    rv = sh.stat("genetic_relatedness_vector", wt, windows=w10, mode="branch", centre=True, nodes=s)
    torch.cuda.synchronize(); barrier(); rv_dt = time.perf_counter() - t0
# matrices and decode: site / branch divergence matrix of 2000 / 64 samples, genotype
# decode of 1024 samples, sharded by genome range / site
sub = s[:: len(s) // 2000][:2000].astype(np.int32)
sub_b = s[:: len(s) // 64][:64].astype(np.int32)
w4 = np.linspace(0, t.sequence_length, 5)
mat = {}
for mode, ss in (("site", sub), ("branch", sub_b)):
    for rep in range(2):
        barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        m = sh.divergence_matrix(w4, sample_sets=ss, sample_set_sizes=np.ones(len(ss), dtype=np.uint64), mode=mode)
        torch.cuda.synchronize(); barrier(); dt = time.perf_counter() - t0
    mat[mode] = (m, dt)
dsub = s[:: len(s) // 1024][:1024].astype(np.int32)
for rep in range(2):
    barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    first, block = sh.genotype_matrix(samples=dsub, gather=False)
    torch.cuda.synchronize(); barrier(); dec_dt = time.perf_counter() - t0
if rank == 0:
    whole = LLTreeSequence(t, device=local)
    res = {"world": world, "ranges": sh.ranges, "workload": bench.workload_name("c2", t, W), "calls": {}}
    for name, idx in calls:
        for rep in range(2):
            t0 = time.perf_counter(); ref = getattr(whole, name)(sizes, flat, idx, windows=windows, mode="branch"); d1 = time.perf_counter() - t0
        err = float(np.max(np.abs(out[name] - ref)) / np.max(np.abs(ref)))
        res["calls"][name] = {"tuples": len(idx), "sharded_ms": times[name] * 1e3, "single_gpu_ms": d1 * 1e3,
                              "max_abs_err_over_max": err}
        assert err < 1e-10, (name, err)
    for rep in range(2):
        t0 = time.perf_counter(); ref = whole.genetic_relatedness_vector(wt, w10, mode="branch", centre=True, nodes=s); d1 = time.perf_counter() - t0
    err = float(np.max(np.abs(rv - ref)) / np.max(np.abs(ref)))
    res["calls"]["genetic_relatedness_vector"] = {"columns": 2, "windows": 10, "sharded_ms": rv_dt * 1e3,
                                                   "single_gpu_ms": d1 * 1e3, "max_abs_err_over_max": err}
    assert err < 1e-10, ("genetic_relatedness_vector", err)
    for mode, ss in (("site", sub), ("branch", sub_b)):
        for rep in range(2):
            t0 = time.perf_counter()
            ref = whole.divergence_matrix(w4, sample_sets=ss, sample_set_sizes=np.ones(len(ss), dtype=np.uint64), mode=mode)
            d1 = time.perf_counter() - t0
        got = mat[mode][0]
        err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        res["calls"]["divergence_matrix_" + mode] = {
            "samples": len(ss), "windows": 4, "sharded_ms": mat[mode][1] * 1e3, "single_gpu_ms": d1 * 1e3,
            "max_abs_err_over_max": err, "bit_identical": bool(np.array_equal(got, ref))}
        assert err < 1e-10, (mode, err)
    for rep in range(2):
        t0 = time.perf_counter(); G = whole.genotype_matrix(samples=dsub); d1 = time.perf_counter() - t0
    res["calls"]["genotype_matrix"] = {
        "samples": len(dsub), "sites": int(t.num_sites), "sharded_ms_rows_stay_sharded": dec_dt * 1e3,
        "single_gpu_ms": d1 * 1e3, "rank0_block_bit_exact": bool(np.array_equal(block, G[first:first + len(block)]))}
    assert res["calls"]["genotype_matrix"]["rank0_block_bit_exact"]
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
