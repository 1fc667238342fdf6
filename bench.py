#!/usr/bin/env python
"""bench.py -- headline measurement of the general-statistics hot path.

Workload (BASELINE.json configs[1], "C2"): windowed branch-mode diversity + divergence on a
seeded synthetic Wright-Fisher ARG of 100k samples, 100 Mb, ~10^7 edges, 1000 windows.
One step = one diversity call (1 sample set) + one divergence call (2 sample sets): two full
sweeps, i.e. 2 x num_edge_diffs edge diffs.  Metric: edge-diffs/s (whole job).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--small]

N > 1 is launched by torchrun, one rank per GPU.  Windows (genome) shard naturally: every rank
owns one 100 Mb chromosome-sized shard of windows (weak scaling) and the per-window results are
gathered with one NCCL all_gather; there is no collective on the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tskit_b200.sim import add_mutations, wright_fisher  # noqa: E402
from tskit_b200.tables import Tables  # noqa: E402

CONFIGS = {
    # name: (samples, generations, L, crossovers per meiosis, mutation draws, windows)
    "c2": (100_000, 2500, 1e8, 2, 1_000_000, 1000),
    "small": (2_000, 500, 1e7, 1, 20_000, 100),
}


def cache_dir():
    d = os.environ.get("TSKB_CACHE", "/tmp/tskb_cache")
    os.makedirs(d, exist_ok=True)
    return d


def load_workload(name, rank=0, barrier=None):
    """Seeded synthetic ARG (ancestry seed 42, mutation seed 1), cached on local disk."""
    n, G, L, nc, muts, W = CONFIGS[name]
    path = os.path.join(cache_dir(), f"wf_{name}_n{n}_g{G}_L{int(L)}_x{nc}_s42.npz")
    gen_s = 0.0
    if rank == 0 and not os.path.exists(path):
        t0 = time.time()
        t = wright_fisher(n, G, L, ncross=nc, seed=42)
        add_mutations(t, muts, seed=1)
        gen_s = time.time() - t0
        t.save(path + ".tmp.npz")
        os.replace(path + ".tmp.npz", path)
    if barrier is not None:
        barrier()
    t = Tables.load(path).ensure_derived()
    return t, W, gen_s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named phase's kernel, from the
    committed ncu --set full capture (profiles/traffic.json, written by tools/ncu_summary.py)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    key = {"sweep": "k_sweep<SVec<int, 1>>", "summary": "k_branch_summary<0, SVec<int, 1>>"}.get(kernel)
    return d.get(key)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ reference arm

def reference_step_fn(t, W):
    """The reference's own C implementation (oracle/_ref) of one sweep: branch diversity."""
    from oracle import ref
    r = ref.RefTreeSequence(t)
    s = t.samples
    windows = np.linspace(0, t.sequence_length, W + 1)

    def one():
        return r.one_way("diversity", [s], windows=windows, mode="branch")
    return one, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtskit_ref.so not built"}))
        return
    name = "small" if args.small else "c2"
    t, W, _ = load_workload(name)
    one, r = reference_step_fn(t, W)
    # edge diffs per sweep, exactly as the reference counts them: edges removed before L + inserted
    nev = int(t.num_edges + np.count_nonzero(t.edges_right < t.sequence_length))
    cores = os.cpu_count() or 1
    pool = ThreadPoolExecutor(cores)

    def step():
        # every host thread runs one full sweep (the reference path is single-threaded and
        # re-entrant on a const tree sequence; ctypes releases the GIL)
        list(pool.map(lambda _: one(), range(cores)))

    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = cores * nev * steps / dt
    unit = "edge-diffs/s"
    print(json.dumps({
        "impl": "reference", "metric": "branch-mode general-stat throughput", "value": value,
        "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(name, t, W),
                   "step": "one branch-diversity sweep per host thread (bounded sample: "
                           "1 of the 2 sweeps of the GPU arm's step)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference",
                         "sample": f"{steps} steps x {cores} concurrent full sweeps of "
                                   "tsk_treeseq_diversity (branch, 1000 windows)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(name, t, W):
    return (f"{name}: branch diversity + divergence, n={t.num_samples}, L={t.sequence_length:.0f}, "
            f"E={t.num_edges}, N={t.num_nodes}, {W} windows")


# ------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from tskit_b200 import _lib
    from tskit_b200.lowlevel import LLTreeSequence, STAT_BRANCH, STAT_SPAN_NORMALISE
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    name = "small" if args.small else "c2"
    t, W, gen_s = load_workload(name, rank, barrier)
    windows = np.linspace(0, t.sequence_length, W + 1)
    s = t.samples
    n = len(s)
    t0 = time.perf_counter()
    ll = LLTreeSequence(t, device=local)
    stage_s = time.perf_counter() - t0
    st = ll.engine_stats()
    nev, visits, levels = st["num_events"], st["num_visits"], st["num_levels"]
    dbar = visits / max(nev, 1)

    sizes1 = np.array([n], dtype=np.uint64)
    sizes2 = np.array([n // 2, n - n // 2], dtype=np.uint64)
    idx = np.array([[0, 1]], dtype=np.int32)
    options = STAT_BRANCH | STAT_SPAN_NORMALISE
    L = _lib.lib()

    # inputs resident in HBM for the kernel-only number (torch owns the buffers)
    d_sets = torch.from_numpy(s.copy()).to(f"cuda:{local}")
    d_res1 = torch.empty(W * 1, dtype=torch.float64, device=f"cuda:{local}")
    d_res2 = torch.empty(W * 1, dtype=torch.float64, device=f"cuda:{local}")
    torch.cuda.synchronize()

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    phase_ms = np.zeros(6)
    launches = [0]

    def step_device():
        ms = 0.0
        for stat_id, sizes, ntup, tup, out in ((0, sizes1, 0, None, d_res1), (3, sizes2, 1, idx, d_res2)):
            ret = L.tskb_treeseq_stat_device(ll._h, stat_id, len(sizes), p(sizes),
                                             C.c_void_p(d_sets.data_ptr()), ntup,
                                             None if tup is None else p(tup), W, p(windows),
                                             options, C.c_void_p(out.data_ptr()))
            if ret != 0:
                raise RuntimeError(L.tskb_strerror(ret).decode() + L.tskb_last_cuda_error().decode())
            es = ll.engine_stats()
            ms += es["last_call_ms"]
            phase_ms[:] += np.array(es["last_kernel_ms"][:6])
            launches[0] += es["last_launches"]
        return ms

    # pinned host buffers for the end-to-end number through the C ABI
    h_sets = torch.from_numpy(s.copy()).pin_memory()
    h_res1 = torch.empty((W, 1), dtype=torch.float64).pin_memory()
    h_res2 = torch.empty((W, 1), dtype=torch.float64).pin_memory()
    h2d = int(h_sets.numel() * 4 * 2 + (W + 1) * 8 * 2 + idx.nbytes + 16 + 8)
    d2h = int(W * 8 * 2)

    def step_e2e():
        for fn, sizes, tup, out in ((L.tskb_treeseq_diversity, sizes1, None, h_res1),
                                    (L.tskb_treeseq_divergence, sizes2, idx, h_res2)):
            a = [ll._h, len(sizes), p(sizes), C.c_void_p(h_sets.data_ptr())]
            if tup is not None:
                a += [1, p(tup)]
            a += [W, p(windows), options, C.c_void_p(out.data_ptr())]
            ret = fn(*a)
            if ret != 0:
                raise RuntimeError(L.tskb_strerror(ret).decode())

    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()
    phase_ms[:] = 0
    launches[0] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_ms += step_device()
    torch.cuda.synchronize()
    wall_dev = time.perf_counter() - t0
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # per-window results of every shard -> all ranks (the only collective on this path)
    gathered = None
    if world > 1:
        out = [torch.empty_like(d_res1) for _ in range(world)]
        dist.all_gather(out, d_res1)
        gathered = torch.stack(out)
        times = torch.tensor([dev_ms, wall_e2e, wall_dev], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dev_ms, wall_e2e, wall_dev = [float(x) for x in times.cpu()]

    if rank == 0:
        sweeps = 2
        diffs_per_step = sweeps * nev * world
        value = diffs_per_step * args.steps / (dev_ms / 1e3)
        e2e_value = diffs_per_step * args.steps / wall_e2e
        peak, peak_src = peaks()
        K_avg = 1.5  # one sweep with K = 1 state column and one with K = 2
        b_branch = 28 + dbar * (12 + 8 * K_avg)
        per_step_ms = phase_ms / args.steps
        names = ["weights", "sweep", "summary", "integrate", "idle", "d2h"]
        dom = int(np.argmax(per_step_ms))
        # algorithmic share of the dominant phase ("Roofline accounting")
        share = {"sweep": 20 + dbar * (4 + 8 * K_avg), "summary": 8 + dbar * 8}.get(
            names[dom], b_branch)
        achieved = share * sweeps * nev / (per_step_ms[dom] / 1e3) / 1e9
        sweep_achieved = b_branch * sweeps * nev / (dev_ms / args.steps / 1e3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(t, W, nev)
        line = {
            "metric": "branch-mode general-stat throughput (edge-diffs/s)", "value": value,
            "unit": "edge-diffs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(name, t, W),
                "step": "diversity(branch, 1 set) + divergence(branch, 2 sets): 2 sweeps",
                "edge_diffs_per_sweep": nev, "visits_per_sweep": visits, "d_bar": dbar,
                "levels": levels, "l2": "inputs larger than L2 (plan arrays %.2f GB)" %
                                        (st["device_bytes"] / 1e9),
                "sharding": "one 100 Mb window shard per rank (same synthetic ARG per rank), "
                            "per-window results all_gathered over NCCL",
                "stage_s": stage_s, "generate_s": gen_s,
                "phase_ms_per_step": dict(zip(names, [float(x) for x in per_step_ms])),
                "wall_ms_per_step_device_inputs": wall_dev / args.steps * 1e3,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "edge-diffs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": wall_e2e / args.steps * 1e3,
                    "cold_including_staging_value": diffs_per_step /
                    (world * stage_s + wall_e2e / args.steps)},
            "gpu_launches": int(launches[0]),
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(names[dom]),
                         "traffic_note": "bytes per launch of the K=1 instantiation, ncu capture under profiles/",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_edge_diff": share,
                         "whole_sweep": {"achieved": sweep_achieved, "frac": sweep_achieved / peak,
                                         "algorithmic_bytes_per_edge_diff": b_branch}},
            "cpu_baseline": cpu,
        }
        if gathered is not None:
            line["config"]["gathered_shape"] = list(gathered.shape)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(t, W, nev):
    """The reference C implementation (oracle/_ref) on this box's host: one core, the full
    workload step (2 sweeps), best of 2."""
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "edge-diffs/s", "cores": 1, "kind": "reference",
                "sample": "oracle/_ref not built"}
    r = ref.RefTreeSequence(t)
    s = t.samples
    n = len(s)
    windows = np.linspace(0, t.sequence_length, W + 1)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        r.one_way("diversity", [s], windows=windows, mode="branch")
        r.k_way("divergence", [s[: n // 2], s[n // 2:]], [[0, 1]], windows=windows, mode="branch")
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": 2 * nev / best, "unit": "edge-diffs/s", "cores": 1, "kind": "reference",
            "sample": "full step (tsk_treeseq_diversity + tsk_treeseq_divergence, branch, "
                      f"{W} windows), best of 2, {best:.2f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny workload (smoke test of the bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
