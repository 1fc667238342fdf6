#!/usr/bin/env python
"""bench.py -- headline measurement of the general-statistics hot path.

Workload (BASELINE.json configs[1], "C2"): windowed branch-mode diversity + divergence on a
seeded synthetic Wright-Fisher ARG of 100k samples, 100 Mb, ~10^7 edges, 1000 windows.
One step = one diversity call (1 sample set) + one divergence call (2 sample sets): two full
sweeps, i.e. 2 x num_edge_diffs edge diffs.  Metric: edge-diffs/s (whole job).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--small]
                  [--config c2|c3]

N > 1 (launched by torchrun, one rank per GPU) is weak scaling over GENOME SHARDS of one tree
sequence: the genome is N x 100 Mb (the C2 ARG repeated along the genome, tskit_b200.sim.
repeat_genome: N unlinked chromosomes over the same 100k samples, N x 10^7 edges, N x 1000
windows), tskit_b200.sharding.plan_shards cuts it into N ranges of equal edge-diff count, every
rank stages only its range (tskb_treeseq_init(range_left, range_right) on the rows meeting it)
and computes un-normalised per-window partials; ONE all_reduce (NCCL) of the device-resident
W x M partials per statistic, INSIDE the timed region, gives every rank the whole result.
--config c3 runs BASELINE.json configs[2] (f2/f3/f4 + Fst over 8 sample sets, 10^6 samples,
1 Gb) as a strong-scaling study over the same sharding (see run_c3).
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

from tskit_b200.sim import add_mutations, repeat_genome, wright_fisher  # noqa: E402
from tskit_b200.tables import Tables  # noqa: E402

METRIC = "edge-diffs/s (branch-mode general stat: diversity + 2-set divergence, windowed)"
UNIT = "edge-diffs/s"
RTOL = 1e-9  # north_star: fp64 statistics within relative 1e-9 of the C reference

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

    def __init__(self, index, interval_ms=50):
        self.index = index
        self.interval_ms = interval_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", str(self.interval_ms), "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.3)  # first samples arrive before the timed region starts
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
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named phase's kernel, from the
    committed ncu --set full capture (profiles/traffic.json, written by tools/ncu_summary.py)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    for key in {"sweep": ("k_sweep<SVec<int, 1>>",),
                "summary": ("k_branch_summary_runs<0, SVec<int, 1>>", "k_branch_summary<0, SVec<int, 1>>")
                }.get(kernel, ()):
        if key in d:
            return d[key]
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(name, t, W, copies=1):
    rep = "" if copies == 1 else f" x {copies} along the genome (one tree sequence, genome-sharded)"
    return (f"{name}: branch diversity + divergence, n={t.num_samples}, L={t.sequence_length:.0f}, "
            f"E={t.num_edges}, N={t.num_nodes}, {W} windows{rep}")


def rel_err(got, want):
    """Largest |got - want| / |want| over the entries (0 where both are 0; inf where only one is
    finite or they differ in NaN-ness)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if got.shape != want.shape:
        return float("inf")
    bad = np.isnan(got) != np.isnan(want)
    if bad.any():
        return float("inf")
    m = ~np.isnan(want)
    d = np.abs(got[m] - want[m])
    den = np.abs(want[m])
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(d == 0, 0.0, d / den)
    return float(r.max()) if r.size else 0.0


# ------------------------------------------------------------------ the reference (checker / CPU arm)

def reference_step(r, t, W):
    """One step of the workload on the reference's own C implementation (oracle/_ref):
    tsk_treeseq_diversity (trees.c:3950) + tsk_treeseq_divergence (trees.c:4711), branch mode."""
    s = t.samples
    n = len(s)
    windows = np.linspace(0, t.sequence_length, W + 1)
    a = r.one_way("diversity", [s], windows=windows, mode="branch")
    b = r.k_way("divergence", [s[: n // 2], s[n // 2:]], [[0, 1]], windows=windows, mode="branch")
    return a, b


def edge_diffs_per_sweep(t):
    """Edge diffs of one sweep exactly as the reference replays them (trees.c:1424-1507): every edge
    inserted, every edge that ends before L removed."""
    return int(t.num_edges + np.count_nonzero(t.edges_right < t.sequence_length))


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
    r = ref.RefTreeSequence(t)
    nev = edge_diffs_per_sweep(t)
    cores = os.cpu_count() or 1
    pool = ThreadPoolExecutor(cores)

    def step():
        # every host thread runs one full step (the reference path is single-threaded and re-entrant on
        # a const tree sequence; ctypes releases the GIL): the same two calls as the GPU arm's step
        list(pool.map(lambda _: reference_step(r, t, W), range(cores)))

    budget_s = 240.0  # the whole run stays within a few minutes
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    warmup = max(0, min(args.warmup, int(budget_s * 0.2 / first)))
    steps = max(1, min(args.steps, int(budget_s * 0.8 / first)))
    for _ in range(max(0, warmup - 1)):  # the probe step above was the first warm-up step
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = cores * 2 * nev * steps / dt
    note = None
    if steps != args.steps or warmup != args.warmup:
        note = (f"one step = {first:.1f} s on {cores} threads: --steps {args.steps} --warmup {args.warmup} "
                f"bounded to {steps} / {max(warmup, 1)} to stay within {budget_s:.0f} s")
    copies = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(warmup, 1), "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(name, t, W),
                   "step": "diversity(branch, 1 set) + divergence(branch, 2 sets): 2 sweeps, one "
                           "full step per host thread, all threads concurrently",
                   "edge_diffs_per_sweep": nev},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{steps} steps x {cores} concurrent full steps of "
                                   f"tsk_treeseq_diversity + tsk_treeseq_divergence (branch, {W} windows) "
                                   f"on the {name} ARG"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if copies > 1:
        line["config"]["sample_note"] = (
            f"the GPU arm at {copies} GPUs runs this ARG repeated {copies} x along the genome; the "
            "reference's rate per edge diff does not depend on the genome length, so the bounded "
            "sample is one copy")
    if note:
        line["config"]["steps_note"] = note
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm

def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from tskit_b200 import _lib, sharding
    from tskit_b200.lowlevel import STAT_BRANCH, STAT_SPAN_NORMALISE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    torch.empty(1, device=dev)
    torch.cuda.synchronize()                # the CUDA context exists before the staging is timed
    name = "small" if args.small else "c2"
    base, W1, gen_s = load_workload(name, rank, barrier)
    t0 = time.perf_counter()
    t = repeat_genome(base, world)          # world == 1: the C2 ARG itself
    W = W1 * world
    windows = np.linspace(0, t.sequence_length, W + 1)
    tile_s = time.perf_counter() - t0
    s = t.samples
    n = len(s)
    t0 = time.perf_counter()
    sh = sharding.ShardedTreeSequence(t, windows, rank, world, device=local)
    stage_s = time.perf_counter() - t0      # plan_shards + restrict_tables + tskb_treeseq_init(range)
    ll = sh.engine
    st = ll.engine_stats()
    init_s = st["stage_ms"] / 1e3
    nev_local, visits_local, levels = st["num_events"], st["num_visits"], st["num_levels"]
    nev_base = edge_diffs_per_sweep(base)
    nev_total = edge_diffs_per_sweep(t)     # what the reference would replay over the whole genome
    dbar = visits_local / max(nev_local, 1)

    sizes1 = np.array([n], dtype=np.uint64)
    sizes2 = np.array([n // 2, n - n // 2], dtype=np.uint64)
    idx = np.array([[0, 1]], dtype=np.int32)
    options = STAT_BRANCH | STAT_SPAN_NORMALISE
    L = _lib.lib()

    # inputs resident in HBM for the kernel-only number (torch owns the buffers)
    d_sets = torch.from_numpy(s.copy()).to(dev)
    # the step's two (W x 1) results; two buffers, alternating: with several GPUs the all_reduce of
    # one step is still in flight on torch's stream while the next step's sweeps write their results
    d_bufs = [torch.empty((2, W), dtype=torch.float64, device=dev) for _ in range(2)]
    pending = [None, None]   # event after the last collective that used each buffer
    step_no = [0]
    torch.cuda.synchronize()

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    phase_ms = np.zeros(6)
    launches = [0]
    coll_ev = []
    # TSKB_BENCH_SYNC_COLL=1: the host waits for each step's collective before the next step's sweeps
    # (no overlap: the NCCL kernel then never shares the SMs with the cooperative sweep)
    sync_coll = os.environ.get("TSKB_BENCH_SYNC_COLL", "0") == "1"
    # The sum of the ranks' partials: pushed over NVLink peer memory by the engine's own kernels
    # (sharding.PeerExchange, default) or, with TSKB_BENCH_NCCL=1, one NCCL all_reduce per step.
    exchange = None
    coll_wall = [0.0]
    if world > 1 and os.environ.get("TSKB_BENCH_NCCL", "0") != "1":
        exchange = sh.use_peer_exchange(2 * W)

    def step_device(collect=True):
        """One step with inputs and outputs in HBM; returns the engine's device time (CUDA events on
        the engine's stream).  Several GPUs: both statistics' un-normalised partials go through ONE
        all_reduce (NCCL, on torch's stream: it overlaps the next step's first sweep on the engine's
        stream) followed by the span normalisation; coll_ev brackets it."""
        ms = 0.0
        b = step_no[0] & 1
        step_no[0] += 1
        d_both = d_bufs[b]
        if pending[b] is not None:
            pending[b].synchronize()   # the collective of two steps ago has released this buffer
        for nm, sizes, tup, out in (("diversity", sizes1, None, d_both[0]), ("divergence", sizes2, idx, d_both[1])):
            ll.stat_device(nm, sizes, d_sets.data_ptr(), tup, windows, options if world == 1 else STAT_BRANCH,
                           out.data_ptr())
            if collect:
                es = ll.engine_stats()
                ms += es["last_call_ms"]
                phase_ms[:] += np.array(es["last_kernel_ms"][:6])
                launches[0] += es["last_launches"]
        if world > 1 and exchange is not None:
            # on the engine's stream, behind the two statistics; the host waits for it only in the
            # instrumented steps (collect), where its wall time is the exchange incl. the slowest rank
            t_c = time.perf_counter()
            exchange.sum_into(d_both, d_both, windows, window_axis=1, wait=collect)
            if collect:
                coll_wall[0] += (time.perf_counter() - t_c) * 1e3
        elif world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sharding.combine(d_both, windows, True, window_axis=1)
            e1.record()
            coll_ev.append((e0, e1))
            pending[b] = e1
            if sync_coll:
                e1.synchronize()
        return ms

    # host buffers for the end-to-end number: the reference-facing C ABI at one GPU, the sharded host
    # API (pinned staging, all_reduce on the device, one read-back) at several
    h_sets = torch.from_numpy(s.copy()).pin_memory()
    h_res1 = torch.empty((W, 1), dtype=torch.float64).pin_memory()
    h_res2 = torch.empty((W, 1), dtype=torch.float64).pin_memory()
    h2d = int(h_sets.numel() * 4 * 2 + (W + 1) * 8 * 2 + idx.nbytes + 16 + 8)
    d2h = int(W * 8 * 2)
    e2e_out = {}

    def step_e2e():
        if world == 1:
            for fn, sizes, tup, out in ((L.tskb_treeseq_diversity, sizes1, None, h_res1),
                                        (L.tskb_treeseq_divergence, sizes2, idx, h_res2)):
                a = [ll._h, len(sizes), p(sizes), C.c_void_p(h_sets.data_ptr())]
                if tup is not None:
                    a += [1, p(tup)]
                a += [W, p(windows), options, C.c_void_p(out.data_ptr())]
                ret = fn(*a)
                if ret != 0:
                    raise RuntimeError(L.tskb_strerror(ret).decode())
            e2e_out["res"] = (h_res1.numpy(), h_res2.numpy())
        else:
            a = sh.stat_host("diversity", sizes1, h_sets.numpy(), None, windows, options)
            b = sh.stat_host("divergence", sizes2, h_sets.numpy(), idx, windows, options)
            e2e_out["res"] = (a, b)

    def collective_ms():
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in coll_ev) + coll_wall[0]
        coll_ev.clear()
        coll_wall[0] = 0.0
        return ms

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
        step_e2e()
    collective_ms()
    # calibration: the timed region lasts at least ~1 s (blocks of exactly --steps steps)
    t0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    one = max(time.perf_counter() - t0, 1e-4)
    collective_ms()
    blocks = max(1, int(np.ceil(1.0 / (one * args.steps))))
    if world > 1:
        b_t = torch.tensor([blocks], device=dev)
        dist.all_reduce(b_t, op=dist.ReduceOp.MAX)
        blocks = int(b_t.item())
    blocks = min(blocks, 500)
    if os.environ.get("TSKB_BENCH_BLOCKS"):  # profiler runs: a short timed region
        blocks = max(1, int(os.environ["TSKB_BENCH_BLOCKS"]))
    phase_ms[:] = 0
    launches[0] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    dev_ms = 0.0
    span0, span1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    span0.record()
    for _ in range(blocks):
        for _ in range(args.steps):
            dev_ms += step_device(collect=world == 1)
    if exchange is not None:
        exchange.status()   # the last exchange has completed (and none timed out) before the closing event
    span1.record()          # follows the last sum + normalisation; the engine calls are synchronous
    torch.cuda.synchronize()
    wall_dev = time.perf_counter() - t0
    coll_ms = collective_ms()
    if world > 1:
        # Engine work (its own stream) and collectives (torch's stream) overlap and cannot be added up:
        # the device time of the timed region is the span between two events that bracket it.  The
        # per-phase engine times come from a few extra steps outside the timed region.
        dev_ms = span0.elapsed_time(span1)
        for _ in range(5):
            step_device(collect=True)
        phase_ms *= nsteps_scale(blocks * args.steps, 5)
        launches[0] = int(launches[0] * nsteps_scale(blocks * args.steps, 5))
        c5 = collective_ms()
        if exchange is not None:
            coll_ms = c5 * nsteps_scale(blocks * args.steps, 5)   # host-waited exchanges of the 5 extra steps
    barrier()
    torch.cuda.synchronize()
    e2e_blocks = max(1, blocks // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_blocks * args.steps):
        step_e2e()
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    nsteps = blocks * args.steps
    nsteps_e2e = e2e_blocks * args.steps

    if world > 1:
        times = torch.tensor([dev_ms, wall_e2e, wall_dev, coll_ms, stage_s, init_s], dtype=torch.float64,
                             device=dev)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dev_ms, wall_e2e, wall_dev, coll_ms, stage_s, init_s = [float(x) for x in times.cpu()]
        counts = torch.tensor([nev_local, visits_local], dtype=torch.float64, device=dev)
        gathered = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(gathered, counts)
        nev_ranks = [int(g[0].item()) for g in gathered]
    else:
        nev_ranks = [nev_local]

    if rank == 0:
        sweeps = 2
        diffs_per_step = sweeps * nev_total
        value = diffs_per_step * nsteps / (dev_ms / 1e3)
        e2e_value = diffs_per_step * nsteps_e2e / wall_e2e
        peak, peak_src = peaks()
        K_avg = 1.5  # one sweep with K = 1 state column and one with K = 2
        b_branch = 28 + dbar * (12 + 8 * K_avg)
        per_step_ms = phase_ms / nsteps
        names = ["weights", "sweep", "summary", "integrate", "idle", "d2h"]
        dom = int(np.argmax(per_step_ms))
        # algorithmic share of the dominant phase ("Roofline accounting"): per launch =
        # share x the edge diffs one launch covers (this rank's range), / its mean duration
        share = {"sweep": 20 + dbar * (4 + 8 * K_avg), "summary": 8 + dbar * 8}.get(names[dom], b_branch)
        achieved = share * sweeps * nev_local / (per_step_ms[dom] / 1e3) / 1e9
        step_achieved = b_branch * diffs_per_step / world / (dev_ms / nsteps / 1e3) / 1e9

        # parity of THIS run's result against the reference C library on the same tables
        got1, got2 = e2e_out["res"]
        parity, cpu = None, None
        if not args.no_cpu_baseline:
            cpu, want = cpu_baseline(base, W1, nev_base, best_of=2 if world == 1 else 1)
            if want is not None:
                # every copy of the genome repeats the base ARG: each block of W1 windows is the base result
                d_last = d_bufs[(step_no[0] - 1) & 1]   # the last device-resident step's results
                d_res1, d_res2 = d_last[0], d_last[1]
                e1 = max(rel_err(got1.reshape(world, W1), np.broadcast_to(want[0].reshape(1, W1), (world, W1))),
                         rel_err(d_res1.cpu().numpy().reshape(world, W1),
                                 np.broadcast_to(want[0].reshape(1, W1), (world, W1))))
                e2 = max(rel_err(got2.reshape(world, W1), np.broadcast_to(want[1].reshape(1, W1), (world, W1))),
                         rel_err(d_res2.cpu().numpy().reshape(world, W1),
                                 np.broadcast_to(want[1].reshape(1, W1), (world, W1))))
                parity = {"max_rel_err": max(e1, e2), "rtol": RTOL, "ok": bool(max(e1, e2) <= RTOL),
                          "against": "oracle/_ref (reference C library) on the same tables: "
                                     "tsk_treeseq_diversity + tsk_treeseq_divergence, every window, "
                                     "device-resident and host-buffer results",
                          "windows_checked": int(2 * 2 * W)}
            if world > 1:
                cpu = None  # the CPU baseline is reported at one GPU only
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": dev_ms / nsteps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(name, base, W1, world),
                "step": "diversity(branch, 1 set) + divergence(branch, 2 sets): 2 sweeps",
                "edge_diffs_per_sweep": nev_total, "edge_diffs_per_sweep_by_rank": nev_ranks,
                "visits_per_sweep_rank0": visits_local, "d_bar": dbar, "levels": levels,
                "l2": "inputs larger than L2 (plan arrays %.2f GB per rank)" % (st["device_bytes"] / 1e9),
                "sharding": ("whole genome on one GPU" if world == 1 else
                             f"genome ranges from sharding.plan_shards, one per rank; per statistic one "
                             f"sum of both statistics' {W} x 1 device-resident partials inside the timed "
                             f"region ({coll_ms / nsteps:.3f} ms per step incl. waiting for the slowest rank), "
                             + ("pushed over NVLink peer memory by the engine's kernels (tskb_exchange_sum), "
                                "summed in rank order on every rank" if exchange is not None else
                                "NCCL all_reduce on torch's stream") +
                             "; value = edge diffs / device span of the timed region (CUDA events around it)"),
                "timed_blocks": blocks, "timed_steps": nsteps, "timed_seconds_device": dev_ms / 1e3,
                "stage_s": stage_s, "init_s": init_s, "tile_s": tile_s, "generate_s": gen_s,
                "phase_ms_per_step": dict(zip(names, [float(x) for x in per_step_ms])),
                "collective_ms_per_step": coll_ms / nsteps,
                "wall_ms_per_step_device_inputs": wall_dev / nsteps * 1e3,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": wall_e2e / nsteps_e2e * 1e3,
                    "steps": nsteps_e2e,
                    "through": ("tskb_treeseq_diversity / tskb_treeseq_divergence (C ABI, host pointers)"
                                if world == 1 else "sharding.ShardedTreeSequence.stat_host (pinned "
                                "staging, all_reduce on the device, one read-back)"),
                    "cold_including_staging_value": diffs_per_step / (stage_s + wall_e2e / nsteps_e2e)},
            "gpu_launches": int(launches[0]) + (nsteps if exchange is not None else 0),
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(names[dom]),
                         "traffic_note": "bytes per launch of the K=1 instantiation, ncu capture under profiles/",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_edge_diff": share,
                         "whole_step": {"achieved": step_achieved, "frac": step_achieved / peak,
                                        "algorithmic_bytes_per_edge_diff": b_branch}},
            "parity": parity,
            "cpu_baseline": cpu,
        }
        if world == 1 and not args.no_secondary:
            try:
                line["site_and_matrix"] = secondary(ll, base, W1, args)
            except Exception as e:  # the headline line is still printed
                line["site_and_matrix"] = {"error": repr(e)}
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            if world > 1:
                dist.destroy_process_group()
            raise SystemExit(f"bench.py: parity against the reference failed: {parity}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def nsteps_scale(timed_steps, sampled_steps):
    """Per-phase engine times are sampled over a few steps and reported per timed step."""
    return timed_steps / float(sampled_steps)


def cpu_baseline(t, W, nev, best_of=2):
    """The reference C implementation (oracle/_ref) on this box's host: one core, the full
    workload step (2 sweeps).  Also returns its results: the in-run parity check."""
    from oracle import ref
    if not ref.available():
        return ({"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                 "sample": "oracle/_ref not built"}, None)
    r = ref.RefTreeSequence(t)
    best, res = None, None
    for _ in range(best_of):
        t0 = time.perf_counter()
        res = reference_step(r, t, W)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ({"value": 2 * nev / best, "unit": UNIT, "cores": 1, "kind": "reference",
             "sample": "full step (tsk_treeseq_diversity + tsk_treeseq_divergence, branch, "
                       f"{W} windows), best of {best_of}, {best:.2f} s"}, res)


# ------------------------------------------------------------------ site mode, decode, matrix (sample.sites/s)

def int8_peak():
    """Dense int8 tensor-core throughput of this GPU through cuBLASLt (torch._int_mm, 8192^3),
    best of 10 and sustained over ~2 s: the denominator of the matrix path's roofline."""
    import torch
    n = 8192
    a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device="cuda")
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = None
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int(2000 / best))
    e0.record()
    for _ in range(reps):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    ops = 2.0 * n ** 3
    return {"burst_tops": ops / (best / 1e3) / 1e12, "sustained_tops": ops * reps / (e0.elapsed_time(e1) / 1e3) / 1e12,
            "how": "torch._int_mm (cuBLASLt int8 -> int32) 8192^3, best of 10 / back to back for ~2 s"}


def secondary(ll, t, W, args):
    """The sample.sites/s half of BASELINE.json's metric on the same ARG, one B200: site-mode
    statistics (virtual rate: genotypes are never formed), genotype decode, and the site-mode
    divergence matrix (tcgen05 int8) against a measured int8 peak."""
    import torch
    from oracle import ref
    out = {}
    s = t.samples
    n, S = len(s), t.num_sites
    windows = np.linspace(0, t.sequence_length, W + 1)
    peak, _ = peaks()
    # ---- site-mode diversity + divergence through the mirror of the reference's low-level interface
    sizes2 = np.array([n // 2, n - n // 2], dtype=np.uint64)
    best, phases = None, None
    for _ in range(5):
        t0 = time.perf_counter()
        a = ll.diversity(np.array([n], dtype=np.uint64), s, windows=windows, mode="site")
        p1 = ll.engine_stats()["last_kernel_ms"][:6]
        b = ll.divergence(sizes2, s, [[0, 1]], windows=windows, mode="site")
        p2 = ll.engine_stats()["last_kernel_ms"][:6]
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, phases = dt, [float(x + y) for x, y in zip(p1, p2)]
    st = ll.engine_stats()
    nev, dbar = st["num_events"], st["num_visits"] / max(st["num_events"], 1)
    dev_s = sum(phases) / 1e3
    b_site = 20 + dbar * (4 + 8 * 1.5)
    site = {"calls": "diversity(site, 1 set) + divergence(site, 2 sets), %d windows, host buffers" % W,
            "sites": S, "samples": n, "seconds_e2e": best, "seconds_device": dev_s,
            "sample_sites_per_s_e2e": 2.0 * n * S / best, "sample_sites_per_s_device": 2.0 * n * S / dev_s,
            "edge_diffs_per_s_device": 2.0 * nev / dev_s,
            "roofline": {"bound": "hbm", "achieved": b_site * 2 * nev / dev_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": b_site * 2 * nev / dev_s / 1e9 / peak,
                         "algorithmic_bytes_per_edge_diff": b_site},
            "phase_ms": dict(zip(["weights", "sweep", "site_summary", "window_sums", "idle", "d2h"], phases))}
    if ref.available() and not args.no_cpu_baseline:
        r = ref.RefTreeSequence(t)
        t0 = time.perf_counter()
        wa = r.one_way("diversity", [s], windows=windows, mode="site")
        wb = r.k_way("divergence", [s[: n // 2], s[n // 2:]], [[0, 1]], windows=windows, mode="site")
        cpu_s = time.perf_counter() - t0
        e = max(rel_err(a, wa), rel_err(b, wb))
        site["parity"] = {"max_rel_err": e, "rtol": RTOL, "ok": bool(e <= RTOL)}
        site["cpu_reference_1core"] = {"seconds": cpu_s, "sample_sites_per_s": 2.0 * n * S / cpu_s}
    out["site_stats"] = site
    # ---- genotype decode (genotypes.c:473-594): int8 [sites x samples] for a sample subset
    sub = s[:: max(1, n // 2048)][:2048].copy()
    g = None
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        g = ll.genotype_matrix(samples=sub)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    dec = {"samples": len(sub), "sites": S, "seconds_e2e": best,
           "sample_sites_per_s_e2e": len(sub) * S / best,
           "note": "includes the read-back of the %.1f GB int8 matrix into pageable host memory" % (g.nbytes / 1e9)}
    if ref.available() and not args.no_cpu_baseline:
        # bit-exact against tsk_variant_decode over every site, for a 512-sample subset
        chk = s[:: max(1, n // 512)][:512].copy()
        r = ref.RefTreeSequence(t)
        t0 = time.perf_counter()
        want = r.genotype_matrix(samples=chk)
        cpu_s = time.perf_counter() - t0
        got = ll.genotype_matrix(samples=chk)
        dec["parity"] = {"bit_exact": bool(np.array_equal(got, want)), "samples": len(chk), "sites": S}
        dec["cpu_reference_1core"] = {"seconds": cpu_s, "sample_sites_per_s": len(chk) * S / cpu_s}
        del want, got
    out["decode"] = dec
    del g
    # ---- site-mode divergence matrix of a sample subset: decode + tcgen05 int8 Gram + fp64 epilogue
    nm = 8192 if n >= 8192 else n
    subm = s[:: max(1, n // nm)][:nm].copy()
    sets = subm.astype(np.int32)
    sizes = np.ones(len(subm), dtype=np.uint64)
    w1 = np.array([0.0, t.sequence_length])
    best, ph = None, None
    D = None
    for _ in range(3):
        t0 = time.perf_counter()
        D = ll.divergence_matrix(w1, sample_sets=sets, sample_set_sizes=sizes, mode="site", span_normalise=False)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, ph = dt, ll.matrix_phase_ms()
    nm = len(subm)
    ops = float(nm) * (nm + 1) * S
    pk = int8_peak()
    gemm_s = ph["gemm"] / 1e3
    D = D[0]
    mat = {"samples": nm, "sites": S, "alleles": ph["alleles"], "seconds_e2e": best, "phase_ms": ph,
           "sample_sites_per_s_e2e": nm * S / best,
           "useful_int_ops": ops, "contraction_tops": ops / gemm_s / 1e12,
           "int8_peak": pk,
           "roofline": {"bound": "tensor", "achieved": ops / gemm_s / 1e12, "peak": pk["sustained_tops"],
                        "unit": "Tint8op/s", "frac": ops / gemm_s / 1e12 / pk["sustained_tops"],
                        "note": "useful ops n(n+1)S of the upper triangle over the contraction kernel's time"},
           "properties": {"symmetric": bool(np.array_equal(D, D.T)), "zero_diagonal": bool((np.diag(D) == 0).all()),
                          "integers": bool(np.array_equal(D, np.round(D)))}}
    out["site_divergence_matrix"] = mat
    torch.cuda.synchronize()
    return out


# ------------------------------------------------------------------ C3: f2/f3/f4 + Fst, 8 sets, 10^6 samples, 1 Gb

C3 = dict(samples=1_000_000, generations=40, arms=8, arm_length=1.25e8, ncross=1, windows=10_000)
C3_SMALL = dict(samples=4_000, generations=40, arms=8, arm_length=1.25e6, ncross=1, windows=400)


def c3_workload(cfg, rank, world, barrier):
    """BASELINE.json configs[2]: one tree sequence of `arms` unlinked chromosome arms over the same 10^6
    samples (seeded Wright-Fisher, ancestry seed 42 + arm; 40 generations back: every local tree is a
    forest of ~n / 21 subtrees, which is a valid input -- a fully coalesced 10^6-sample, 1 Gb ARG has
    ~10^9+ edges and fits neither the generator's time budget nor one GPU).  The ranks simulate the arms
    between them, exchange them through the local disk and every rank assembles the whole tables."""
    from concurrent.futures import ThreadPoolExecutor
    from tskit_b200.sim import concat_genomes
    tag = "c3_n%d_g%d_a%d_L%d" % (cfg["samples"], cfg["generations"], cfg["arms"], int(cfg["arm_length"]))

    def path(a):
        return os.path.join(cache_dir(), f"{tag}_arm{a}.npz")

    def make(a):
        if not os.path.exists(path(a)):
            t = wright_fisher(cfg["samples"], cfg["generations"], cfg["arm_length"], ncross=cfg["ncross"],
                              seed=42 + a, num_threads=1)
            t.save(path(a) + ".tmp.npz")
            os.replace(path(a) + ".tmp.npz", path(a))

    t0 = time.time()
    mine = [a for a in range(cfg["arms"]) if a % world == rank]
    threads = max(1, (os.cpu_count() or 1) // world)
    with ThreadPoolExecutor(min(threads, max(1, len(mine)))) as pool:
        list(pool.map(make, mine))
    gen_s = time.time() - t0
    barrier()
    t0 = time.time()
    t = concat_genomes([Tables.load(path(a)) for a in range(cfg["arms"])])
    return t, gen_s, time.time() - t0


def run_c3(args):
    """Strong scaling of BASELINE.json configs[2] over genome shards: one step = f2 (28 pairs) + f3 (8
    triples) + f4 (2 quadruples) + Fst (= diversity of the 8 sets + divergence of the 28 pairs,
    trees.py:10076-10126), branch mode, 10^4 windows: 5 sweeps.  Every rank stages its genome range
    and computes un-normalised partials; per statistic one all_reduce of the device-resident W x M
    partials inside the timed region."""
    import torch
    import torch.distributed as dist

    from tskit_b200 import sharding
    from tskit_b200.lowlevel import STAT_BRANCH, STAT_SPAN_NORMALISE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    cfg = C3_SMALL if args.small else C3
    t, gen_s, concat_s = c3_workload(cfg, rank, world, barrier)
    W = cfg["windows"]
    windows = np.linspace(0, t.sequence_length, W + 1)
    s = t.samples
    sets = np.array_split(s, 8)
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    pairs = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], dtype=np.int32)
    triples = np.array([(i, (i + 1) % 8, (i + 2) % 8) for i in range(8)], dtype=np.int32)
    quads = np.array([(0, 1, 2, 3), (4, 5, 6, 7)], dtype=np.int32)
    calls = [("f2", pairs), ("f3", triples), ("f4", quads), ("diversity", None), ("divergence", pairs)]
    t0 = time.perf_counter()
    sh = sharding.ShardedTreeSequence(t, windows, rank, world, device=local)
    stage_s = time.perf_counter() - t0
    ll = sh.engine
    st = ll.engine_stats()
    nev_total = edge_diffs_per_sweep(t)
    d_sets = torch.from_numpy(s.copy()).to(dev)
    outs = {nm: torch.empty((W, len(sizes) if ix is None else len(ix)), dtype=torch.float64, device=dev)
            for nm, ix in calls}
    phase = {}
    worst = {}
    kphase = {}
    coll_ev = []
    coll_wall = [0.0]
    exchange = None
    if world > 1 and os.environ.get("TSKB_BENCH_NCCL", "0") != "1":
        exchange = sh.use_peer_exchange(W * len(pairs))
    p0 = torch.from_numpy(pairs[:, 0].astype(np.int64)).to(dev)
    p1 = torch.from_numpy(pairs[:, 1].astype(np.int64)).to(dev)

    def step():
        ms = 0.0
        for nm, ix in calls:
            ll.stat_device(nm, sizes, d_sets.data_ptr(), ix, windows, STAT_BRANCH, outs[nm].data_ptr())
            es = ll.engine_stats()
            ms += es["last_call_ms"]
            phase[nm] = phase.get(nm, 0.0) + es["last_call_ms"]
            worst[nm] = max(worst.get(nm, 0.0), es["last_call_ms"])
            kphase[nm] = kphase.get(nm, np.zeros(6)) + np.array(es["last_kernel_ms"][:6])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if exchange is not None:   # host-synchronous, on the engine's stream
                t_c = time.perf_counter()
                exchange.sum_into(outs[nm], outs[nm], windows)
                coll_wall[0] += (time.perf_counter() - t_c) * 1e3
            else:
                sharding.combine(outs[nm], windows, True)
            if nm == "divergence":  # Fst = 1 - 2 (pi_u + pi_v) / (pi_u + pi_v + 2 d_uv)
                pi = outs["diversity"]
                su = pi[:, p0] + pi[:, p1]
                outs["Fst"] = 1 - 2 * su / (su + 2 * outs["divergence"])
            e1.record()
            coll_ev.append((e0, e1))
        return ms

    def collective_ms():
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in coll_ev)
        coll_ev.clear()
        if exchange is not None:   # the events are on torch's stream and do not see the engine's
            ms = coll_wall[0]
        coll_wall[0] = 0.0
        return ms

    for _ in range(max(1, args.warmup)):
        step()
    collective_ms()
    phase.clear()
    worst.clear()
    kphase.clear()
    sampler = ClockSampler(local, interval_ms=250)  # steps last 0.03-0.3 s: a few samples per step
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    engine_ms = 0.0
    span0, span1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    span0.record()
    for _ in range(args.steps):
        engine_ms += step()
    span1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    coll_ms = collective_ms()
    # device time of the timed region: the span between the two events that bracket it (the engine's
    # calls are synchronous, the collectives run on torch's stream and may overlap the next sweep)
    dev_ms = span0.elapsed_time(span1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        times = torch.tensor([dev_ms, wall, coll_ms, stage_s, gen_s], dtype=torch.float64, device=dev)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dev_ms, wall, coll_ms, stage_s, gen_s = [float(x) for x in times.cpu()]
        mem = torch.tensor([st["device_bytes"], st["num_events"], engine_ms / args.steps], dtype=torch.float64,
                           device=dev)
        g = [torch.empty_like(mem) for _ in range(world)]
        dist.all_gather(g, mem)
        per_rank = [[int(x[0].item()), int(x[1].item()), round(float(x[2].item()), 3)] for x in g]
    else:
        per_rank = [[st["device_bytes"], st["num_events"], round(engine_ms / args.steps, 3)]]
    if rank == 0:
        # parity on a subsample: the reference C library on the rows meeting the first windows
        parity = None
        if not args.no_cpu_baseline:
            parity = c3_parity(t, windows, sets, calls, {k: v.cpu().numpy() for k, v in outs.items()})
        checksum = {k: float(torch.nan_to_num(v).sum().item()) for k, v in outs.items()}
        sweeps = len(calls)
        value = sweeps * nev_total * args.steps / (dev_ms / 1e3)
        line = {
            "metric": "edge-diffs/s (branch-mode f2 + f3 + f4 + Fst over 8 sample sets, windowed)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": ("c3: f2 (28 pairs) + f3 (8) + f4 (2) + Fst (28 pairs), 8 sample sets, branch mode, "
                             f"n={t.num_samples}, L={t.sequence_length:.0f}, E={t.num_edges}, N={t.num_nodes}, "
                             f"{W} windows; {cfg['arms']} arms x {cfg['generations']} generations"),
                "edge_diffs_per_sweep": nev_total, "sweeps_per_step": sweeps,
                "sharding": ("whole genome on one GPU" if world == 1 else
                             f"{world} genome ranges from sharding.plan_shards; per statistic one sum of the "
                             "device-resident partials inside the timed region, "
                             + ("pushed over NVLink peer memory (tskb_exchange_sum)" if exchange is not None
                                else "NCCL all_reduce")),
                "per_rank_plan_bytes_edge_diffs_engine_ms": per_rank,
                "ms_per_step_by_statistic": {k: v / args.steps for k, v in phase.items()},
                "slowest_call_ms_by_statistic_rank0": {k: round(v, 3) for k, v in worst.items()},
                "phases_ms_by_statistic_rank0": {
                    k: dict(zip(["weights", "sweep", "summary", "integrate", "idle", "result"],
                                [round(float(x) / args.steps, 3) for x in v])) for k, v in kphase.items()},
                "collective_ms_per_step": coll_ms / args.steps, "wall_ms_per_step": wall / args.steps * 1e3,
                "engine_ms_per_step_rank0": engine_ms / args.steps,
                "generate_s": gen_s, "concat_s": concat_s, "stage_s": stage_s, "levels": st["num_levels"],
                "checksum": checksum},
            "clocks": clocks, "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def c3_parity(t, windows, sets, calls, got, nwin=3):
    """The first `nwin` windows of every statistic against the reference C library (oracle/_ref) run on
    the rows of the tables that meet those windows."""
    from oracle import ref
    from tskit_b200 import sharding
    if not ref.available():
        return None
    hi = float(windows[nwin])
    sub = sharding.restrict_tables(t, 0.0, hi)
    r = ref.RefTreeSequence(sub)
    w = np.concatenate([windows[: nwin + 1], [t.sequence_length]])
    worst = 0.0
    t0 = time.perf_counter()
    want = {}
    for nm, ix in calls:
        if ix is None:
            want[nm] = r.one_way(nm, sets, windows=w, mode="branch")[:nwin]
        else:
            want[nm] = r.k_way(nm, sets, ix, windows=w, mode="branch")[:nwin]
        scale = np.nanmax(np.abs(want[nm]))
        d = np.abs(got[nm][:nwin] - want[nm])
        # f2/f3/f4 are differences of large terms: absolute floor of rtol x the largest entry
        worst = max(worst, float(np.nanmax(d / np.maximum(np.abs(want[nm]), scale))))
    return {"max_rel_err": worst, "rtol": RTOL, "ok": bool(worst <= RTOL), "windows": nwin,
            "against": "oracle/_ref on the rows meeting the first windows (sharding.restrict_tables)",
            "edges_in_subsample": int(sub.num_edges), "reference_seconds": time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny workload (smoke test of the bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the CPU reference legs (and with them the in-run parity check)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the site / decode / matrix block")
    ap.add_argument("--config", default="c2", choices=["c2", "c3"])
    args = ap.parse_args()
    if args.config == "c3":
        run_c3(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
