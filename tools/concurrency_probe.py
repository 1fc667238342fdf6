"""Experiment: do the sweep and the branch summary overlap when they run side by side?  Two engines
on one GPU (own streams), two host threads issuing branch diversity calls; total calls/s against one
thread.  TSKB_SWEEP_MAX_PER_SM caps the cooperative sweep's resident CTAs so that the other
stream's kernels find room."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tskit_b200.lowlevel import LLTreeSequence
t, W, _ = bench.load_workload("c2")
w = np.linspace(0, t.sequence_length, W + 1)
s = t.samples
sz = np.array([len(s)], dtype=np.uint64)
engines = [LLTreeSequence(t), LLTreeSequence(t)]


def loop(ll, n):
    for _ in range(n):
        ll.diversity(sz, s, windows=w, mode="branch")


for ll in engines:
    loop(ll, 5)
N = 200
t0 = time.perf_counter(); loop(engines[0], N); one = time.perf_counter() - t0
ths = [threading.Thread(target=loop, args=(ll, N)) for ll in engines]
t0 = time.perf_counter()
for th in ths: th.start()
for th in ths: th.join()
two = time.perf_counter() - t0
print(f"cap={os.environ.get('TSKB_SWEEP_MAX_PER_SM')}: one thread {one / N * 1e3:.3f} ms/call; two threads {two / (2 * N) * 1e3:.3f} ms/call "
      f"(speed-up {one / N / (two / (2 * N)):.2f}x)")
