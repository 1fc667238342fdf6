"""The window-run branch summaries (k_branch_summary_runs and k_branch_summary_bypos_runs, with
k_runs_reduce + k_runs_finalize) replayed on the CPU over the numpy plan model (tests/plan_model.py), with the kernel's arithmetic -- the
window of a position as min(trunc((x - w0) / step), W - 1) on edges w0 + i * step, register sums per
thread of 8 consecutive pieces, the difference array over entirely covered windows, bins restricted to the
windows an engine's genome range meets -- and checked against the oracle's restatement of the reference.
Guards two claims the kernel rests on: the parts of a piece measured against the nominal edges add up to
the piece whatever window the ulp-tolerant lookup names, and every bin index stays inside the range of
windows make_run_args allots (an index outside would be a silent out-of-bounds reduction on the device)."""
import numpy as np
import pytest

from oracle import port
from tests import fixtures as fx
from tests import plan_model

NO_PIECE = 0xFFFFFFFF
RUN_IPT = 8


def exactly_uniform(w):
    """make_run_args: every edge but the last is w0 + i * step bit for bit, the last one lies beyond."""
    W = len(w) - 1
    step = w[1] - w[0] if W > 1 else w[W] - w[0]
    if not step > 0:
        return False
    for i in range(W):
        if w[0] + np.float64(i) * step != w[i]:
            return False
    return bool(w[W] > w[0] + np.float64(W - 1) * step)


def window_range(w, left, right):
    """make_run_args: [wlo, wlo + Wl), the windows meeting [left, right] with a window of margin."""
    W = len(w) - 1
    i0 = int(np.searchsorted(w, left, side="right"))   # first edge > left
    i1 = int(np.searchsorted(w, right, side="left"))   # first edge >= right
    lo = i0 - 2 if i0 >= 2 else 0
    hi = min(W, i1 + 1)
    wlo = min(lo, W - 1)
    return wlo, max(hi, wlo + 1) - wlo


def piece_states(t, m, sample_set):
    """the sweep on the model: a piece is the sum of the pieces it references (sample counts)"""
    npp, n = len(m["q_bp0"]), t.num_samples
    lb, q_off, refs = m["level_begin"], m["q_off"], m["refs"]
    state = np.zeros(npp + n + 1, dtype=np.int64)
    index = {int(u): i for i, u in enumerate(t.samples)}
    for u in sample_set:
        state[npp + index[int(u)]] = 1
    for h in range(len(lb) - 1):
        for j in range(lb[h], lb[h + 1]):
            if m["q_bp1"][j] != NO_PIECE:
                state[j] = state[refs[q_off[j]:q_off[j + 1]]].sum()
    return state[:npp]


def window_runs(m, state, n, w, left, right, threads=5):
    """diversity, branch mode, unpolarised, not span-normalised, by the kernel's formulation"""
    W = len(w) - 1
    assert exactly_uniform(w)
    first, last = np.float64(w[0]), np.float64(w[W])
    step = np.float64(w[1] - w[0] if W > 1 else w[W] - w[0])
    inv = np.float64(1.0) / step
    wlo, Wl = window_range(w, left, right)
    R, C = np.zeros(Wl), np.zeros(Wl + 1)

    def add(arr, idx, v):
        assert 0 <= idx - wlo < len(arr), (idx, wlo, Wl)   # the claim: inside the allotted windows
        arr[idx - wlo] += v

    def cell(x):
        return min(int((x - first) * inv), W - 1)   # trunc: x >= first

    npp = len(state)
    inv_den = 1.0 / (n * (n - 1.0))
    for th in range(threads):   # grid-stride over groups of RUN_IPT consecutive pieces
        wc, acc = 0, 0.0
        for g in range(th, (npp + RUN_IPT - 1) // RUN_IPT, threads):
            for j in range(g * RUN_IPT, min(npp, (g + 1) * RUN_IPT)):
                b1 = m["q_bp1"][j]
                if b1 == NO_PIECE:
                    continue
                x = float(state[j])
                G = m["q_bl"][j] * (x * (n - x) * inv_den + (n - x) * (n - (n - x)) * inv_den)
                a, e = max(m["bp_pos"][m["q_bp0"][j]], first), min(m["bp_pos"][b1], last)
                if not (a < e and G != 0.0):
                    continue
                w0, w1 = cell(a), cell(e)
                hi0 = last if w0 + 1 >= W else first + np.float64(w0 + 1) * step
                lo1 = first + np.float64(w1) * step
                if w0 != wc:
                    if acc != 0.0:
                        add(R, wc, acc)
                    acc, wc = 0.0, w0
                if w1 == w0:
                    acc += G * (e - a)
                else:
                    add(R, w0, acc + G * (hi0 - a))
                    if w1 > w0 + 1:
                        add(C, w0 + 1, G)
                        add(C, w1, -G)
                    wc, acc = w1, G * (e - lo1)
        if acc != 0.0:
            add(R, wc, acc)
    out = np.zeros(W)
    S = 0.0
    for i in range(Wl):   # k_runs_finalize
        S += C[i]
        out[wlo + i] = R[i] + (w[wlo + i + 1] - w[wlo + i]) * S
    return out


def window_of(w, x, end):
    """run_window_of: the window with w[k] <= x < w[k + 1] (start) or w[k] < x <= w[k + 1] (end)"""
    W = len(w) - 1
    k = int(np.searchsorted(w, x, side="left" if end else "right")) - 1
    return min(max(k, 0), W - 1)


def window_runs_by_position(m, state, n, w, left, right, chunk=7):
    """k_branch_summary_bypos_runs (one column): the pieces sorted by the breakpoint they start at, a
    register sum per walker, the last part of a piece that ends in a later window sent there directly;
    general windows by binary search, uniform ones by arithmetic (run_piece_windows)"""
    W = len(w) - 1
    exact = exactly_uniform(w)
    first, last = np.float64(w[0]), np.float64(w[W])
    step = np.float64(w[1] - w[0] if W > 1 else w[W] - w[0])
    inv = np.float64(1.0) / step
    wlo, Wl = window_range(w, left, right)
    R, C = np.zeros(Wl), np.zeros(Wl + 1)

    def add(arr, idx, v):
        assert 0 <= idx - wlo < len(arr), (idx, wlo, Wl)
        arr[idx - wlo] += v

    real = [j for j in range(len(state)) if m["q_bp1"][j] != NO_PIECE]
    real.sort(key=lambda j: m["q_bp0"][j])   # the summary order
    inv_den = 1.0 / (n * (n - 1.0))
    for c0 in range(0, len(real), chunk):    # one walker per chunk
        wc, acc = None, 0.0
        for j in real[c0:c0 + chunk]:
            a, e = max(m["bp_pos"][m["q_bp0"][j]], first), min(m["bp_pos"][m["q_bp1"][j]], last)
            if not a < e or m["q_bl"][j] == 0.0:
                continue
            if exact:
                w0, w1 = min(int((a - first) * inv), W - 1), min(int((e - first) * inv), W - 1)
                hi0 = last if w0 + 1 >= W else first + np.float64(w0 + 1) * step
                lo1 = first + np.float64(w1) * step
            else:
                w0, w1 = window_of(w, a, False), window_of(w, e, True)
                hi0, lo1 = w[w0 + 1], w[w1]
            d0 = e - a if w1 == w0 else hi0 - a
            d1 = 0.0 if w1 == w0 else e - lo1
            x = float(state[j])
            G = m["q_bl"][j] * (x * (n - x) * inv_den + (n - x) * (n - (n - x)) * inv_den)
            if w0 != wc:
                if wc is not None and acc != 0.0:
                    add(R, wc, acc)
                wc, acc = w0, 0.0
            acc += G * d0
            if w1 != w0 and G != 0.0:
                add(R, w1, G * d1)
                if w1 > w0 + 1:
                    add(C, w0 + 1, G)
                    add(C, w1, -G)
        if wc is not None and acc != 0.0:
            add(R, wc, acc)
    out = np.zeros(W)
    S = 0.0
    for i in range(Wl):
        S += C[i]
        out[wlo + i] = R[i] + (w[wlo + i + 1] - w[wlo + i]) * S
    return out


def check(t, w, ranges):
    o = port.Oracle(t)
    s = t.samples
    want = o.stat("diversity", [s], windows=w, mode="branch", span_normalise=False).reshape(-1)
    got = np.zeros_like(want)
    for a, b in ranges:
        m = plan_model.build(t, a, b)
        got += window_runs(m, piece_states(t, m, s), len(s), np.asarray(w, dtype=np.float64), a, b)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(want).max())), (w, ranges)


def check_by_position(t, w, ranges):
    o = port.Oracle(t)
    s = t.samples
    w = np.asarray(w, dtype=np.float64)
    want = o.stat("diversity", [s], windows=w, mode="branch", span_normalise=False).reshape(-1)
    got = np.zeros_like(want)
    for a, b in ranges:
        m = plan_model.build(t, a, b)
        got += window_runs_by_position(m, piece_states(t, m, s), len(s), w, a, b)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(want).max())), (w, ranges)


@pytest.mark.parametrize("name", ["multiroot", "paper", "internal_sample", "unary", "missing"])
def test_window_runs_match_oracle_on_fixtures(name):
    t = fx.load(name)
    L = t.sequence_length
    for count in (1, 2, 3, 7, 64):   # 64 windows: most pieces cover several windows entirely
        w = np.linspace(0, L, count + 1)
        if not exactly_uniform(w):
            continue
        check(t, w, [(0.0, L)])
        check(t, w, [(0.0, L / 3), (L / 3, L)])                 # cuts inside windows
        if count >= 2:
            check(t, w, [(0.0, w[1]), (w[1], L)])               # a cut on a window edge


def test_window_runs_match_oracle_wright_fisher(wf_small):
    L = wf_small.sequence_length
    check(wf_small, np.linspace(0, L, 41), [(0.0, 21000.5), (21000.5, 50000.0), (50000.0, L)])
    # a step that is not representable: edges are w0 + i * step with two roundings
    w = np.float64(0.0) + np.arange(8, dtype=np.float64) * np.float64(L / 7)
    w[-1] = L
    assert exactly_uniform(w)
    check(wf_small, w, [(0.0, L)])


@pytest.mark.parametrize("name", ["multiroot", "paper", "internal_sample", "unary", "missing"])
def test_window_runs_by_position_match_oracle_on_fixtures(name):
    t = fx.load(name)
    L = t.sequence_length
    for w in (np.linspace(0, L, 2), np.linspace(0, L, 4), np.linspace(0, L, 65),
              np.array([0.0, L / 5, L / 4, 0.9 * L, L])):      # the last one: general windows
        check_by_position(t, w, [(0.0, L)])
        check_by_position(t, w, [(0.0, L / 3), (L / 3, L)])
        check_by_position(t, w, [(0.0, w[1]), (w[1], L)] if len(w) > 2 else [(0.0, L)])


def test_window_runs_by_position_wright_fisher(wf_small):
    L = wf_small.sequence_length
    cuts = [(0.0, 21000.5), (21000.5, 50000.0), (50000.0, L)]
    check_by_position(wf_small, np.linspace(0, L, 41), cuts)
    check_by_position(wf_small, np.array([0.0, 500.0, 21000.5, 30000.0, 77777.0, 90000.0, L]), cuts)


def test_which_windows_are_uniform():
    assert not exactly_uniform(np.array([0.0, 10.5, 11.0, 50000.25, 1e5]))
    assert not exactly_uniform(np.array([0.0, 1.0, 2.5, 3.0]))
    # the last window may be shorter or longer than a step: its right edge is carried separately
    assert exactly_uniform(np.array([0.0, 1.0, 2.0, 2.5]))
    assert exactly_uniform(np.array([0.0, 1.0, 2.0, 3.5]))
    assert exactly_uniform(np.linspace(0, 1e8, 1001)) and exactly_uniform(np.linspace(0, 1e5, 8))
