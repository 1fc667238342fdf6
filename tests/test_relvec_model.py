"""The device algorithm of genetic_relatedness_vector (the transposed sweep over the replay
plan) replayed on the CPU with the numpy plan model (tests/plan_model.py) and checked against the oracle's
step-by-step restatement of the reference.  Guards the structural claim the kernel rests on: a piece lies inside
the span of every piece it references, so its whole accumulated area belongs to each of them."""
import numpy as np
import pytest

from oracle import port
from tests import fixtures as fx
from tests import plan_model

NO_PIECE = 0xFFFFFFFF


def transposed_sweep(t, W, win):
    m = plan_model.build(t)
    npp, n, K = len(m["q_bp0"]), t.num_samples, W.shape[1]
    lb, q_off, refs = m["level_begin"], m["q_off"], m["refs"]

    def ref_slots(j):
        return refs[q_off[j]:(q_off[j + 1] if j + 1 < len(q_off) else len(refs))]

    state = np.zeros((npp + n + 1, K))
    state[npp:npp + n] = W
    for h in range(len(lb) - 1):  # the sweep: a piece is the sum of the pieces it references
        for j in range(lb[h], lb[h + 1]):
            if m["q_bp1"][j] != NO_PIECE:
                state[j] = state[ref_slots(j)].sum(axis=0)
    out = np.zeros((len(win) - 1, n, K))
    for w in range(len(win) - 1):
        G = np.zeros_like(state)
        for h in range(len(lb) - 2, -1, -1):  # the push: tallest height first
            for j in range(lb[h], lb[h + 1]):
                b1 = m["q_bp1"][j]
                if b1 == NO_PIECE:
                    continue
                x0, x1 = m["bp_pos"][m["q_bp0"][j]], m["bp_pos"][b1]
                for r in ref_slots(j):  # the structural claim, where the referenced slot is a piece
                    if r < npp:
                        assert m["bp_pos"][m["q_bp0"][r]] <= x0 and x1 <= m["bp_pos"][m["q_bp1"][r]]
                length = min(x1, win[w + 1]) - max(x0, win[w])
                g = G[j].copy()
                if length > 0:
                    g += m["q_bl"][j] * length * state[j]
                for r in ref_slots(j):
                    G[r] += g
        out[w] = G[npp:npp + n]
    return out


@pytest.mark.parametrize("name", ["multiroot", "paper", "internal_sample", "unary", "missing"])
def test_transposed_sweep_matches_oracle_on_fixtures(name):
    t = fx.load(name)
    W = np.random.default_rng(1).normal(size=(t.num_samples, 2))
    L = t.sequence_length
    for win in ([0, L], [0, L / 3, L], [L / 4, L / 2]):
        got = transposed_sweep(t, W, win)
        want = port.Oracle(t).genetic_relatedness_vector(W, windows=win, centre=False, span_normalise=False)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), (name, win)


def test_transposed_sweep_matches_oracle_wright_fisher(wf_small):
    W = np.random.default_rng(2).normal(size=(wf_small.num_samples, 1))
    L = wf_small.sequence_length
    win = [L / 7, 0.9 * L]
    got = transposed_sweep(wf_small, W, win)
    want = port.Oracle(wf_small).genetic_relatedness_vector(W, windows=win, centre=False, span_normalise=False)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
