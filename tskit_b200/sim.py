"""Seeded synthetic inputs: infinite-sites mutation placement here;
the Wright-Fisher ancestry generator lives in tskit_b200/csrc/wfsim.cpp."""
import numpy as np

from .tables import Tables


def add_mutations(t: Tables, target, seed=1):
    """Infinite-sites mutations: `target` draws of an edge with probability
    proportional to span x branch length, position = floor(uniform within the
    edge's span), duplicate positions dropped (one mutation per site, ancestral
    "0", derived "1")."""
    rng = np.random.default_rng(seed)
    span = t.edges_right - t.edges_left
    bl = t.nodes_time[t.edges_parent] - t.nodes_time[t.edges_child]
    wgt = span * bl
    cdf = np.cumsum(wgt)
    e = np.searchsorted(cdf, rng.random(target) * cdf[-1], side="right")
    e = np.minimum(e, t.num_edges - 1)
    pos = np.floor(t.edges_left[e] + rng.random(target) * span[e])
    pos = np.minimum(pos, np.nextafter(t.edges_right[e], -np.inf))
    pos = np.maximum(pos, t.edges_left[e])
    pos, first = np.unique(pos, return_index=True)
    e = e[first]
    S = len(pos)
    t.sites_position = pos.astype(np.float64)
    t.sites_ancestral_state = np.full(S, ord("0"), dtype=np.int8)
    t.sites_ancestral_state_offset = np.arange(S + 1, dtype=np.uint64)
    t.mutations_site = np.arange(S, dtype=np.int32)
    t.mutations_node = t.edges_child[e].astype(np.int32)
    t.mutations_parent = np.full(S, -1, dtype=np.int32)
    t.mutations_derived_state = np.full(S, ord("1"), dtype=np.int8)
    t.mutations_derived_state_offset = np.arange(S + 1, dtype=np.uint64)
    return t
