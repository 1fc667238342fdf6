"""Writes tests/golden/wf_small_golden.json by importing the PYTHON reference (tskit built
from /root/reference/python).  Run in the build container only:
    PYTHONPATH=<built reference python dir> python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import tskit

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tskit_b200.tables import Tables  # noqa: E402


def to_tskit(t):
    tc = tskit.TableCollection(t.sequence_length)
    tc.time_units = "generations"
    tc.nodes.set_columns(flags=t.nodes_flags, time=t.nodes_time)
    tc.edges.set_columns(left=t.edges_left, right=t.edges_right, parent=t.edges_parent,
                         child=t.edges_child)
    tc.sites.set_columns(position=t.sites_position, ancestral_state=t.sites_ancestral_state,
                         ancestral_state_offset=t.sites_ancestral_state_offset)
    tc.mutations.set_columns(site=t.mutations_site, node=t.mutations_node,
                             derived_state=t.mutations_derived_state,
                             derived_state_offset=t.mutations_derived_state_offset)
    tc.build_index()
    tc.compute_mutation_parents()
    return tc.tree_sequence()


def main():
    t = Tables.load(os.path.join(os.path.dirname(HERE), "data", "wf_200_500_100000.npz"))
    ts = to_tskit(t)
    s = ts.samples()
    sets = [s[:50], s[50:120], s[120:]]
    W = np.linspace(0, ts.sequence_length, 8)
    idx = {"divergence": [[0, 1], [1, 2], [0, 0]], "Y2": [[0, 1], [2, 1]],
           "f2": [[0, 1], [0, 2]], "genetic_relatedness": [[0, 1], [2, 2]],
           "Y3": [[0, 1, 2]], "f3": [[0, 1, 2], [2, 1, 0]], "f4": [[0, 1, 2, 0]]}
    stats = {}
    for mode in ("branch", "site"):
        for name in ("diversity", "segregating_sites", "Y1"):
            stats[f"{name}/{mode}"] = getattr(ts, name)(sets, windows=W, mode=mode).tolist()
        for name, ix in idx.items():
            kw = dict(proportion=False) if name == "genetic_relatedness" else {}
            stats[f"{name}/{mode}"] = getattr(ts, name)(
                sets, indexes=ix, windows=W, mode=mode, **kw).tolist()
    G = ts.genotype_matrix()
    sub = [[int(x)] for x in s[:12]]
    out = dict(
        tskit_version=tskit.__version__, windows=W.tolist(), indexes=idx, stats=stats,
        genotypes_head=G[:64].tolist(),
        divmat_site=ts.divergence_matrix(sub, windows=W[:3], mode="site").tolist(),
        divmat_branch=ts.divergence_matrix(sub, windows=W[:3], mode="branch").tolist())
    json.dump(out, open(os.path.join(HERE, "wf_small_golden.json"), "w"))
    print("wrote", len(stats), "stat arrays")


if __name__ == "__main__":
    main()
