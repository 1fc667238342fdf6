"""Writes tests/golden/wf_small_relvec_golden.json by importing the PYTHON reference (the unmodified tskit
package installed under baseline/_ref): genetic_relatedness_vector on the small Wright-Fisher input with
deterministic weights.  Run in the build container only:
    python tests/golden/make_golden_relvec.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import tskit  # noqa: E402
from tskit_b200 import dropin  # noqa: E402
from tskit_b200.tables import Tables  # noqa: E402


def weights(n):
    """Deterministic, generator-independent weight columns."""
    j = np.arange(n, dtype=np.float64)
    return np.stack([np.sin(1.0 + j), (j % 7) - 3.0 + 0.25 * np.cos(j * j)], axis=1)


def main():
    t = Tables.load(os.path.join(ROOT, "tests", "data", "wf_200_500_100000.npz")).ensure_derived()
    ts = dropin.from_tables(t)
    L = ts.sequence_length
    W = weights(ts.num_samples)
    windows = [L / 7, L / 3, 0.9 * L]
    nodes = [0, 5, ts.num_nodes - 1, ts.num_nodes // 2, 5]
    out = dict(tskit_version=tskit.__version__, windows=windows, nodes=nodes, results={})
    for centre in (True, False):
        for span in (True, False):
            r = ts.genetic_relatedness_vector(W, windows=windows, mode="branch", centre=centre, span_normalise=span)
            out["results"][f"samples/centre={centre}/span={span}"] = r.tolist()
    r = ts.genetic_relatedness_vector(W, windows=windows, mode="branch", centre=False, nodes=nodes)
    out["results"]["nodes/centre=False/span=True"] = r.tolist()
    json.dump(out, open(os.path.join(HERE, "wf_small_relvec_golden.json"), "w"))
    print("wrote", len(out["results"]), "arrays")


if __name__ == "__main__":
    main()
