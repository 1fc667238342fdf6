"""`.trees` (kastore) file -> memory-mapped columns -> engine (tskit_b200/trees_file.py)."""
import dataclasses

import numpy as np
import pytest

tskit = pytest.importorskip("tskit", reason="baseline/_ref (reference tskit) not installed")

from tskit_b200 import dropin, trees_file  # noqa: E402


@pytest.fixture(scope="module")
def trees_path(wf_small, tmp_path_factory):
    p = tmp_path_factory.mktemp("kas") / "wf.trees"
    dropin.from_tables(wf_small).dump(str(p))
    return str(p)


def test_columns_are_views_of_the_file(trees_path, wf_small):
    """Every column the path reads comes out of the file bit for bit, as a view of the mapping (no
    copy), equal to what the reference's loader hands to tskit.TreeSequence."""
    k = trees_file.load_tables(trees_path)
    ref = dropin.tables_from_tree_sequence(tskit.load(trees_path))
    for f in dataclasses.fields(k):
        a, b = getattr(k, f.name), getattr(ref, f.name)
        if isinstance(a, np.ndarray):
            assert a.dtype == b.dtype and np.array_equal(a, b), f.name
            if f.name.startswith(("edges_", "edge_", "nodes_")) or f.name in ("sites_position", "mutations_site",
                                                                              "mutations_node", "mutations_parent"):
                assert not a.flags["OWNDATA"], f.name  # the large columns are used in place
        else:
            assert a == b, f.name
    store = trees_file.read_kastore(trees_path)
    assert store["format/name"].tobytes() == b"tskit.trees"


def test_malformed_files_are_refused(trees_path, tmp_path):
    raw = open(trees_path, "rb").read()
    for name, data in (("magic", b"XXXX" + raw[4:]), ("short", raw[:-8]), ("tiny", raw[:40]),
                       ("type", raw[:64] + bytes([200]) + raw[65:])):
        p = tmp_path / (name + ".trees")
        p.write_bytes(data)
        with pytest.raises(trees_file.FileFormatError):
            trees_file.read_kastore(str(p))


@pytest.mark.gpu
def test_file_to_device(trees_path, wf_small):
    from oracle import port
    ll = trees_file.load(trees_path)
    o = port.Oracle(wf_small)
    s = wf_small.samples
    w = np.linspace(0, wf_small.sequence_length, 7)
    sizes = np.array([len(s)], dtype=np.uint64)
    for mode in ("branch", "site"):
        assert np.allclose(ll.diversity(sizes, s, windows=w, mode=mode), o.stat("diversity", [s], windows=w, mode=mode),
                           rtol=1e-9, atol=0)
    assert np.array_equal(ll.genotype_matrix(), o.genotype_matrix())
