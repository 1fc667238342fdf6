"""CPU suite: the C-ABI library loads and exports every symbol include/tskit_b200.h declares;
without a device the engine refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import HAS_GPU, ROOT
from tskit_b200 import _lib


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "tskit_b200.h")).read()
    names = set(re.findall(r"\b(tskb_[a-zA-Z0-9_]+)\s*\(", txt))
    return sorted(n for n in names if not n.endswith("_t"))  # function types (typedefs) are not symbols


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), name
    assert sorted(_lib.SYMBOLS) == names


def test_strerror_messages():
    L = _lib.lib()
    assert b"TSK_ERR_BAD_WINDOWS" in L.tskb_strerror(-901)
    assert b"TSK_ERR_DUPLICATE_SAMPLE" in L.tskb_strerror(-600)
    assert b"no CPU fallback" in L.tskb_strerror(-20004)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_no_device_means_error_not_fallback(wf_small):
    from tskit_b200.lowlevel import LibraryError, LLTreeSequence
    with pytest.raises(LibraryError) as e:
        LLTreeSequence(wf_small)
    assert e.value.code == -20004


def test_lowlevel_argument_parsing():
    """The extension-level checks of _tskitmodule.c:6588-6636, 799-851 (no device needed)."""
    from tskit_b200 import lowlevel as ll
    assert ll.parse_stats_mode(None) == 1 and ll.parse_stats_mode("branch") == 2
    with pytest.raises(ValueError):
        ll.parse_stats_mode("bogus")
    with pytest.raises(ValueError):
        ll.parse_windows([0.0])
    with pytest.raises(ValueError):
        ll.parse_sample_sets([2, 2], [0, 1, 2])
    sizes, sets = ll.parse_sample_sets([1, 2], [0, 1, 2])
    assert sizes.dtype == np.uint64 and sets.dtype == np.int32
