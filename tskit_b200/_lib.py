"""ctypes binding of libtskb.so (include/tskit_b200.h).

The shared library is built in-tree (``tskit_b200/csrc/Makefile``) for sm_100a.
There is no fallback: if the library is missing, or no CUDA device is visible,
every compute call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSKB_LIB") or os.path.join(_HERE, "libtskb.so")  # TSKB_LIB: A/B builds

u64, i32p, u64p, f64p, vp = C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p


class Tables(C.Structure):
    """tskb_tables_t"""
    _fields_ = [
        ("sequence_length", C.c_double),
        ("time_uncalibrated", C.c_int32),
        ("num_nodes", C.c_uint64),
        ("node_flags", C.c_void_p),
        ("node_time", C.c_void_p),
        ("num_edges", C.c_uint64),
        ("edge_left", C.c_void_p),
        ("edge_right", C.c_void_p),
        ("edge_parent", C.c_void_p),
        ("edge_child", C.c_void_p),
        ("edge_insertion_order", C.c_void_p),
        ("edge_removal_order", C.c_void_p),
        ("num_sites", C.c_uint64),
        ("site_position", C.c_void_p),
        ("site_ancestral_state", C.c_void_p),
        ("site_ancestral_state_offset", C.c_void_p),
        ("num_mutations", C.c_uint64),
        ("mutation_site", C.c_void_p),
        ("mutation_node", C.c_void_p),
        ("mutation_parent", C.c_void_p),
        ("mutation_derived_state", C.c_void_p),
        ("mutation_derived_state_offset", C.c_void_p),
    ]


class Stats(C.Structure):
    """tskb_stats_t"""
    _fields_ = [
        ("num_events", C.c_uint64),
        ("num_visits", C.c_uint64),
        ("num_levels", C.c_uint64),
        ("stage_ms", C.c_double),
        ("last_call_ms", C.c_double),
        ("last_kernel_ms", C.c_double * 8),
        ("last_launches", C.c_uint64),
        ("device_bytes", C.c_uint64),
    ]


# tskb_general_stat_func_t = general_stat_func_t of c/tskit/trees.h:1032-1033
GENERAL_STAT_FUNC = C.CFUNCTYPE(C.c_int, C.c_uint64, C.POINTER(C.c_double), C.c_uint64, C.POINTER(C.c_double),
                                C.c_void_p)

# every symbol include/tskit_b200.h declares
SYMBOLS = [
    "tskb_treeseq_init", "tskb_treeseq_free", "tskb_strerror", "tskb_last_cuda_error",
    "tskb_treeseq_diversity", "tskb_treeseq_segregating_sites", "tskb_treeseq_Y1",
    "tskb_treeseq_divergence", "tskb_treeseq_Y2", "tskb_treeseq_f2",
    "tskb_treeseq_genetic_relatedness", "tskb_treeseq_Y3", "tskb_treeseq_f3",
    "tskb_treeseq_f4", "tskb_treeseq_sample_count_stat_tabulated",
    "tskb_treeseq_trait_covariance", "tskb_treeseq_trait_correlation",
    "tskb_treeseq_genetic_relatedness_weighted", "tskb_treeseq_genetic_relatedness_vector", "tskb_treeseq_trait_linear_model",
    "tskb_treeseq_allele_frequency_spectrum",
    "tskb_treeseq_divergence_matrix", "tskb_treeseq_genotype_matrix", "tskb_treeseq_decode_sites",
    "tskb_treeseq_general_stat",
    "tskb_treeseq_trees_at", "tskb_treeseq_get_stats", "tskb_treeseq_stat_device",
    "tskb_treeseq_debug_array", "tskb_exchange_create", "tskb_exchange_get_handle", "tskb_exchange_connect",
    "tskb_exchange_connect_local", "tskb_exchange_sum", "tskb_exchange_status", "tskb_exchange_free",
]

_lib = None


def lib():
    """Load libtskb.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `make -C tskit_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.tskb_strerror.restype = C.c_char_p
        L.tskb_strerror.argtypes = [C.c_int]
        L.tskb_last_cuda_error.restype = C.c_char_p
        L.tskb_treeseq_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                        C.c_double, C.c_uint32]
        L.tskb_treeseq_free.argtypes = [C.c_void_p]
        one = [C.c_void_p, u64, C.c_void_p, C.c_void_p, u64, C.c_void_p, C.c_uint32, C.c_void_p]
        for n in ("diversity", "segregating_sites", "Y1", "divergence_matrix"):
            getattr(L, "tskb_treeseq_" + n).argtypes = one
        kway = [C.c_void_p, u64, C.c_void_p, C.c_void_p, u64, C.c_void_p, u64, C.c_void_p,
                C.c_uint32, C.c_void_p]
        for n in ("divergence", "Y2", "f2", "genetic_relatedness", "Y3", "f3", "f4"):
            getattr(L, "tskb_treeseq_" + n).argtypes = kway
        L.tskb_treeseq_sample_count_stat_tabulated.argtypes = [
            C.c_void_p, u64, C.c_void_p, C.c_void_p, u64, u64, C.c_void_p, u64, C.c_void_p,
            C.c_uint32, C.c_void_p]
        for n in ("trait_covariance", "trait_correlation"):
            getattr(L, "tskb_treeseq_" + n).argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_void_p,
                                                        C.c_uint32, C.c_void_p]
        L.tskb_treeseq_allele_frequency_spectrum.argtypes = [
            C.c_void_p, u64, C.c_void_p, C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, C.c_uint32,
            C.c_void_p]
        L.tskb_treeseq_trait_linear_model.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, u64,
                                                      C.c_void_p, C.c_uint32, C.c_void_p]
        L.tskb_treeseq_genetic_relatedness_weighted.argtypes = [
            C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, C.c_void_p, C.c_uint32]
        L.tskb_treeseq_genetic_relatedness_vector.argtypes = [
            C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, C.c_void_p, C.c_uint32]
        L.tskb_treeseq_genotype_matrix.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_uint32,
                                                   C.c_void_p]
        L.tskb_treeseq_decode_sites.argtypes = [C.c_void_p, u64, u64, C.c_void_p, u64, C.c_uint32, C.c_void_p]
        L.tskb_treeseq_trees_at.argtypes = [C.c_void_p, u64, C.c_void_p, C.c_void_p, u64,
                                            C.c_void_p, C.c_void_p]
        L.tskb_treeseq_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.tskb_treeseq_stat_device.argtypes = [C.c_void_p, C.c_int, u64, C.c_void_p, C.c_void_p,
                                               u64, C.c_void_p, u64, C.c_void_p, C.c_uint32,
                                               C.c_void_p]
        L.tskb_treeseq_general_stat.argtypes = [C.c_void_p, u64, C.c_void_p, u64, GENERAL_STAT_FUNC, C.c_void_p,
                                                u64, C.c_void_p, C.c_uint32, C.c_void_p]
        L.tskb_exchange_create.argtypes = [C.c_int, u64, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
        L.tskb_exchange_get_handle.argtypes = [C.c_void_p, C.c_void_p]
        L.tskb_exchange_connect.argtypes = [C.c_void_p, C.c_void_p]
        L.tskb_exchange_connect_local.argtypes = [C.c_void_p, C.c_void_p]
        L.tskb_exchange_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64, C.c_void_p, u64, u64, C.c_void_p,
                                        C.c_uint32]
        L.tskb_exchange_status.argtypes = [C.c_void_p, C.c_void_p]
        L.tskb_exchange_free.argtypes = [C.c_void_p]
        L.tskb_treeseq_debug_array.restype = C.c_int64
        L.tskb_treeseq_debug_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, u64]
        _lib = L
    return _lib
