"""`.trees` file -> HBM without the host product in between.

A `.trees` file is a kastore (``c/subprojects/kastore/kastore.c``): a 64-byte header (magic,
version, item count, file size), one 64-byte descriptor per array (type, key and array offsets and
lengths, ``kastore_read_descriptors``), the keys, then the arrays, 8-byte aligned.  The columns the
statistics path reads are memory-mapped as numpy views (no ``tsk_table_collection_load``, no copy
into table structs) and handed to ``tskb_treeseq_init``, which checks them as ``tsk_treeseq_init`` does
and copies them to the device through its pinned staging buffers straight from the page cache.
"""
import mmap
import struct

import numpy as np

from .tables import Tables

_MAGIC = b"\x89KAS\r\n\x1a\n"
_DTYPES = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32,
           np.float64]   # KAS_INT8 ... KAS_FLOAT64 (kastore.h:108-117)


class FileFormatError(Exception):
    pass


def read_kastore(path):
    """``{key: numpy view}`` over a read-only memory map of the file (``kastore_open`` +
    ``kastore_read_header`` + ``kastore_read_descriptors`` with their consistency checks)."""
    f = open(path, "rb")
    mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    f.close()
    if len(mm) < 64 or mm[:8] != _MAGIC:
        raise FileFormatError("not a kastore file")
    major, minor, num_items, file_size = struct.unpack_from("<HHIQ", mm, 8)
    if major != 1:
        raise FileFormatError(f"kastore version {major}.{minor} not supported")
    if file_size != len(mm) or 64 + 64 * num_items > file_size:
        raise FileFormatError("file size does not match the header")
    out = {}
    offset = 64 + 64 * num_items
    items = []
    for j in range(num_items):
        d = 64 + 64 * j
        typ = mm[d]
        key_start, key_len, array_start, array_len = struct.unpack_from("<QQQQ", mm, d + 8)
        if typ >= len(_DTYPES) or key_start != offset or key_start + key_len > file_size:
            raise FileFormatError("bad item descriptor")
        offset += key_len
        items.append((typ, key_start, key_len, array_start, array_len))
    for typ, key_start, key_len, array_start, array_len in items:
        offset = (offset + 7) // 8 * 8
        dt = np.dtype(_DTYPES[typ])
        if array_start != offset or array_start + array_len * dt.itemsize > file_size:
            raise FileFormatError("bad array packing")
        offset += array_len * dt.itemsize
        key = bytes(mm[key_start:key_start + key_len]).decode()
        out[key] = np.frombuffer(mm, dtype=dt, count=array_len, offset=array_start)
    if offset != file_size:
        raise FileFormatError("trailing bytes")
    return out


def load_tables(path):
    """The columns of a ``.trees`` file as a ``Tables`` of memory-mapped views."""
    k = read_kastore(path)
    need = ["sequence_length", "nodes/flags", "nodes/time", "edges/left", "edges/right", "edges/parent",
            "edges/child", "indexes/edge_insertion_order", "indexes/edge_removal_order"]
    for key in need:
        if key not in k:
            raise FileFormatError(f"{key} missing: not an indexed tree sequence file")
    units = bytes(k["time_units"]).decode() if "time_units" in k else "unknown"
    return Tables(
        float(k["sequence_length"][0]), k["nodes/flags"], k["nodes/time"], k["edges/left"], k["edges/right"],
        k["edges/parent"], k["edges/child"],
        sites_position=k.get("sites/position"), sites_ancestral_state=k.get("sites/ancestral_state"),
        sites_ancestral_state_offset=k.get("sites/ancestral_state_offset"),
        mutations_site=k.get("mutations/site"), mutations_node=k.get("mutations/node"),
        mutations_parent=k.get("mutations/parent"), mutations_derived_state=k.get("mutations/derived_state"),
        mutations_derived_state_offset=k.get("mutations/derived_state_offset"),
        time_uncalibrated=(units == "uncalibrated"),
        edge_insertion_order=k["indexes/edge_insertion_order"], edge_removal_order=k["indexes/edge_removal_order"])


def load(path, device=0, genome_range=None):
    """``.trees`` file -> device-resident ``LLTreeSequence``."""
    from .lowlevel import LLTreeSequence
    return LLTreeSequence(load_tables(path), device=device, genome_range=genome_range)
