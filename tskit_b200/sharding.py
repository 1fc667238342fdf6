"""Multi-GPU decomposition of the windowed statistics: one process per GPU, the genome cut into
contiguous ranges balanced by edge-diff count.

The path shards naturally: a window's value depends only on the trees overlapping it, and each
rank seeds the state at its range's left edge locally from the tables (the engine's
``genome_range``).  There is no collective on the data path.  The only exchange is the final
sum of the per-window partial results -- windows owned by one rank receive zeros from the
others, windows straddling a cut are the sum of their parts -- done with one ``all_reduce`` of
``W x M`` doubles (NCCL on device-resident partials; gloo in the CPU tests), before span
normalisation (``trees.c:1920-1934`` divides after accumulation).  The relatedness vector shards
the same way: its rows are integrals over the genome, and centring the output rows is linear, so
the ranks' centred partial rows add up to the centred whole.
"""
import numpy as np

from .tables import Tables


def edge_diff_positions(tables):
    """Sorted positions of every edge diff of a left-to-right sweep (``trees.c:1424-1507``):
    each edge is inserted at ``left`` and, unless it reaches the end, removed at ``right``."""
    L = tables.sequence_length
    r = tables.edges_right[tables.edges_right < L]
    pos = np.concatenate([tables.edges_left, r])
    pos.sort()
    return pos


def _diffs_up_to(tables):
    """x -> number of edge diffs at positions <= x, from the two edge indexes (already sorted by
    left / right: no sort of the 2 E positions)."""
    tables.ensure_derived()
    L = tables.sequence_length
    lefts = tables.edges_left[tables.edge_insertion_order]
    rights = tables.edges_right[tables.edge_removal_order]
    rights = rights[:np.searchsorted(rights, L, side="left")]  # edges reaching L are never removed
    return (lambda x: np.searchsorted(lefts, x, side="right") + np.searchsorted(rights, x, side="right"),
            len(lefts) + len(rights))


def plan_shards(tables, windows, world):
    """``world`` contiguous genome ranges covering [0, L) with ~equal edge-diff counts; cuts are
    snapped to window edges when there are at least ``world`` windows (every rank then owns
    whole windows), else they fall inside windows."""
    L = float(tables.sequence_length)
    if world <= 1:
        return [(0.0, L)]
    windows = np.asarray(windows, dtype=np.float64)
    inner = windows[1:-1]
    count, total = _diffs_up_to(tables)
    cuts = []
    if total and len(inner) >= world - 1:
        # the window edge whose cumulative diff count is nearest each target
        c = count(inner)
        for r in range(1, world):
            target = total * r / world
            j = int(np.searchsorted(c, target))
            if j > 0 and (j == len(c) or target - c[j - 1] <= c[j] - target):
                j -= 1
            cuts.append(float(inner[j]))
    elif total:
        # fewer windows than ranks: cut inside windows, at the position of the target diff
        pos = edge_diff_positions(tables)
        cuts = [float(pos[min(len(pos) - 1, (len(pos) * r) // world)]) for r in range(1, world)]
    cuts = sorted(set(c for c in cuts if 0.0 < c < L))
    # degenerate inputs: fall back to even cuts so that every rank has a non-empty range
    if len(cuts) != world - 1:
        cuts = [L * r / world for r in range(1, world)]
    edges = [0.0] + cuts + [L]
    return list(zip(edges[:-1], edges[1:]))


def restrict_tables(tables, a, b):
    """The rows a rank needs for the genome range [a, b): the edges meeting it (coordinates
    untouched: the engine clips and seeds the tree at ``a`` itself), the sites inside it with their
    mutations, and the whole node table (ids are global).  Edge indexes are the sub-sequences of
    the global ones, so the order of the diffs is the reference's."""
    tables.ensure_derived()
    keep = (tables.edges_left < b) & (tables.edges_right > a)
    new_id = np.cumsum(keep, dtype=np.int64) - 1

    def sub_order(order):
        o = order[keep[order]]
        return new_id[o].astype(np.int32)

    kw = {}
    if tables.num_sites:
        s0 = int(np.searchsorted(tables.sites_position, a, side="left"))
        s1 = int(np.searchsorted(tables.sites_position, b, side="left"))
        m0 = int(np.searchsorted(tables.mutations_site, s0, side="left"))
        m1 = int(np.searchsorted(tables.mutations_site, s1, side="left"))
        ao, do = tables.sites_ancestral_state_offset, tables.mutations_derived_state_offset
        kw = dict(
            sites_position=tables.sites_position[s0:s1],
            sites_ancestral_state=tables.sites_ancestral_state[int(ao[s0]):int(ao[s1])],
            sites_ancestral_state_offset=ao[s0:s1 + 1] - ao[s0],
            mutations_site=tables.mutations_site[m0:m1] - s0,
            mutations_node=tables.mutations_node[m0:m1],
            mutations_parent=np.where(tables.mutations_parent[m0:m1] >= 0,
                                      tables.mutations_parent[m0:m1] - m0, -1),
            mutations_derived_state=tables.mutations_derived_state[int(do[m0]):int(do[m1])],
            mutations_derived_state_offset=do[m0:m1 + 1] - do[m0])
    return Tables(tables.sequence_length, tables.nodes_flags, tables.nodes_time,
                  tables.edges_left[keep], tables.edges_right[keep], tables.edges_parent[keep],
                  tables.edges_child[keep], time_uncalibrated=tables.time_uncalibrated,
                  edge_insertion_order=sub_order(tables.edge_insertion_order),
                  edge_removal_order=sub_order(tables.edge_removal_order), **kw)


_SPANS = {}


def _spans(windows, like, window_axis=0):
    """Window spans as a tensor on ``like``'s device, broadcastable against it along ``window_axis``
    (cached: the same windows are used call after call)."""
    import torch
    w = np.ascontiguousarray(windows, dtype=np.float64)
    key = (hash(w.tobytes()), str(like.device), like.dim(), window_axis)
    t = _SPANS.get(key)
    if t is None:
        if len(_SPANS) > 16:
            _SPANS.clear()
        shape = [1] * like.dim()
        shape[window_axis] = len(w) - 1
        t = torch.from_numpy((w[1:] - w[:-1]).reshape(shape)).to(like.device)
        if t.is_cuda:
            # the engine's kernels read it on their own (non-blocking) stream: the upload must have landed
            torch.cuda.current_stream(t.device).synchronize()
        _SPANS[key] = t
    return t


def combine(local, windows, span_normalise, group=None, device=None, window_axis=0):
    """Sum the ranks' un-normalised ``(W, M)`` (relatedness vector: ``(W, nodes, K)``) partial results
    and span-normalise.  ``local`` is this rank's result computed with ``span_normalise=False`` over
    its own genome range: a numpy array (moved to ``device`` for NCCL, returned as numpy) or a
    tensor already resident on the device (summed and normalised in place, returned as is: no host
    round trip; several statistics stacked along another axis go through in ONE all_reduce, with
    ``window_axis`` naming the axis of the windows)."""
    import torch
    import torch.distributed as dist
    on_device = isinstance(local, torch.Tensor)
    total = local if on_device else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if not on_device and device is not None:
            total = total.to(device)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    if span_normalise:
        total /= _spans(windows, total, window_axis)
    if on_device:
        return total
    return total.cpu().numpy().copy()


class ShardedTreeSequence:
    """Same low-level statistics methods as ``lowlevel.LLTreeSequence``; this rank computes its
    genome range, every rank returns the full result."""

    def __init__(self, tables, windows_for_cuts, rank, world, device=0, engine_factory=None,
                 group=None, restrict=True):
        self.ranges = plan_shards(tables, windows_for_cuts, world)
        self.range = self.ranges[rank]
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        self.num_samples = tables.num_samples
        if engine_factory is None:
            from .lowlevel import LLTreeSequence
            engine_factory = lambda t, rng: LLTreeSequence(t, device=device, genome_range=rng)  # noqa: E731
        local = restrict_tables(tables, *self.range) if (restrict and world > 1) else tables
        self.local_tables = local
        self.engine = engine_factory(local, self.range)
        self._bufs = {}

    def stat(self, name, *args, windows, span_normalise=True, **kwargs):
        local = getattr(self.engine, name)(*args, windows=windows, span_normalise=False, **kwargs)
        dev = None
        try:
            import torch
            if torch.cuda.is_available():
                dev = f"cuda:{self.device}"
        except Exception:
            pass
        return combine(local, windows, span_normalise, group=self.group, device=dev)

    # ---- device-resident path (sample-count statistics): sets in, result out through pinned host
    # memory once each; the partial result never leaves HBM before the all_reduce
    def _buf(self, key, shape, dtype, pinned=False):
        import torch
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape):
            if pinned:
                b = torch.empty(shape, dtype=dtype).pin_memory()
            else:
                b = torch.empty(shape, dtype=dtype, device=f"cuda:{self.device}")
            self._bufs[key] = b
        return b

    def stat_device(self, name, sizes, d_sets, indexes, windows, options, out=None):
        """This rank's partial of a sample-count statistic over sets already in HBM (``d_sets``: int32
        tensor), summed over the ranks and span-normalised on the device.  Returns the ``(W, M)``
        result tensor (``out`` if given)."""
        import torch
        from .lowlevel import STAT_SPAN_NORMALISE
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        idx = None if indexes is None else np.ascontiguousarray(indexes, dtype=np.int32)
        w = np.ascontiguousarray(windows, dtype=np.float64)
        M = len(sizes) if idx is None else idx.shape[0]
        if out is None:
            out = self._buf(("res", name, M), (len(w) - 1, M), torch.float64)
        self.engine.stat_device(name, sizes, d_sets.data_ptr(), idx, w,
                                options & ~STAT_SPAN_NORMALISE, out.data_ptr())
        ex = getattr(self, "_exchange", None)
        if ex is not None and out.numel() <= ex.count:
            return ex.sum_into(out, out, w if (options & STAT_SPAN_NORMALISE) else None)
        return combine(out, w, bool(options & STAT_SPAN_NORMALISE), group=self.group)

    def use_peer_exchange(self, count):
        """Sum the partials of ``stat_device`` / ``stat_host`` over NVLink peer memory (``PeerExchange``)
        instead of an NCCL all_reduce, for results of at most ``count`` doubles.  Collective: every rank
        of the group must call it.  Returns the exchange, or None (on every rank alike, the all_reduce
        stays) when some rank could not map its peers -- GPUs without peer access, IPC not permitted."""
        import torch
        import torch.distributed as dist
        ex, why = None, ""
        try:
            ex = PeerExchange(self.engine, count, self.rank, self.world, device=self.device, group=self.group)
        except Exception as e:  # noqa: BLE001 -- reported, and agreed on below
            why = f"{type(e).__name__}: {e}"
        if self.world > 1:
            dev = f"cuda:{self.device}" if torch.cuda.is_available() else "cpu"
            ok = torch.tensor([0 if ex is None else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if ex is not None:
                    ex.close()
                ex = None
        self._exchange = ex
        self.exchange_note = why
        return ex

    def stat_host(self, name, sizes, sets, indexes, windows, options):
        """The call a user makes: host sample sets in, full host result out on every rank."""
        import torch
        sets = np.ascontiguousarray(sets, dtype=np.int32)
        h_sets = self._buf(("hsets", len(sets)), (len(sets),), torch.int32, pinned=True)
        h_sets.numpy()[:] = sets
        d_sets = self._buf(("dsets", len(sets)), (len(sets),), torch.int32)
        d_sets.copy_(h_sets, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the engine runs on its own stream
        res = self.stat_device(name, sizes, d_sets, indexes, windows, options)
        h_res = self._buf(("hres", name, tuple(res.shape)), tuple(res.shape), torch.float64, pinned=True)
        h_res.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h_res.numpy()

    # ---- matrices
    def divergence_matrix(self, windows, sample_sets=None, sample_set_sizes=None, mode=None,
                          span_normalise=True):
        """``divergence_matrix`` over the whole genome from per-range partials: this rank contracts the
        sites (site mode: exact integer counts of differences, so the sum is bit-identical for any
        number of ranks) or sweeps the trees (branch mode) of its own range, one ``all_reduce`` adds
        the ``(W, n, n)`` partials, span normalisation follows the sum (``trees.c:8876-8899``)."""
        return self.stat("divergence_matrix", windows=windows, sample_sets=sample_sets,
                         sample_set_sizes=sample_set_sizes, mode=mode, span_normalise=span_normalise)

    def genotype_matrix(self, samples=None, isolated_as_missing=True, gather=True):
        """Genotype decode sharded by site (``genotypes.c:473-594``; sites are independent): this rank
        decodes the sites of its range.  ``gather=False`` returns ``(first_site, block)`` -- the rows
        stay sharded -- else every rank gets the whole ``(num_sites, n)`` int8 matrix (all_gather of
        the blocks, padded to the largest)."""
        import torch
        import torch.distributed as dist
        block = self.engine.genotype_matrix(samples=samples, isolated_as_missing=isolated_as_missing)
        counts = [self.local_tables.num_sites]
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi:
            counts = [None] * self.world
            dist.all_gather_object(counts, self.local_tables.num_sites, group=self.group)
        first = int(sum(counts[: self.rank])) if multi else 0
        if not gather or not multi:
            return (first, block) if not gather else block
        rows, n = max(counts), block.shape[1]
        dev = f"cuda:{self.device}" if torch.cuda.is_available() else "cpu"
        pad = torch.zeros((rows, n), dtype=torch.int8, device=dev)
        pad[: block.shape[0]] = torch.from_numpy(block).to(dev)
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


class PeerExchange:
    """Sum of the ranks' device-resident partials over NVLink peer memory (``tskb_exchange_*``): every
    rank pushes its partial into a receive slab on every peer, then adds the world's slots in rank
    order -- no NCCL launch on the path, and the same bits on every rank.  One process per GPU of one
    node: the slabs' CUDA IPC handles (64 bytes each) are exchanged once over the process group and
    mapped by the library from this rank's device.  ``members`` instead connects exchange objects of
    THIS process (one per device or stream; no IPC).  ``count`` doubles per call at most."""

    ASYNC = 1

    def __init__(self, engine, count, rank, world, device=0, group=None, connect=True):
        import ctypes as C

        from . import _lib
        from .lowlevel import _handle
        self.engine, self.count, self.rank, self.world, self.device = engine, int(count), rank, world, device
        L = _lib.lib()
        self._x = None
        err = None
        try:
            h = C.c_void_p()
            _handle(L.tskb_exchange_create(int(device), self.count, world, rank, C.byref(h)))
            self._x = h
        except Exception as e:  # noqa: BLE001 -- the collectives below still have to be entered
            err = e
        if world > 1 and connect:
            import torch.distributed as dist
            mine = C.create_string_buffer(64)
            if err is None:
                try:
                    _handle(L.tskb_exchange_get_handle(self._x, mine))
                except Exception as e:  # noqa: BLE001
                    err = e
            everyone = [None] * world
            dist.all_gather_object(everyone, None if err is not None else mine.raw, group=group)
            if err is None and all(h is not None for h in everyone):
                try:
                    _handle(L.tskb_exchange_connect(self._x, b"".join(everyone)))
                except Exception as e:  # noqa: BLE001
                    err = e
            elif err is None:
                err = RuntimeError("PeerExchange: a peer could not export its receive slab")
            dist.barrier(group=group)   # every slab is mapped everywhere before the first push
        if err is not None:
            self.close()
            raise err

    @staticmethod
    def connect_local(members):
        """Connect exchange objects created in this process with ``connect=False`` (rank order)."""
        import ctypes as C

        from . import _lib
        from .lowlevel import _handle
        arr = (C.c_void_p * len(members))(*[m._x for m in members])
        for m in members:
            _handle(_lib.lib().tskb_exchange_connect_local(m._x, arr))

    def sum_into(self, local, out, windows=None, window_axis=0, wait=True):
        """``out`` = sum over the ranks of ``local`` (both device tensors of at most ``count`` doubles;
        they may be the same tensor), span-normalised along ``window_axis`` when ``windows`` is given.
        Runs on the engine's stream; ``wait=False`` returns without waiting for the device (``status()``
        later reports a peer that never delivered)."""
        import ctypes as C

        from . import _lib
        from .lowlevel import _handle
        n = local.numel()
        if n > self.count or out.numel() != n or not local.is_contiguous() or not out.is_contiguous():
            raise ValueError("PeerExchange: tensors must be contiguous and at most `count` doubles")
        spans, stride, scount = None, 1, 1
        if windows is not None:
            sp = _spans(windows, local, window_axis).reshape(-1)
            spans, scount = sp.data_ptr(), sp.numel()
            stride = 1
            for d in local.shape[window_axis + 1:]:
                stride *= d
        _handle(_lib.lib().tskb_exchange_sum(
            self._x, None if self.engine is None else self.engine._h, C.c_void_p(local.data_ptr()), n,
            None if spans is None else C.c_void_p(spans), stride, scount, C.c_void_p(out.data_ptr()),
            0 if wait else self.ASYNC))
        return out

    def status(self):
        from . import _lib
        from .lowlevel import _handle
        _handle(_lib.lib().tskb_exchange_status(self._x, None if self.engine is None else self.engine._h))

    def close(self):
        if getattr(self, "_x", None):
            from . import _lib
            _lib.lib().tskb_exchange_free(self._x)
            self._x = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
