"""Multi-GPU decomposition of the windowed statistics: one process per GPU, the genome cut into
contiguous ranges balanced by edge-diff count.

The path shards naturally: a window's value depends only on the trees overlapping it, and each
rank seeds the state at its range's left edge locally from the replicated tables (the engine's
``genome_range``).  There is no collective on the data path.  The only exchange is the final
sum of the per-window partial results -- windows owned by one rank receive zeros from the
others, windows straddling a cut are the sum of their parts -- done with one ``all_reduce`` of
``W x M`` doubles (NCCL on GPUs, gloo in the CPU tests), before span normalisation
(``trees.c:1920-1934`` divides after accumulation).  The relatedness vector shards the same way:
its rows are integrals over the genome, and centring the output rows is linear, so the ranks'
centred partial rows add up to the centred whole.
"""
import numpy as np


def edge_diff_positions(tables):
    """Sorted positions of every edge diff of a left-to-right sweep (``trees.c:1424-1507``):
    each edge is inserted at ``left`` and, unless it reaches the end, removed at ``right``."""
    L = tables.sequence_length
    r = tables.edges_right[tables.edges_right < L]
    pos = np.concatenate([tables.edges_left, r])
    pos.sort()
    return pos


def plan_shards(tables, windows, world):
    """``world`` contiguous genome ranges covering [0, L) with ~equal edge-diff counts; cuts are
    snapped to window edges when there are at least ``world`` windows (every rank then owns
    whole windows), else they fall inside windows."""
    L = float(tables.sequence_length)
    if world <= 1:
        return [(0.0, L)]
    pos = edge_diff_positions(tables)
    windows = np.asarray(windows, dtype=np.float64)
    targets = [pos[min(len(pos) - 1, (len(pos) * r) // world)] if len(pos) else L * r / world
               for r in range(1, world)]
    cuts = []
    inner = windows[1:-1]
    for x in targets:
        if len(inner) >= world - 1:
            j = int(np.argmin(np.abs(inner - x)))
            x = float(inner[j])
        cuts.append(float(x))
    cuts = sorted(set(c for c in cuts if 0.0 < c < L))
    # degenerate inputs: fall back to even cuts so that every rank has a non-empty range
    if len(cuts) != world - 1:
        cuts = [L * r / world for r in range(1, world)]
    edges = [0.0] + cuts + [L]
    return list(zip(edges[:-1], edges[1:]))


def combine(local, windows, span_normalise, group=None, device=None):
    """Sum the ranks' un-normalised ``(W, M)`` (relatedness vector: ``(W, nodes, K)``) partial results
    and span-normalise.  ``local`` is
    this rank's result computed with ``span_normalise=False`` over its own genome range."""
    import torch
    import torch.distributed as dist
    total = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if device is not None:
            total = total.to(device)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        total = total.cpu()
    out = total.numpy().copy()
    if span_normalise:
        w = np.asarray(windows, dtype=np.float64)
        out /= (w[1:] - w[:-1]).reshape((-1,) + (1,) * (out.ndim - 1))
    return out


class ShardedTreeSequence:
    """Same low-level statistics methods as ``lowlevel.LLTreeSequence``; this rank computes its
    genome range, every rank returns the full result."""

    def __init__(self, tables, windows_for_cuts, rank, world, device=0, engine_factory=None,
                 group=None):
        self.ranges = plan_shards(tables, windows_for_cuts, world)
        self.range = self.ranges[rank]
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        if engine_factory is None:
            from .lowlevel import LLTreeSequence
            engine_factory = lambda t, rng: LLTreeSequence(t, device=device, genome_range=rng)  # noqa: E731
        self.engine = engine_factory(tables, self.range)

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
