"""Plain column container for the tables the statistics path reads.

The reference keeps these as SoA columns (``c/tskit/tables.h:309-601``) and the
Python ``TreeSequence`` re-exports them as numpy views (``python/tskit/trees.py:
4180-4209``).  This class is the hand-off format between the host product and
the engine: exactly those columns, nothing else.
"""
from __future__ import annotations

import dataclasses
import io

import numpy as np

NODE_IS_SAMPLE = 1
NULL = -1


def _ragged(strings):
    data = b"".join(strings)
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    if len(strings):
        off[1:] = np.cumsum([len(s) for s in strings])
    return np.frombuffer(data, dtype=np.int8).copy(), off


@dataclasses.dataclass
class Tables:
    sequence_length: float
    nodes_flags: np.ndarray
    nodes_time: np.ndarray
    edges_left: np.ndarray
    edges_right: np.ndarray
    edges_parent: np.ndarray
    edges_child: np.ndarray
    sites_position: np.ndarray = None
    sites_ancestral_state: np.ndarray = None          # int8 bytes
    sites_ancestral_state_offset: np.ndarray = None   # uint64 [S+1]
    mutations_site: np.ndarray = None
    mutations_node: np.ndarray = None
    mutations_parent: np.ndarray = None               # may be None -> computed
    mutations_derived_state: np.ndarray = None
    mutations_derived_state_offset: np.ndarray = None
    time_uncalibrated: bool = False
    # edge indexes (tables.c:11392-11459); None -> built by build_indexes()
    edge_insertion_order: np.ndarray = None
    edge_removal_order: np.ndarray = None

    def __post_init__(self):
        c = np.ascontiguousarray
        self.nodes_flags = c(self.nodes_flags, dtype=np.uint32)
        self.nodes_time = c(self.nodes_time, dtype=np.float64)
        self.edges_left = c(self.edges_left, dtype=np.float64)
        self.edges_right = c(self.edges_right, dtype=np.float64)
        self.edges_parent = c(self.edges_parent, dtype=np.int32)
        self.edges_child = c(self.edges_child, dtype=np.int32)
        if self.sites_position is None:
            self.sites_position = np.zeros(0, dtype=np.float64)
        self.sites_position = c(self.sites_position, dtype=np.float64)
        S = len(self.sites_position)
        if self.sites_ancestral_state is None:
            self.sites_ancestral_state = np.full(S, ord("0"), dtype=np.int8)
            self.sites_ancestral_state_offset = np.arange(S + 1, dtype=np.uint64)
        self.sites_ancestral_state = c(self.sites_ancestral_state, dtype=np.int8)
        self.sites_ancestral_state_offset = c(
            self.sites_ancestral_state_offset, dtype=np.uint64)
        if self.mutations_site is None:
            self.mutations_site = np.zeros(0, dtype=np.int32)
            self.mutations_node = np.zeros(0, dtype=np.int32)
        self.mutations_site = c(self.mutations_site, dtype=np.int32)
        self.mutations_node = c(self.mutations_node, dtype=np.int32)
        Mu = len(self.mutations_site)
        if self.mutations_derived_state is None:
            self.mutations_derived_state = np.full(Mu, ord("1"), dtype=np.int8)
            self.mutations_derived_state_offset = np.arange(Mu + 1, dtype=np.uint64)
        self.mutations_derived_state = c(self.mutations_derived_state, dtype=np.int8)
        self.mutations_derived_state_offset = c(
            self.mutations_derived_state_offset, dtype=np.uint64)
        if self.mutations_parent is not None:
            self.mutations_parent = c(self.mutations_parent, dtype=np.int32)
        if self.edge_insertion_order is not None:
            self.edge_insertion_order = c(self.edge_insertion_order, dtype=np.int32)
            self.edge_removal_order = c(self.edge_removal_order, dtype=np.int32)

    # ------------------------------------------------------------------ sizes
    @property
    def num_nodes(self):
        return len(self.nodes_time)

    @property
    def num_edges(self):
        return len(self.edges_left)

    @property
    def num_sites(self):
        return len(self.sites_position)

    @property
    def num_mutations(self):
        return len(self.mutations_site)

    @property
    def samples(self):
        """Sample node ids in id order (``init_nodes``, c/tskit/trees.c:404-453)."""
        return np.nonzero(self.nodes_flags & NODE_IS_SAMPLE)[0].astype(np.int32)

    @property
    def num_samples(self):
        return int(np.count_nonzero(self.nodes_flags & NODE_IS_SAMPLE))

    # ---------------------------------------------------------------- derived
    def build_indexes(self):
        """Edge insertion/removal orders with the reference's sort keys
        (``tsk_table_collection_build_index``, c/tskit/tables.c:11392-11459):
        insertion by (left, time[parent], parent, child); removal by
        (right, -time[parent], -parent, -child)."""
        t = self.nodes_time[self.edges_parent]
        p = self.edges_parent.astype(np.int64)
        ch = self.edges_child.astype(np.int64)
        self.edge_insertion_order = np.lexsort(
            (ch, p, t, self.edges_left)).astype(np.int32)
        self.edge_removal_order = np.lexsort(
            (-ch, -p, -t, self.edges_right)).astype(np.int32)
        return self

    def compute_mutation_parents(self):
        """Fill ``mutations_parent``: the closest mutation at the same site on
        the path to the root (semantics of
        ``tsk_table_collection_compute_mutation_parents``, tables.c).  Pure
        host logic, O(path) per mutation; used only when a caller did not
        supply the column."""
        Mu = self.num_mutations
        par = np.full(Mu, NULL, dtype=np.int32)
        if Mu == 0:
            self.mutations_parent = par
            return self
        counts = np.bincount(self.mutations_site, minlength=self.num_sites)
        if counts.max(initial=0) <= 1:
            self.mutations_parent = par
            return self
        # sites with several mutations: walk the marginal tree
        order = np.lexsort((self.edges_left, self.edges_child))
        cl = self.edges_left[order]
        cr = self.edges_right[order]
        cp = self.edges_parent[order]
        cc = self.edges_child[order]
        start = np.searchsorted(cc, np.arange(self.num_nodes + 1))

        def parent_at(u, x):
            a, b = start[u], start[u + 1]
            k = np.searchsorted(cl[a:b], x, side="right") - 1
            if k >= 0 and cr[a + k] > x:
                return int(cp[a + k])
            return NULL

        first = np.concatenate([[0], np.cumsum(counts)])
        for s in np.nonzero(counts > 1)[0]:
            x = self.sites_position[s]
            last = {}
            for m in range(first[s], first[s + 1]):
                u = int(self.mutations_node[m])
                v = u
                found = NULL
                # a mutation over the same node listed earlier is the parent
                while v != NULL:
                    if v in last and (v != u or last[v] != m):
                        found = last[v]
                        break
                    v = parent_at(v, x)
                par[m] = found
                last[u] = m
        self.mutations_parent = par
        return self

    def ensure_derived(self):
        if self.edge_insertion_order is None:
            self.build_indexes()
        if self.mutations_parent is None:
            self.compute_mutation_parents()
        return self

    # ------------------------------------------------------------------- text
    @classmethod
    def from_text(cls, nodes, edges, sites=None, mutations=None, sequence_length=0,
                  time_uncalibrated=False):
        """Parse the whitespace-separated text tables used by the reference's
        fixtures (``c/tests/testlib.c:31-400``; ``tskit.load_text``)."""
        def rows(txt):
            lines = [ln.split() for ln in io.StringIO(txt) if ln.strip()]
            if lines and not _is_number(lines[0][0]):
                header, lines = lines[0], lines[1:]
            else:
                header = None
            return header, lines

        def col(header, default_order, name):
            return (header or default_order).index(name)

        h, r = rows(nodes)
        order = ["is_sample", "time", "population", "individual"]
        i_s, i_t = col(h, order, "is_sample"), col(h, order, "time")
        flags = np.array([int(x[i_s]) for x in r], dtype=np.uint32)
        time = np.array([float(x[i_t]) for x in r], dtype=np.float64)
        h, r = rows(edges)
        order = ["left", "right", "parent", "child"]
        il, ir, ip, ic = (col(h, order, k) for k in order)
        el, er, ep, ec = [], [], [], []
        for x in r:
            for child in x[ic].split(","):
                el.append(float(x[il]))
                er.append(float(x[ir]))
                ep.append(int(x[ip]))
                ec.append(int(child))
        kw = {}
        if sites is not None:
            h, r = rows(sites)
            order = ["position", "ancestral_state"]
            ipos, ia = col(h, order, "position"), col(h, order, "ancestral_state")
            kw["sites_position"] = np.array([float(x[ipos]) for x in r])
            a, off = _ragged([x[ia].encode() for x in r])
            kw["sites_ancestral_state"], kw["sites_ancestral_state_offset"] = a, off
        if mutations is not None:
            h, r = rows(mutations)
            order = ["site", "node", "derived_state", "parent"]
            isite, inode, ider = (col(h, order, k) for k in order[:3])
            kw["mutations_site"] = np.array([int(x[isite]) for x in r], dtype=np.int32)
            kw["mutations_node"] = np.array([int(x[inode]) for x in r], dtype=np.int32)
            d, off = _ragged([x[ider].encode() for x in r])
            kw["mutations_derived_state"], kw["mutations_derived_state_offset"] = d, off
            hh = h or order
            if "parent" in hh and all(len(x) > hh.index("parent") for x in r):
                ipar = hh.index("parent")
                kw["mutations_parent"] = np.array(
                    [int(x[ipar]) for x in r], dtype=np.int32)
        L = sequence_length
        if not L:
            L = max(er) if er else 1.0
        t = cls(L, flags, time, el, er, ep, ec, time_uncalibrated=time_uncalibrated, **kw)
        t.sort_edges()
        return t

    def sort_edges(self):
        """Canonical edge order (time[parent], parent, child, left) required by
        the reference (``tsk_table_collection_sort``)."""
        o = np.lexsort((self.edges_left, self.edges_child, self.edges_parent,
                        self.nodes_time[self.edges_parent]))
        self.edges_left = self.edges_left[o]
        self.edges_right = self.edges_right[o]
        self.edges_parent = self.edges_parent[o]
        self.edges_child = self.edges_child[o]
        self.edge_insertion_order = None
        self.edge_removal_order = None
        return self

    @classmethod
    def from_tskit(cls, ts):
        """Columns of a ``tskit.TreeSequence`` (``trees.py:4180-4209``)."""
        t = ts.tables
        return cls(
            ts.sequence_length, t.nodes.flags, t.nodes.time, t.edges.left,
            t.edges.right, t.edges.parent, t.edges.child,
            sites_position=t.sites.position,
            sites_ancestral_state=t.sites.ancestral_state,
            sites_ancestral_state_offset=t.sites.ancestral_state_offset,
            mutations_site=t.mutations.site, mutations_node=t.mutations.node,
            mutations_parent=t.mutations.parent,
            mutations_derived_state=t.mutations.derived_state,
            mutations_derived_state_offset=t.mutations.derived_state_offset,
            time_uncalibrated=(ts.time_units == "uncalibrated"),
            edge_insertion_order=t.indexes.edge_insertion_order,
            edge_removal_order=t.indexes.edge_removal_order)

    def save(self, path):
        d = {f.name: getattr(self, f.name) for f in dataclasses.fields(self)
             if getattr(self, f.name) is not None}
        np.savez(path, **d)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        d = {k: z[k] for k in z.files}
        d["sequence_length"] = float(d["sequence_length"])
        d["time_uncalibrated"] = bool(d.get("time_uncalibrated", False))
        return cls(**d)


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False
