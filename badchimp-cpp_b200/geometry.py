"""In-memory geometry ingest: voxel array -> the reference's per-rank tables.

Replaces the reference's pipeline  PythonScripts/vtklb.py -> ASCII .vtklb -> LBvtk/Grid/Nodes/
BndMpi  for this hot path, with bit-identical numbering:

* node numbering and neighbour table      PythonScripts/vtklb.py:70-124, 182-218, 6-44
* node ranks / types                      src/lbsolver/LBnodes.h:112-138, 141-209
* ghost-node types from the owner         src/lbsolver/LBbndmpi.h:317-331
* bulk / boundary node lists              src/lbsolver/LBgeometry.h:11-21, 37-56
* half-way bounce-back link classes       src/lbsolver/LBhalfwayhelperclass.h:110-161
* processor-boundary exchange lists       src/lbsolver/LBbndmpi.h:182-203, 207-315

`geo[x, y(, z)]` is an integer array: 0 solid, k+1 fluid owned by rank k.  Everything here is
host logic (numpy); the device tables are produced from these by the C-ABI builder.
"""
from __future__ import annotations

import numpy as np

# Direction sets in the order of the reference headers (LBd2q9.h:33, LBd3q19.h:33) which is also
# vtklb.py's system basis (vtklb.py:168-176); D3Q27 follows the same ordering contract.
BASIS = {
    "D2Q9": np.array([[1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1], [1, -1], [0, 0]], dtype=np.int64),
    "D3Q19": np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1],
                       [0, 1, 1], [0, 1, -1], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, 0], [-1, 1, 0],
                       [-1, 0, -1], [-1, 0, 1], [0, -1, -1], [0, -1, 1], [0, 0, 0]], dtype=np.int64),
    "D3Q27": np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1],
                       [0, 1, 1], [0, 1, -1], [1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1],
                       [-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, 0], [-1, 1, 0], [-1, 0, -1], [-1, 0, 1],
                       [0, -1, -1], [0, -1, 1], [-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1],
                       [0, 0, 0]], dtype=np.int64),
}
LATTICE_ID = {"D2Q9": 0, "D3Q19": 1, "D3Q27": 2}


def n_dir_pairs(lattice: str) -> int:
    return (len(BASIS[lattice]) - 1) // 2


def reverse_direction(lattice: str, q: int) -> int:
    nq = len(BASIS[lattice])
    return q if q == nq - 1 else (q + n_dir_pairs(lattice)) % (nq - 1)


def sphere_pack(shape, radius, porosity, seed):
    """Periodic random pack of overlapping spheres (Boolean model): 1 = fluid, 0 = solid.

    The number of spheres is n = -ln(porosity) * V / V_sphere; centres come from
    numpy.random.default_rng(seed).  (the reference's own pore-pack
    scripts are unseeded, so the synthetic inputs are defined here.)
    """
    shape = tuple(int(s) for s in shape)
    nd = len(shape)
    vol = float(np.prod(shape))
    vs = np.pi * radius ** 2 if nd == 2 else 4.0 / 3.0 * np.pi * radius ** 3
    n_sph = max(1, int(round(-np.log(porosity) * vol / vs)))
    rng = np.random.default_rng(seed)
    centres = rng.random((n_sph, nd)) * np.array(shape)
    geo = np.ones(shape, dtype=np.uint8)
    r = int(np.ceil(radius)) + 1
    off = np.arange(-r, r + 1)
    for c in centres:
        base = np.floor(c).astype(np.int64)
        idx = [(base[d] + off) for d in range(nd)]
        d2 = None
        for d in range(nd):
            dd = (idx[d] - c[d]) ** 2
            sh = [1] * nd
            sh[d] = -1
            d2 = dd.reshape(sh) if d2 is None else d2 + dd.reshape(sh)
        inside = d2 < radius ** 2
        sub = np.ix_(*[np.mod(idx[d], shape[d]) for d in range(nd)])
        block = geo[sub]
        block[inside] = 0
        geo[sub] = block
    return geo


def z_slab_rank_map(geo, n_ranks):
    """geo (0 solid / 1 fluid) -> 0 solid, k+1 fluid of rank k; equal-thickness slabs along the last axis
    (equals relperm_input.py:16-35 with nproc=[1,1,P])."""
    geo = (np.asarray(geo) > 0).astype(np.int64)
    nz = geo.shape[-1]
    out = np.zeros_like(geo)
    for k in range(n_ranks):
        lo, hi = (k * nz) // n_ranks, ((k + 1) * nz) // n_ranks
        out[..., lo:hi] = geo[..., lo:hi] * (k + 1)
    return out


def _set_iteration_order(chunks):
    """Iteration order of the Python set the reference builds with
    `rank_set = rank_set.union(set(values))` per direction (vtklb.py:199-203): the same distinct
    values are inserted in the same first-occurrence order, so the hash-table layout -- and with it
    the order in which ghost blocks are appended -- is reproduced."""
    rank_set = set()
    for vals in chunks:
        if vals.size == 0:
            continue
        uniq, first = np.unique(vals, return_index=True)
        ordered = [int(v) for v in uniq[np.argsort(first)]]
        rank_set = rank_set.union(set(ordered))
    return rank_set


class LatticeGeometry:
    """All ranks of one decomposed geometry (the in-memory equivalent of one vtklb(...) call)."""

    def __init__(self, geo, lattice="D3Q19", periodic=""):
        geo = np.asarray(geo)
        self.lattice = lattice
        self.basis = BASIS[lattice]
        self.nq = len(self.basis)
        self.nd = geo.ndim
        if self.basis.shape[1] != self.nd:
            raise ValueError("lattice %s needs a %d-dimensional geometry" % (lattice, self.basis.shape[1]))
        self.shape = geo.shape
        self.n_ranks = int(geo.max())
        self.periodic = periodic.lower()
        pshape = tuple(n + 2 for n in geo.shape)
        inner = tuple([slice(1, -1)] * self.nd)
        self.geo = np.full(pshape, -1, dtype=np.int64)
        self.geo[inner] = geo
        self.bulk = np.zeros(pshape, dtype=bool)
        self.bulk[inner] = True
        self.label = np.zeros(pshape, dtype=np.int64)
        for rank in range(1, self.n_ranks + 1):
            m = (self.geo == rank) & self.bulk
            self.label[m] = np.arange(1, 1 + np.count_nonzero(m))
        for ax, name in enumerate("xyz"[: self.nd]):
            if name in self.periodic:
                self._wrap(self.geo, ax)
                self._wrap(self.label, ax)
        self._ranks = {}

    def _wrap(self, arr, ax):
        sl = [slice(None)] * self.nd

        def at(i):
            s = list(sl)
            s[ax] = i
            return tuple(s)

        arr[at(0)] = arr[at(-2)]
        arr[at(-1)] = arr[at(1)]

    def pad_attribute(self, val):
        """vtklb.append_data_set (vtklb.py:327-345): zero rim + periodic copies."""
        val = np.asarray(val)
        out = np.zeros(tuple(n + 2 for n in val.shape), dtype=val.dtype)
        out[tuple([slice(1, -1)] * self.nd)] = val
        for ax, name in enumerate("xyz"[: self.nd]):
            if name in self.periodic:
                self._wrap(out, ax)
        return out

    def rank(self, r):
        """tables of rank r (0-based), cached"""
        if r not in self._ranks:
            self._ranks[r] = RankTables(self, r)
        return self._ranks[r]

    def all_ranks(self):
        tabs = [self.rank(r) for r in range(self.n_ranks)]
        for t in tabs:
            t.finish_exchange(tabs)
        return tabs


class RankTables:
    """What LBvtk + Grid + Nodes + BndMpi hold for one rank."""

    def __init__(self, g: LatticeGeometry, r: int):
        self.g = g
        self.my_rank = r
        rank = r + 1
        geo, label, basis = g.geo, g.label, g.basis
        nd = g.nd
        # --- vtklb.setup_processor_labels (vtklb.py:182-218)
        own = np.where((geo == rank) & g.bulk)
        order = np.argsort(label[own], kind="stable")
        own = tuple(a[order] for a in own)
        self.n_fluid = len(own[0])
        added = np.zeros(geo.shape, dtype=bool)
        shifted = [tuple(own[d] + int(v[d]) for d in range(nd)) for v in basis]
        for s in shifted:
            added[s] |= geo[s] == 0
        blocks = [own, np.where(added)]
        self.n_solid = len(blocks[1][0])
        rank_set = _set_iteration_order([geo[s] for s in shifted])
        for drop in (-1, 0, rank):
            rank_set.discard(drop)
        self.file_neighbor_order = [int(x) - 1 for x in rank_set]  # order of PROCESSOR blocks in the file
        ghost_counts = []
        for nr in rank_set:
            added[:] = False
            for s in shifted:
                added[s] |= geo[s] == nr
            blk = np.where(added)
            ghost_counts.append(len(blk[0]))
            blocks.append(blk)
        ind_local = tuple(np.concatenate([b[d] for b in blocks]) for d in range(nd))
        self.ind_local = ind_local
        n_points = len(ind_local[0])
        self.size = n_points + 1  # node 0 is the shared dummy (USE_ZERO_GHOST_NODE)
        local_label = np.zeros(geo.shape, dtype=np.int64)
        local_label[ind_local] = np.arange(1, n_points + 1)
        self.pos = np.full((self.size, nd), -1, dtype=np.int32)  # Grid::pos_ (LBgrid.h:127-139)
        self.pos[1:] = (np.array(ind_local).T - 1)
        # --- neighbour table (vtklb.py:6-44 -> LBgrid.h:141-147)
        neigh = np.zeros((self.size, g.nq), dtype=np.int32)
        P = np.array(ind_local)
        shape = np.array(geo.shape).reshape(nd, 1)
        for q, c in enumerate(basis):
            nb = P + c.reshape(nd, 1)
            ok = np.all((nb >= 0) & (nb < shape), axis=0)
            idx = tuple(nb[d][ok] for d in range(nd))
            lab = local_label[idx]
            per = (lab == 0) & (geo[idx] == rank)
            lab = np.where(per, label[idx], lab)
            col = np.zeros(n_points, dtype=np.int64)
            col[ok] = lab
            neigh[1:, q] = col
        self.neigh = neigh
        # --- PARALLEL_COMPUTING blocks (vtklb.py:271-281), then sorted by rank (LBvtk.h:534-544)
        geo_pts = geo[ind_local]
        procs = []
        for nr in rank_set:
            sel = np.nonzero(geo_pts == nr)[0]
            cur = (sel + 1).astype(np.int32)  # local labels of the ghost nodes
            adj = label[tuple(a[sel] for a in ind_local)].astype(np.int32)  # their labels on the owner
            procs.append((int(nr) - 1, cur, adj))
        procs.sort(key=lambda t: t[0])
        self.neig_ranks = [p[0] for p in procs]
        self.cur_proc_nodes = [p[1] for p in procs]
        self.adj_proc_nodes = [p[2] for p in procs]
        # --- Nodes (LBnodes.h:112-209)
        node_rank = np.full(self.size, r, dtype=np.int32)
        node_rank[0] = -1
        for nr, cur in zip(self.neig_ranks, self.cur_proc_nodes):
            node_rank[cur] = nr
        self.node_rank = node_rank
        file_type = np.zeros(self.size, dtype=np.int16)
        file_type[1:] = (geo_pts > 0)
        self.file_nodetype = file_type
        fluid = np.zeros(self.size, dtype=bool)
        fluid[1:] = file_type[1:] != 0
        nb_fluid = fluid[neigh]  # [size, nq]
        typ = np.full(self.size, -1, dtype=np.int16)
        solid_rows = ~fluid
        typ[solid_rows] = np.where(nb_fluid[solid_rows].any(axis=1), 1, 0)
        fl_rows = fluid.copy()
        has_solid = (~nb_fluid).any(axis=1)
        typ[fl_rows] = np.where(has_solid[fl_rows], 2, 3)
        typ[fluid & (node_rank != r)] = 4
        typ[0] = -1
        self.node_type = typ  # ghost types (4) are resolved by finish_exchange
        self._exchange_done = False
        self.send_lists = None

    # ---- BndMpi::setup + setupNodeType (LBbndmpi.h:207-331), evaluated for all ranks at once
    def finish_exchange(self, all_tabs):
        if self._exchange_done:
            return
        g = self.g
        nqnz = g.nq - 1
        # ghost node types come from the owner's classification (LBbndmpi.h:317-331)
        typ = self.node_type
        for nr, cur, adj in zip(self.neig_ranks, self.cur_proc_nodes, self.adj_proc_nodes):
            typ[cur] = all_tabs[nr].node_type_own()[adj]
        # makeDirList (LBbndmpi.h:182-203)
        self.recv_nodes, self.recv_ndir, self.recv_dirs = [], [], []
        for cur in self.cur_proc_nodes:
            nb = self.neigh[cur, :nqnz]
            mine = self.node_rank[nb] == self.my_rank
            self.recv_nodes.append(cur.astype(np.int32))
            self.recv_ndir.append(mine.sum(axis=1).astype(np.int32))
            self.recv_dirs.append(np.nonzero(mine)[1].astype(np.int32))
        self._exchange_done = True

    def node_type_own(self):
        """types of own-rank nodes never change after the constructor"""
        return self.node_type

    def send_side(self, all_tabs):
        """nodesToSend / nDirPerNodeToSend / dirListToSend per neighbour: what the neighbour asked for
        in the handshake (LBbndmpi.h:225-308)."""
        out = []
        for nr in self.neig_ranks:
            other = all_tabs[nr]
            k = other.neig_ranks.index(self.my_rank)
            out.append((other.adj_proc_nodes[k].astype(np.int32), other.recv_ndir[k], other.recv_dirs[k]))
        return out

    # ---- node lists (LBgeometry.h:11-21, 37-56)
    def is_fluid(self):
        return self.node_type > 1

    def bulk_nodes(self):
        n = np.arange(self.size)
        return n[(self.node_type > 1) & (self.node_rank == self.my_rank) & (n > 0)].astype(np.int32)

    def fluid_bnd_nodes(self):
        n = np.arange(self.size)
        return n[(self.node_type == 2) & (self.node_rank == self.my_rank) & (n > 0)].astype(np.int32)

    def solid_bnd_nodes(self):
        n = np.arange(self.size)
        return n[(self.node_type == 1) & (n > 0)].astype(np.int32)

    # ---- BoundaryHalwWayHelper (LBhalfwayhelperclass.h:110-161)
    def halfway_bb(self, bnd_nodes):
        """returns (nodes, nBeta, nGamma, nDelta, links[n, nDirPairs]) with links ordered beta, gamma, delta"""
        g = self.g
        npairs = (g.nq - 1) // 2
        bnd_nodes = np.asarray(bnd_nodes, dtype=np.int32)
        fluid = self.is_fluid()
        fq = fluid[self.neigh[bnd_nodes, :npairs]]
        fr = fluid[self.neigh[bnd_nodes, npairs:2 * npairs]]
        q = np.arange(npairs)
        is_gamma = fq & fr
        is_beta_q = fq & ~fr      # unknown direction is q
        is_beta_r = ~fq & fr      # unknown direction is reverse(q)
        is_delta = ~fq & ~fr
        n = len(bnd_nodes)
        links = np.zeros((n, npairs), dtype=np.int32)
        n_beta = (is_beta_q | is_beta_r).sum(axis=1).astype(np.int32)
        n_gamma = is_gamma.sum(axis=1).astype(np.int32)
        n_delta = is_delta.sum(axis=1).astype(np.int32)
        beta_val = np.where(is_beta_q, q, q + npairs)
        for cls, mask, val, start in ((0, is_beta_q | is_beta_r, beta_val, np.zeros(n, dtype=np.int64)),
                                      (1, is_gamma, np.broadcast_to(q, (n, npairs)), n_beta.astype(np.int64)),
                                      (2, is_delta, np.broadcast_to(q, (n, npairs)), (n_beta + n_gamma).astype(np.int64))):
            rank_in_class = np.cumsum(mask, axis=1) - 1
            rows, cols = np.nonzero(mask)
            links[rows, start[rows] + rank_in_class[rows, cols]] = val[rows, cols]
        return bnd_nodes, n_beta, n_gamma, n_delta, links

    def attribute(self, padded_val):
        """values of a padded attribute array at this rank's nodes, row 0 (dummy node) = 0"""
        out = np.zeros(self.size, dtype=padded_val.dtype)
        out[1:] = padded_val[self.ind_local]
        return out

    # ---- interop: the same file vtklb.py writes (vtklb.py:222-325), for the reference's own reader
    def write_vtklb(self, path, attributes=None, version="na"):
        g = self.g
        with open(path, "w") as fh:
            fh.write("# BADChIMP vtklb Version {}\n".format(version))
            fh.write("Geometry file for process {}\n".format(self.my_rank))
            fh.write("ASCII\n")
            fh.write("DATASET UNSTRUCTURED_LB_GRID\n")
            fh.write("NUM_DIMENSIONS {}\n".format(g.nd))
            fh.write("GLOBAL_DIMENSIONS" + "".join(" {}".format(s) for s in g.geo.shape) + "\n")
            fh.write("USE_ZERO_GHOST_NODE\n")
            fh.write("POINTS {} int\n".format(self.size - 1))
            np.savetxt(fh, self.pos[1:], fmt="%d", delimiter=" ", newline="\n")
            fh.write("LATTICE {} int\n".format(g.nq))
            np.savetxt(fh, g.basis, fmt="%d", delimiter=" ", newline="\n")
            fh.write("NEIGHBORS int\n")
            np.savetxt(fh, self.neigh[1:], fmt="%d", delimiter=" ", newline="\n")
            fh.write("PARALLEL_COMPUTING {}\n".format(self.my_rank))
            by_rank = {nr: (cur, adj) for nr, cur, adj in zip(self.neig_ranks, self.cur_proc_nodes, self.adj_proc_nodes)}
            for nr in self.file_neighbor_order:
                cur, adj = by_rank[nr]
                fh.write("PROCESSOR {} {}\n".format(len(cur), nr))
                np.savetxt(fh, np.array([cur, adj]).T, fmt="%d", delimiter=" ", newline="\n")
            fh.write("POINT_DATA {}\n".format(self.size - 1))
            fh.write("SCALARS nodetype int\n")
            np.savetxt(fh, self.file_nodetype[1:].astype(int), fmt="%d", delimiter=" ", newline="\n")
            for name, val in (attributes or {}).items():
                vals = self.attribute(g.pad_attribute(val))[1:]
                if np.issubdtype(vals.dtype, np.integer):
                    fh.write("SCALARS {} int\n".format(name))
                    np.savetxt(fh, vals, fmt="%d", delimiter=" ", newline="\n")
                else:
                    fh.write("SCALARS {} float\n".format(name))
                    np.savetxt(fh, vals, delimiter=" ", newline="\n")
