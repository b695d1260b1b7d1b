"""Halo index maps for LazyESN (host side).

The reference builds every group's input with `dask.array.overlap.overlap`
(xesn/lazyesn.py:165-166, :199-200, :235): the chunked domain is extended by
`depth` cells per overlapped axis, from the neighbouring chunk or -- at the
domain edge -- from the boundary rule ("periodic", "reflect" or a constant).
Here the same construction is applied ONCE to an array of flat cell indices, so
that on the device a halo is a plain index gather: negative entries encode the
constant-fill values.  The device never sees dask's graph; it sees

    in_map  (G, I) int32 : cell feeding input i of group g, or -1-k -> fill[k]
    out_map (G, O) int32 : cell predicted by output o of group g

with groups in C order over the group grid and cells flattened in C order inside
a block, exactly the order `_flatten_space` / `_get_inner_mask`
(xesn/lazyesn.py:330-354) produce.
"""
import itertools

import numpy as np


def normalise_boundary(boundary, dims, depth):
    """Per-axis boundary rule for the overlapped space axes.

    `boundary` is what the user gave LazyESN: a string / number for all axes or a
    dict keyed by dimension name (lazyesn.py:257-261)."""
    rules = []
    for ax, name in enumerate(dims):
        if depth[ax] == 0:
            rules.append(None)
            continue
        if isinstance(boundary, dict):
            if name not in boundary:
                raise ValueError(f"LazyESN: no boundary rule for overlapped axis '{name}'")
            rule = boundary[name]
        else:
            rule = boundary
        if isinstance(rule, str) and rule not in ("periodic", "reflect"):
            raise NotImplementedError(f"LazyESN: boundary '{rule}' not recognized, use 'periodic', 'reflect' "
                                      f"or a fill value")
        rules.append(rule)
    return rules


def build_maps(space_shape, chunks, depth, rules):
    """Returns (grid, in_map, out_map, fill).

    space_shape / chunks / depth / rules are per space axis (time excluded)."""
    nsp = len(space_shape)
    for n, c in zip(space_shape, chunks):
        if c <= 0 or n % c:
            raise ValueError(f"LazyESN: domain size {tuple(space_shape)} must be divisible by the chunk sizes "
                             f"{tuple(chunks)}")
    grid = tuple(n // c for n, c in zip(space_shape, chunks))
    n_cells = int(np.prod(space_shape))
    if n_cells >= 2**31:
        raise ValueError("LazyESN: too many cells for int32 maps")
    cells = np.arange(n_cells, dtype=np.int64).reshape(space_shape)

    fills = []
    padded = cells
    for ax in range(nsp):
        d = depth[ax]
        if d == 0:
            continue
        width = [(0, 0)] * nsp
        width[ax] = (d, d)
        rule = rules[ax]
        if rule == "periodic":
            padded = np.pad(padded, width, mode="wrap")
        elif rule == "reflect":
            padded = np.pad(padded, width, mode="symmetric")
        else:
            value = float(rule)
            # a NaN boundary cell is dropped from the input (lazyesn.py:287-289) == contributes 0
            value = 0.0 if np.isnan(value) else value
            if value not in fills:
                fills.append(value)
            padded = np.pad(padded, width, mode="constant", constant_values=-1 - fills.index(value))

    in_rows, out_rows = [], []
    for g in itertools.product(*[range(n) for n in grid]):
        halo = tuple(slice(g[a] * chunks[a], g[a] * chunks[a] + chunks[a] + 2 * depth[a]) for a in range(nsp))
        core = tuple(slice(g[a] * chunks[a], (g[a] + 1) * chunks[a]) for a in range(nsp))
        in_rows.append(padded[halo].reshape(-1))
        out_rows.append(cells[core].reshape(-1))
    in_map = np.ascontiguousarray(np.stack(in_rows).astype(np.int32))
    out_map = np.ascontiguousarray(np.stack(out_rows).astype(np.int32))
    fill = np.asarray(fills if fills else [0.0], dtype=np.float64)
    return grid, in_map, out_map, fill


def wout_to_blocks(wout_groups, grid):
    """(G, O, N) -> the array layout dask assembles in the reference with
    chunks=_Wout_chunks, drop_axis=-1 (lazyesn.py:72-77, :169-185; shapes pinned by
    xesn/test/lazy.py:46-88):  1-D grid -> (O, G*N);  2-D -> (Gx*O, Gy*N);
    3-D -> (Gx*O, Gy*N, Gz)."""
    G, O, N = wout_groups.shape
    w = wout_groups.reshape(tuple(grid) + (O, N))
    if len(grid) == 1:
        return np.ascontiguousarray(w.transpose(1, 0, 2).reshape(O, grid[0] * N))
    if len(grid) == 2:
        return np.ascontiguousarray(w.transpose(0, 2, 1, 3).reshape(grid[0] * O, grid[1] * N))
    if len(grid) == 3:
        return np.ascontiguousarray(w.transpose(0, 3, 1, 4, 2).reshape(grid[0] * O, grid[1] * N, grid[2]))
    raise NotImplementedError("LazyESN: Wout block layout is defined for 1 to 3 space axes")


def blocks_to_wout(blocks, grid, O, N):
    """Inverse of `wout_to_blocks` (used when a stored readout is loaded)."""
    blocks = np.asarray(blocks)
    if len(grid) == 1:
        w = blocks.reshape(O, grid[0], N).transpose(1, 0, 2)
    elif len(grid) == 2:
        w = blocks.reshape(grid[0], O, grid[1], N).transpose(0, 2, 1, 3)
    elif len(grid) == 3:
        w = blocks.reshape(grid[0], O, grid[1], N, grid[2]).transpose(0, 2, 4, 1, 3)
    else:
        raise NotImplementedError("LazyESN: Wout block layout is defined for 1 to 3 space axes")
    return np.ascontiguousarray(w.reshape(-1, O, N))


SRC_STATIC = -2**31   # src entry of an input cell that no group produces (its value never changes)


def flow_tables(in_map, out_map, n_cells, world, rank):
    """Tables of the dataflow forecast kernel (csrc/predict_flow.cu) for rank `rank` of `world` ranks that own
    contiguous, equally sized slabs of the group list (`partition_groups`).  Replaces the whole-field re-halo of
    xesn/lazyesn.py:235 by neighbour-only traffic: a rank IMPORTS exactly the cells its groups read that groups of
    other ranks produce, and EXPORTS exactly the cells other ranks import from it.

    in_map (G_all, I) / out_map (G_all, O): global cell indices (in_map < 0: constant fill).  Ring slots of a rank:
    s < G*O = its local output (g, o); G*O + j = its j-th imported cell (sorted by cell index).  Returns a dict:
      src (G, I) int32        ring slot feeding input k of local group g | negative fill code | SRC_STATIC
      import_cell (n_imp,)    cell of every import slot
      import_owner (n_imp,)   rank that produces it
      exp_ptr (G*O+1,), exp_peer, exp_slot : CSR over the local slots of (importing rank, slot in its ring)
      n_import_all (world,)   number of import slots of every rank (ring sizes)
    Every rank can compute every rank's tables: no communication is needed to build them."""
    G_all, O = out_map.shape
    if G_all % world:
        raise ValueError(f"flow_tables: {G_all} groups do not divide over {world} ranks")
    per = G_all // world
    owner = np.full(n_cells, -1, dtype=np.int64)          # global slot g*O + o producing a cell
    owner[out_map.reshape(-1)] = np.arange(G_all * O)

    def imports_of(p):
        m = in_map[p * per:(p + 1) * per]
        cells = np.unique(m[m >= 0])
        own = owner[cells]
        return cells[(own >= 0) & ((own // O < p * per) | (own // O >= (p + 1) * per))]

    imports = [imports_of(p) for p in range(world)]
    mine = imports[rank]
    g0 = rank * per
    m = in_map[g0:g0 + per].astype(np.int64)
    own = owner[np.maximum(m, 0)]
    local = (own >= g0 * O) & (own < (g0 + per) * O)
    src = np.where(m < 0, m, np.where(own < 0, SRC_STATIC,
                                      np.where(local, own - g0 * O, per * O + np.searchsorted(mine, np.maximum(m, 0)))))
    entries = []          # (local slot, peer, slot in the peer's ring)
    for p in range(world):
        if p == rank:
            continue
        own_p = owner[imports[p]]
        sel = np.nonzero((own_p >= g0 * O) & (own_p < (g0 + per) * O))[0]
        entries += [(int(own_p[j] - g0 * O), p, per * O + int(j)) for j in sel]
    entries.sort()
    exp_ptr = np.zeros(per * O + 1, dtype=np.int32)
    for s_, _, _ in entries:
        exp_ptr[s_ + 1] += 1
    exp_ptr = np.cumsum(exp_ptr).astype(np.int32)
    return {
        "src": np.ascontiguousarray(src.astype(np.int32)),
        "import_cell": np.ascontiguousarray(mine.astype(np.int32)),
        "import_owner": (owner[mine] // O // per).astype(np.int32),
        "exp_ptr": exp_ptr,
        "exp_peer": np.asarray([e[1] for e in entries], dtype=np.int32),
        "exp_slot": np.asarray([e[2] for e in entries], dtype=np.int32),
        "n_import_all": np.asarray([len(i) for i in imports], dtype=np.int64),
    }
