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
