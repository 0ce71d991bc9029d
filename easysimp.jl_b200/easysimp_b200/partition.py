"""x-slab partition of a structured grid across ranks -- the host-side statement of the
layout esimp_set_grid_structured builds on each rank (csrc/api.cu), used by the gloo tests and bench.py.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Slab:
    rank: int
    nranks: int
    ex0: int  # owned element layers [ex0, ex1)
    ex1: int
    gp0: int  # global index of local node plane 0
    npx: int  # local node planes stored (owned + ghosts)
    own_p0: int  # owned local node planes [own_p0, own_p1)
    own_p1: int
    ge0: int  # global index of local element layer 0
    nex: int  # local element layers stored
    own_e0: int
    own_e1: int
    mv_e0: int  # element layers the operator visits: owned + the redundant left interface layer
    mv_e1: int
    halo_e: int


def slab_partition(nx: int, rank: int, nranks: int, halo_e: int = 1) -> Slab:
    """Rank r owns element layers [nx*r/R, nx*(r+1)/R) and node planes ex0..ex1-1 (the last rank also
    owns plane nx).  It stores one ghost node plane per interior side and `halo_e` ghost element layers
    per interior side; the left interface layer is computed redundantly so that only u (never y) moves."""
    ex0, ex1 = nx * rank // nranks, nx * (rank + 1) // nranks
    left, right = int(rank > 0), int(rank < nranks - 1)
    he = halo_e if nranks > 1 else 0
    gl, gr = he * left, he * right
    npx = (ex1 - ex0) + 1 + left
    return Slab(rank, nranks, ex0, ex1, ex0 - left, npx, left, npx - right, ex0 - gl, (ex1 - ex0) + gl + gr, gl,
                gl + (ex1 - ex0), gl - left, gl + (ex1 - ex0), he)


def multigrid_levels(nel, nranks: int, min_layers: int = 1, max_levels: int = 12):
    """Grid hierarchy csrc/multigrid.cu builds (mg_build_levels): halve every dimension while all three are even and
    every rank keeps >= min_layers whole cell layers (coarse nx divisible by nranks).  Returns [(nx, ny, nz), ...]."""
    nx, ny, nz = nel
    levels = [(nx, ny, nz)]
    while len(levels) < max_levels:
        if nx % 2 or ny % 2 or nz % 2:
            break
        cx = nx // 2
        if cx % nranks or cx // nranks < min_layers:
            break
        nx, ny, nz = cx, ny // 2, nz // 2
        levels.append((nx, ny, nz))
    return levels


def level_slab(level_nx: int, rank: int, nranks: int) -> Slab:
    """Slab of one multigrid level (one ghost cell layer per interior side on coarse levels)."""
    return slab_partition(level_nx, rank, nranks, 1)
