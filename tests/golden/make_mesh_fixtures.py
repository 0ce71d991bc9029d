"""Convert the reference's imported-mesh test inputs (data/cantilever_beam.vtu, data/Wheel_3d_coarse.msh, data/stul14.vtu, data/stul15.vtu)
into compact .npz fixtures: /root/reference does not exist on the GPU box.  Only derived arrays
(point coordinates + connectivity) are stored, no reference source.  Also extracts the one numeric OUTPUT the
reference ships: the `density` cell field of data/cantilever_beam_opt.vtu (a 60x20x4 cantilever optimised by an
earlier version of the package; mean 0.4007, minimum exactly 0), re-ordered to generate_grid's cell order.
Run from the repo root:
    python tests/golden/make_mesh_fixtures.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.mesh_readers import read_msh41, read_vtu, read_vtu_arrays  # noqa: E402

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent
pts, cells = read_vtu(REF / "cantilever_beam.vtu")
print("cantilever_beam.vtu", pts.shape, cells.shape, cells.min(), cells.max())
np.savez_compressed(OUT / "mesh_cantilever_beam_hex.npz", points=pts, cells=cells.astype(np.int32))
pts, cells = read_msh41(REF / "Wheel_3d_coarse.msh")
print("Wheel_3d_coarse.msh", pts.shape, cells.shape, cells.min(), cells.max())
np.savez_compressed(OUT / "mesh_wheel_coarse_tet.npz", points=pts, cells=cells.astype(np.int32))
for name in ("stul14", "stul15"):  # general hexes: example 04's gripper (stul14) and its refinement (stul15, the c3 bench workload)
    pts, cells = read_vtu(REF / f"{name}.vtu")
    print(f"{name}.vtu", pts.shape, cells.shape)
    np.savez_compressed(OUT / f"mesh_{name}_hex.npz", points=pts, cells=cells.astype(np.int32))
pts, cells = read_vtu(REF / "cantilever_beam_opt.vtu")
rho = read_vtu_arrays(REF / "cantilever_beam_opt.vtu")["density"].astype(np.float64)
c = pts[cells].mean(axis=1)  # unit cubes: centroid (i+.5, j+.5, k+.5)
ijk = np.floor(c).astype(np.int64)
assert ijk.min() == 0 and tuple(ijk.max(axis=0)) == (59, 19, 3) and rho.size == 4800
grid_order = np.empty(4800)
grid_order[ijk[:, 0] + 60 * (ijk[:, 1] + 20 * ijk[:, 2])] = rho
print("cantilever_beam_opt.vtu density: mean %.4f min %.3g max %.3g" % (rho.mean(), rho.min(), rho.max()))
np.savez_compressed(OUT / "ref_cantilever_beam_opt_density.npz", density=grid_order, nel=np.array([60, 20, 4]))
