"""Convert the reference's imported-mesh test inputs (data/cantilever_beam.vtu, data/Wheel_3d_coarse.msh)
into compact .npz fixtures: /root/reference does not exist on the GPU box.  Only derived arrays
(point coordinates + connectivity) are stored, no reference source.  Run from the repo root:
    python tests/golden/make_mesh_fixtures.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.mesh_readers import read_msh41, read_vtu  # noqa: E402

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent
pts, cells = read_vtu(REF / "cantilever_beam.vtu")
print("cantilever_beam.vtu", pts.shape, cells.shape, cells.min(), cells.max())
np.savez_compressed(OUT / "mesh_cantilever_beam_hex.npz", points=pts, cells=cells.astype(np.int32))
pts, cells = read_msh41(REF / "Wheel_3d_coarse.msh")
print("Wheel_3d_coarse.msh", pts.shape, cells.shape, cells.min(), cells.max())
np.savez_compressed(OUT / "mesh_wheel_coarse_tet.npz", points=pts, cells=cells.astype(np.int32))
