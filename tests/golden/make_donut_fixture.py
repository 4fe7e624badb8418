"""Extract the reference's only mesh fixture (src/mpet/test/donut2D.h5, used by
src/mpet/test/test_donut.py:36-40) into a numpy archive that travels with the repo.

Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_donut_fixture.py
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.mesh import read_donut_h5  # noqa: E402

SRC = "/root/reference/src/mpet/test/donut2D.h5"

if __name__ == "__main__":
    mesh = read_donut_h5(SRC)
    out = os.path.join(os.path.dirname(__file__), "donut2D.npz")
    np.savez_compressed(out, coordinates=mesh.coords, topology=mesh.cells)
    print("wrote", out, mesh.coords.shape, mesh.cells.shape)
