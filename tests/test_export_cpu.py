"""Sharded result export: per-rank iteration files and ParaView pieces written from a partition
reassemble to the global fields; files parse as VTK XML."""
import base64
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from easyfea_b200 import dist as efd
from easyfea_b200 import export
from tests.helpers import make_mesh


def _decode(node):
    raw = base64.b64decode(node.text.strip())
    n = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
    dt = {"Float64": np.float64, "Int64": np.int64, "UInt8": np.uint8, "Int32": np.int32, "Float32": np.float32}[node.get("type")]
    return np.frombuffer(raw[8:8 + n], dtype=dt)


@pytest.mark.parametrize("elemType,n,world", [("HEXA8", (4, 3, 3), 3), ("TETRA10", (2, 2, 2), 2), ("TRI3", (6, 5), 4)])
def test_sharded_iteration_files_and_vtu_pieces(tmp_path, elemType, n, world):
    coords, connect = make_mesh(elemType, n)
    Nn = coords.shape[0]
    dim = 2 if elemType.startswith(("TRI", "QUAD")) else 3
    rng = np.random.default_rng(0)
    u = rng.normal(size=Nn * dim)
    d = rng.uniform(size=Nn)
    erank = efd.rcb_element_ranks(coords[connect].mean(1), world)
    pvtu = None
    seen_cells = 0
    for r in range(world):
        part = efd.Partition.from_global(connect, Nn, world, r, erank=erank)
        ul = u.reshape(Nn, dim)[part.nodes].ravel()  # local [owned | halo] vectors, as the drivers hold them
        dl = d[part.nodes]
        export.save_iter(str(tmp_path), 3, {"displacement": ul, "damage": dl}, part, scalars={"Niter": 7})
        svm = rng.uniform(size=part.elem_ids.size)
        f = export.save_vtu(str(tmp_path), "simu", elemType, coords[part.nodes], part.connect, {"u": ul, "d": dl}, {"Svm": svm}, part, 3)
        pvtu = f
        root = ET.parse(os.path.join(tmp_path, f"simu_3_rank{r}.vtu")).getroot()
        piece = root.find("UnstructuredGrid/Piece")
        assert int(piece.get("NumberOfPoints")) == part.n_local and int(piece.get("NumberOfCells")) == part.elem_ids.size
        arrays = {a.get("Name"): a for a in piece.iter("DataArray")}
        pts = _decode(arrays["Points"]).reshape(-1, 3)
        assert np.allclose(pts[:, :dim], coords[part.nodes][:, :dim])
        conn = _decode(arrays["connectivity"]).reshape(-1, connect.shape[1])
        order = export.GMSH_TO_VTK.get(elemType, list(range(connect.shape[1])))
        assert np.array_equal(part.nodes[conn], connect[part.elem_ids][:, order])
        ghosts = [a for a in piece.iter("DataArray") if a.get("Name") == "vtkGhostType"]
        pg, cg = _decode(ghosts[0]), _decode(ghosts[1])
        assert pg.sum() == part.n_halo
        seen_cells += int((cg == 0).sum())
        uu = _decode(arrays["u"]).reshape(-1, 3)
        assert np.array_equal(uu[:, :dim].ravel(), ul)
    assert seen_cells == connect.shape[0]  # every element is a non-ghost cell of exactly one piece
    proot = ET.parse(os.path.join(tmp_path, pvtu)).getroot()
    assert len(proot.findall("PUnstructuredGrid/Piece")) == world
    got = export.load_iter(str(tmp_path), 3, Nn, world)
    assert np.array_equal(got["displacement"], u) and np.array_equal(got["damage"], d) and int(got["scalars"]["Niter"]) == 7
    export.save_pvd(str(tmp_path), "simulation", [pvtu], [0.5])
    assert ET.parse(os.path.join(tmp_path, "simulation.pvd")).getroot().find("Collection/DataSet").get("file") == pvtu


def test_single_rank_iteration_roundtrip(tmp_path):
    coords, connect = make_mesh("QUAD9", (3, 2))
    Nn = coords.shape[0]
    u = np.arange(Nn * 2, dtype=float)
    export.save_iter(str(tmp_path), 0, {"displacement": u})
    assert np.array_equal(export.load_iter(str(tmp_path), 0, Nn)["displacement"], u)
    f = export.save_vtu(str(tmp_path), "m", "QUAD9", coords, connect, {"u": u})
    assert f == "m_0.vtu" and ET.parse(os.path.join(tmp_path, f)).getroot().tag == "VTKFile"
