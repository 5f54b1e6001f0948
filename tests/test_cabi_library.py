"""The C-ABI shared library loads without a GPU and exports every symbol include/easyfea_b200.h declares; host-side
argument logic (broadcast modes, scipy index-dtype rule) behaves like the reference.  No compute calls here."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "easyfea_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(efb_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from easyfea_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/easyfea_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)
    assert lib.efb_version() == 100


def test_no_cpu_fallback():
    import torch

    from easyfea_b200 import _lib, mesh, meshgen, operators

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    coords, connect = meshgen.structured_mesh("HEXA8", 2)
    g = mesh.ElemGroup("HEXA8", connect, coords)
    with pytest.raises(_lib.EfbError):
        operators.LinearizedElasticity(g, np.eye(6))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "easyfea_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), fn


def test_broadcast_rules_match_reference_shapes():
    from easyfea_b200.operators import coef_mode, tensor_mode

    Ne, nPg = 7, 4
    assert coef_mode(3, Ne, nPg) == (None, 0, 3.0)
    assert coef_mode(np.float64(2.5), Ne, nPg)[1:] == (0, 2.5)
    assert coef_mode(np.ones(Ne), Ne, nPg)[1] == 1
    assert coef_mode(np.ones(nPg), Ne, nPg)[1] == 2
    assert coef_mode(np.ones((Ne, nPg)), Ne, nPg)[1] == 3
    assert coef_mode(np.ones(4), 4, 4)[1] == 1  # ambiguous 1-D: (Ne,) wins, _linalg.py:469-473
    with pytest.raises(ValueError):
        coef_mode(np.ones(5), Ne, nPg)
    assert tensor_mode(np.ones((6, 6)), Ne, nPg, 6)[1] == 0
    assert tensor_mode(np.ones((Ne, 6, 6)), Ne, nPg, 6)[1] == 1
    assert tensor_mode(np.ones((Ne, nPg, 6, 6)), Ne, nPg, 6)[1] == 2
    with pytest.raises(ValueError):
        tensor_mode(np.ones((nPg, 6, 6)), Ne, nPg, 6)  # (nPg,) lead is rejected with tensor_ndim=2, _linalg.py:462-465


def test_elem_group_matches_reference_conventions():
    from easyfea_b200 import mesh, meshgen
    from oracle import easyfea_oracle as orc

    coords, connect = meshgen.structured_mesh("QUAD9", (3, 2))
    g = mesh.ElemGroup("QUAD9", connect, coords)
    for dof_n in (1, 2):
        assert np.array_equal(g.Get_assembly_e(dof_n), orc.assembly_e(connect, dof_n))
    u = np.arange(coords.shape[0] * 2.0)
    assert np.array_equal(g.Locates_sol_e(u), orc.locate_sol_e(u, connect, 2))
