"""The drop-in installer rebinds exactly the reference attributes INTEGRATION.md lists and restores them; without a GPU the
patched entry points fail loudly instead of computing on the CPU.  Needs the live reference (skipped on the GPU box)."""
import numpy as np
import pytest

from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="live reference (/root/reference) not present")


def test_install_uninstall_roundtrip():
    import torch

    EasyFEA = import_reference()
    from EasyFEA.FEM import Operators
    from EasyFEA.Simulations._simu import _Simu

    from easyfea_b200 import _lib, dropin, operators

    before = {(m, n): getattr(getattr(Operators, m), n) for m, names in dropin._LEVEL1.items() for n in names}
    csr_before = (_Simu.__dict__["_Simu__Get_csr_map"], _Simu.__dict__["_Simu__Assemble_csr"])
    pf_before = EasyFEA.Models.PhaseField.__dict__["Calc_C"]
    patched = dropin.install(EasyFEA)
    try:
        assert dropin.installed() and len(patched) == 6 + 2 + 3
        assert Operators.Bilinear.LinearizedElasticity is operators.LinearizedElasticity
        assert Operators.Linear.V is operators.V
        assert _Simu.__dict__["_Simu__Assemble_csr"] is not csr_before[1]
        with pytest.raises(RuntimeError):
            dropin.install(EasyFEA)
        if not torch.cuda.is_available():
            # the reference's own simulation now routes through the device path, which must refuse to run on the CPU
            from EasyFEA import Models, Simulations
            from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh

            from easyfea_b200 import meshgen

            coords, connect = meshgen.structured_mesh("HEXA8", 2)
            g = GroupElemFactory.Create(ElemType.HEXA8, connect, coords)
            simu = Simulations.Elastic(Mesh({ElemType.HEXA8: g}), Models.Elastic.Isotropic(3))
            with pytest.raises(_lib.EfbError):
                simu.Get_K_C_M_F()
    finally:
        dropin.uninstall()
    assert not dropin.installed()
    for (m, n), f in before.items():
        assert getattr(getattr(Operators, m), n) is f
    assert (_Simu.__dict__["_Simu__Get_csr_map"], _Simu.__dict__["_Simu__Assemble_csr"]) == csr_before
    assert EasyFEA.Models.PhaseField.__dict__["Calc_C"] is pf_before
    # and the unpatched reference still assembles
    from EasyFEA import Models, Simulations
    from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh

    from easyfea_b200 import meshgen

    coords, connect = meshgen.structured_mesh("HEXA8", 2)
    g = GroupElemFactory.Create(ElemType.HEXA8, connect, coords)
    K = Simulations.Elastic(Mesh({ElemType.HEXA8: g}), Models.Elastic.Isotropic(3)).Get_K_C_M_F()[0]
    assert K.shape == (81, 81) and np.isfinite(K.data).all()
