"""TEST INFRASTRUCTURE ONLY — locate and import the live reference (EasyFEA) in the authoring container.

The reference needs `gmsh` only for meshing (`EasyFEA/FEM/_mesher.py:13`); an empty stub module lets the
whole package import so hand-built meshes can drive its NumPy/SciPy path.  `/root/reference` does not exist
on the GPU box: nothing under `tests/ -m gpu`, `bench.py` or `__graft_entry__.smoke()` imports this file.
It is used (a) by `tests/golden/make_golden.py` to mint fixtures and (b) by the `not gpu` tests that pin the
NumPy oracle restatement (`oracle/easyfea_oracle.py`) against the live reference when it is present.
"""
import os
import sys
import types

REF_PATHS = ("/root/reference",)


def reference_available() -> bool:
    return any(os.path.isdir(os.path.join(p, "EasyFEA")) for p in REF_PATHS)


def import_reference():
    """Returns the imported `EasyFEA` package of the reference (raises ImportError if absent)."""
    if "EasyFEA" in sys.modules:
        return sys.modules["EasyFEA"]
    for p in REF_PATHS:
        if os.path.isdir(os.path.join(p, "EasyFEA")):
            if "gmsh" not in sys.modules:
                try:
                    import gmsh  # noqa: F401
                except Exception:
                    sys.modules["gmsh"] = types.ModuleType("gmsh")
            if p not in sys.path:
                sys.path.insert(0, p)
            import EasyFEA  # noqa: F401

            return sys.modules["EasyFEA"]
    raise ImportError("reference EasyFEA not found (expected at /root/reference)")
