"""TEST INFRASTRUCTURE ONLY — locate and import the live reference (EasyFEA).

Search order: the offline install that TRAVELS with the repository snapshot (`baseline/_ref`, made once by
`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference`, git-ignored, see
`__graft_entry__.build()`), then `/root/reference` (authoring container only), then an `EasyFEA` already importable.
The reference needs `gmsh` only for meshing (`EasyFEA/FEM/_mesher.py:13`); an empty stub module lets the whole package
import so hand-built meshes can drive its NumPy/SciPy path.

Users: `tests/` (oracle pinning, the GPU drop-in tests against the live reference), `tests/golden/make_golden*.py`, and
the CPU arm of `bench.py` (`--impl reference` / `cpu_baseline`).  Nothing under `easyfea_b200/` imports this file.
"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAVEL_PATH = os.path.join(ROOT, "baseline", "_ref")
REF_PATHS = (TRAVEL_PATH, "/root/reference")


def _candidates(travel_only: bool):
    return (TRAVEL_PATH,) if travel_only else REF_PATHS


def reference_available(travel_only: bool = False) -> bool:
    if "EasyFEA" in sys.modules:
        return True
    if any(os.path.isdir(os.path.join(p, "EasyFEA")) for p in _candidates(travel_only)):
        return True
    return not travel_only and importlib.util.find_spec("EasyFEA") is not None


def reference_location(travel_only: bool = False):
    """Directory the reference would be imported from (None: an installed package or nothing)."""
    for p in _candidates(travel_only):
        if os.path.isdir(os.path.join(p, "EasyFEA")):
            return p
    return None


def import_reference(travel_only: bool = False):
    """Returns the imported `EasyFEA` package of the reference (raises ImportError if absent).
    `travel_only`: look at `baseline/_ref` alone — what the `-m gpu` tests and `bench.py` do, because `/root/reference`
    does not exist on the GPU box."""
    if "EasyFEA" in sys.modules:
        return sys.modules["EasyFEA"]
    if "gmsh" not in sys.modules:
        try:
            import gmsh  # noqa: F401
        except Exception:
            sys.modules["gmsh"] = types.ModuleType("gmsh")
    p = reference_location(travel_only)
    if p is not None:
        if p not in sys.path:
            sys.path.insert(0, p)
    elif travel_only or importlib.util.find_spec("EasyFEA") is None:
        raise ImportError("reference EasyFEA not found (looked at baseline/_ref" + ("" if travel_only else ", /root/reference, site-packages") + ")")
    import EasyFEA  # noqa: F401

    return sys.modules["EasyFEA"]
