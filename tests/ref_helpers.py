"""Reference-side case builders of the tests: see oracle/ref_cases.py (shared with the reference legs of bench.py)."""
from oracle.ref_cases import (apply_shear, hexa8_elastic, phasefield_case, quad9_boundary_seg3, readme_cantilever,  # noqa: F401
                              ref_mesh, shear_sets)
