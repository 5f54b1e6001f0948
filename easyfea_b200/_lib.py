"""ctypes binding of the C-ABI shared library `libeasyfea_b200.so` (include/easyfea_b200.h).

There is NO fallback: if the library is missing or no CUDA device is visible, every compute entry point raises.
Build it in-tree with `make` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EASYFEA_B200_LIB: development override used by scripts/ to time alternative builds of the same C ABI
LIB_PATH = os.environ.get("EASYFEA_B200_LIB") or os.path.join(_HERE, "libeasyfea_b200.so")

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class EfbGroup(ctypes.Structure):
    """`efb_group` of include/easyfea_b200.h"""

    _fields_ = [("dim", c_i32), ("nPe", c_i32), ("nPg", c_i32), ("coord_stride", c_i32), ("Ne", c_i64), ("connect", c_vp),
                ("coord", c_vp), ("dN_pg", c_vp), ("N_pg", c_vp), ("w_pg", c_vp)]


class EfbPfMaterial(ctypes.Structure):
    """`efb_pf_material` of include/easyfea_b200.h"""

    _fields_ = [("dim", c_i32), ("split", c_i32), ("planeStress", c_i32), ("_pad", c_i32), ("E", c_f64), ("v", c_f64),
                ("lam", c_f64), ("mu", c_f64), ("bulk", c_f64), ("C", c_f64 * 36), ("sqrtC", c_f64 * 36),
                ("inv_sqrtC", c_f64 * 36)]


MAX_RANKS = 8  # EFB_MAX_RANKS


class EfbPcgSystem(ctypes.Structure):
    """`efb_pcg_system` of include/easyfea_b200.h"""

    _fields_ = [("nrows", c_i64), ("kind", c_i32), ("index_bytes", c_i32), ("dof_n", c_i32), ("lanes", c_i32), ("indptr", c_vp),
                ("indices", c_vp), ("data", c_vp), ("free_mask", c_vp), ("inv_diag", c_vp), ("x", c_vp), ("r", c_vp), ("z", c_vp),
                ("Ap", c_vp), ("partials", c_vp), ("s", c_vp)]


class EfbPcgPeer(ctypes.Structure):
    """`efb_pcg_peer` of include/easyfea_b200.h"""

    _fields_ = [("world", c_i32), ("rank", c_i32), ("n_send", c_i32), ("n_recv", c_i32), ("send_rank", c_i32 * MAX_RANKS),
                ("recv_rank", c_i32 * MAX_RANKS), ("send_ptr", c_i64 * (MAX_RANKS + 1)), ("send_dst", c_i64 * MAX_RANKS),
                ("base", c_vp * MAX_RANKS), ("pbuf_off", (c_i64 * 4) * MAX_RANKS), ("send_idx", c_vp), ("ar_seq", ctypes.c_uint64),
                ("halo_seq", ctypes.c_uint64), ("push_id", c_vp), ("push_ptr", c_vp), ("push_nbr", c_vp), ("push_pos", c_vp)]


_GP = ctypes.POINTER(EfbGroup)
_PP = ctypes.POINTER(c_vp)
_I64P = ctypes.POINTER(c_i64)
_I32P = ctypes.POINTER(c_i32)

# name -> argtypes; every function returns int status except the ones listed in _RESTYPES
SIGNATURES = {
    "efb_geometry": [_GP, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_geometry_parts": [_GP, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_elastic_Ke": [_GP, c_vp, c_vp, ctypes.c_int, c_f64, c_vp, c_vp],
    "efb_mass_Me": [_GP, c_vp, ctypes.c_int, c_f64, ctypes.c_int, c_f64, c_vp, c_vp],
    "efb_diffusion_Ke": [_GP, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_f64, c_f64, c_vp, c_vp],
    "efb_source_Fe": [_GP, c_vp, ctypes.c_int, c_f64, ctypes.c_int, c_f64, c_vp, c_vp],
    "efb_internal_force": [_GP, c_vp, c_vp, c_vp],
    "efb_strain": [_GP, c_vp, c_vp, c_vp, c_vp],
    "efb_hyperelastic_Ke_Re": [_GP, c_vp, c_vp, c_vp, c_vp, c_f64, c_vp, c_vp, c_vp],
    "efb_pf_split": [ctypes.POINTER(EfbPfMaterial), c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_pf_elastic_Ke": [ctypes.POINTER(EfbPfMaterial), _GP, c_vp, c_vp, c_vp, c_f64, c_f64, c_vp, c_vp],
    "efb_pf_degradation": [_GP, c_vp, c_vp, c_f64, c_vp, c_vp],
    "efb_pf_history_rf": [c_vp, c_vp, c_i64, ctypes.c_int, c_f64, c_f64, c_vp, c_vp, c_vp],
    "efb_pf_damage_Ke_Fe": [_GP, c_vp, c_vp, c_vp, c_f64, c_f64, c_vp, c_vp, c_vp],
    "efb_csr_count_node_rows": [ctypes.c_int, _PP, _I64P, _I32P, c_i64, c_vp, c_vp],
    "efb_exclusive_scan_i32": [c_vp, c_i64, c_vp, c_vp, c_vp],
    "efb_csr_fill_node_rows": [ctypes.c_int, _PP, _I64P, _I32P, c_i64, c_vp, c_vp, c_vp, c_vp],
    "efb_csr_count_adj": [ctypes.c_int, _PP, _I64P, _I32P, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_csr_fill_adj": [ctypes.c_int, _PP, _I64P, _I32P, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_csr_expand": [c_i64, ctypes.c_int, c_i64, c_i64, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp],
    "efb_csr_slot_map": [ctypes.c_int, _PP, _I64P, _I32P, c_vp, c_vp, c_vp, c_vp],
    "efb_csr_inv_map": [ctypes.c_int, _PP, _I64P, _I32P, ctypes.c_int, c_vp, c_vp, c_vp, c_vp],
    "efb_csr_row_has_entry": [c_i64, ctypes.c_int, c_i64, c_vp, c_vp, c_vp],
    "efb_csr_compact_rows": [c_i64, c_vp, c_vp, c_vp, c_vp],
    "efb_csr_replay_matrix": [ctypes.c_int, _PP, _I64P, _I32P, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp],
    "efb_csr_replay_vector": [ctypes.c_int, _PP, _I64P, _I32P, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp],
    "efb_assemble_elastic": [_GP, c_vp, c_vp, c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                             c_vp, c_vp],
    "efb_assemble_elastic_smem": [ctypes.c_int] * 6,
    "efb_assemble_elastic_group": [ctypes.c_int] * 3,
    "efb_assemble_elastic_mma": [_GP, c_vp, c_vp, c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_assemble_elastic_mma_smem": [ctypes.c_int] * 3,
    "efb_spmv_csr": [c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, ctypes.c_int, c_vp],
    "efb_spmv_nodeblock": [c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, ctypes.c_int, c_vp],
    "efb_csr_diagonal": [c_i64, c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_pcg_inv_diag": [c_i64, c_vp, c_vp, c_vp, c_vp],
    "efb_pcg_dot": [c_i64, c_vp, c_vp, c_vp, c_vp],
    "efb_pcg_reduce": [c_vp, ctypes.c_int, c_vp, c_vp],
    "efb_pcg_init": [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_pcg_update_xr": [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_pcg_update_p": [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "efb_pack_f64": [c_i64, c_vp, c_vp, c_vp, c_vp],
    "efb_hooke": [c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp, c_vp],
    "efb_field_result": [c_i64, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, c_f64, c_vp, c_vp],
    "efb_energy_e": [c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_f64, c_vp, c_vp],
    "efb_node_values": [c_i64, c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp, c_vp],
    "efb_lincomb": [c_i64, ctypes.c_int, ctypes.POINTER(c_f64), _PP, c_vp, c_vp],
    "efb_pcg_ctrl_bytes": [],
    "efb_pcg_ctrl_layout": [_I32P],
    "efb_pcg_iterate": [ctypes.POINTER(EfbPcgSystem), ctypes.POINTER(EfbPcgPeer), ctypes.c_int, c_i64, c_vp],
    "efb_pcg_iterate_cheb": [ctypes.POINTER(EfbPcgSystem), ctypes.POINTER(EfbPcgPeer), ctypes.c_int, c_i64, ctypes.c_int, c_f64, c_f64,
                             c_vp, c_vp, ctypes.c_int, c_vp],
    "efb_cast_f32": [c_i64, c_vp, c_vp, c_vp],
    "efb_pcg_cheb_update": [c_i64, c_vp, c_vp, c_vp, c_f64, c_f64, c_vp, c_vp, c_vp],
    "efb_pcg_iterate_cg2": [ctypes.POINTER(EfbPcgSystem), ctypes.POINTER(EfbPcgPeer), ctypes.c_int, c_i64, c_vp],
    "efb_pcg_solve_persistent": [ctypes.POINTER(EfbPcgSystem), ctypes.POINTER(EfbPcgPeer), c_i64, c_i64, c_f64, c_vp],
    "efb_peer_alloc": [c_i64, _PP],
    "efb_peer_free": [c_vp],
    "efb_peer_export": [c_vp, c_vp],
    "efb_peer_open": [c_vp, _PP],
    "efb_peer_close": [c_vp],
    "efb_pcg_partials_size": [],
    "efb_version": [],
    "efb_device_count": [],
    "efb_last_error": [],
}
_RESTYPES = {"efb_last_error": ctypes.c_char_p}

_lib = None


class EfbError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises EfbError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EfbError(f"{LIB_PATH} not found: build the CUDA library first (`make` or __graft_entry__.build()); "
                           "easyfea_b200 has no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = lib
    return _lib


def call(name: str, *args):
    """Call an int-status entry point; non-zero status -> EfbError carrying efb_last_error()."""
    lib = load()
    status = getattr(lib, name)(*args)
    if status != 0:
        raise EfbError(f"{name} failed ({status}): {lib.efb_last_error().decode()}")


def require_cuda():
    """The product path needs a CUDA device: fail loudly otherwise (no silent CPU path)."""
    import torch

    if not torch.cuda.is_available():
        raise EfbError("easyfea_b200 needs a CUDA device (torch.cuda.is_available() is False); there is no CPU fallback")
    lib = load()
    if lib.efb_device_count() < 1:
        raise EfbError(f"libeasyfea_b200.so sees no CUDA device: {lib.efb_last_error().decode()}")
