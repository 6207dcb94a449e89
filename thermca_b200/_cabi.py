"""ctypes binding of include/thermca_b200.h.

The shared library is built in-tree (``python -m thermca_b200.build``).  There is no
CPU fallback: if the library is missing or no CUDA device is present, the product path
raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libthermca_b200.so")

i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class NetDesc(C.Structure):
    _fields_ = [("iarena", vp), ("darena", vp), ("cmds_off", i32), ("prog_dir_off", i32),
                ("tab_dir_off", i32), ("N", i32), ("u", i32), ("nnz", i32)] + \
               [(n, i32) for n in ("o_T", "o_heat", "o_capy", "o_vol", "o_volcapy", "o_condy", "o_cond",
                                   "o_clean", "o_calc", "o_inputs", "o_indptr", "o_indices", "o_diag")] + \
               [("n_mean_jobs", i32), ("n_mean_blocks", i32), ("job_of_block", vp), ("first_block", vp),
                ("nblk", vp), ("ent_start", vp), ("ent_idx", vp), ("ent_w", vp), ("n_iarena", i32), ("n_darena", i32)]


class FfeDesc(C.Structure):
    _fields_ = [("n", i32), ("nslices", i32), ("y_off", i64), ("slice_ptr", vp), ("col", vp), ("col16", vp), ("val", vp),
                ("mass", vp), ("adiag", vp),
                ("load_nrows", i32), ("load_row", vp), ("load_start", vp), ("load_surf", vp), ("load_bval", vp),
                ("fd_nrows", i32), ("fd_pos", vp), ("fd_row", vp), ("fd_abody", vp), ("fd_start", vp),
                ("fd_film", vp), ("fd_adata", vp), ("films_doff", i32), ("flux_doff", i32),
                ("tdep", i32), ("condy_tab", i32), ("volcapy_tab", i32), ("ts_nrows", i32),
                ("ts_row", vp), ("ts_start", vp), ("ts_surf", vp), ("ts_qval", vp),
                ("ts_fstart", vp), ("ts_film", vp), ("ts_adata", vp),
                ("j_nslices", i32), ("j_slice_ptr", vp), ("j_col", vp), ("j_val", vp), ("j_mass", vp), ("j_adiag", vp),
                ("fd_rhs", i32)]


class LpDesc(C.Structure):
    _fields_ = [("n", i32), ("nslices", i32), ("y_off", i64), ("slice_ptr_ioff64", i32), ("col_ioff", i32),
                ("val_doff", i32), ("b_doff", i32), ("mass_doff", i32), ("adiag_doff", i32),
                ("F", i32), ("lfm_doff", i32), ("minv_doff", i32)]


class MorDesc(C.Structure):
    _fields_ = [("S", i32), ("r", i32), ("F", i32), ("y_off", i64), ("A_body", vp), ("A_films", vp),
                ("B_T", vp), ("films_doff", i32), ("flux_doff", i32)]


EXPORTS = {
    "tc_last_error": (C.c_char_p, []),
    "tc_version": (i32, []),
    "tc_create": (i32, [C.POINTER(vp), vp]),
    "tc_destroy": (i32, [vp]),
    "tc_set_network": (i32, [vp, C.POINTER(NetDesc)]),
    "tc_add_ffe_part": (i32, [vp, C.POINTER(FfeDesc)]),
    "tc_add_lp_part": (i32, [vp, C.POINTER(LpDesc)]),
    "tc_add_mor_part": (i32, [vp, C.POINTER(MorDesc)]),
    "tc_finalize": (i32, [vp, i64]),
    "tc_set_jac_films": (i32, [vp, i32, vp]),
    "tc_rhs": (i32, [vp, f64, vp, vp]),
    "tc_set_state": (i32, [vp, vp]),
    "tc_get_state": (i32, [vp, vp]),
    "tc_get_state_async": (i32, [vp, vp]),
    "tc_set_history": (i32, [vp, vp, vp, i32, i32]),
    "tc_set_probes": (i32, [vp, vp, i32]),
    "tc_rk_begin": (i32, [vp, i32, f64, f64, f64, f64, f64, f64]),
    "tc_rk_advance": (i32, [vp, i32, i32]),
    "tc_get_info": (i32, [vp, C.POINTER(i64), C.POINTER(f64), C.POINTER(f64)]),
    "tc_reset_history": (i32, [vp]),
    "tc_graph_info": (i32, [vp, vp]),
    "tc_theta_steps": (i32, [vp, i32, f64, f64, f64, i32, i32]),
    "tc_ros2_begin": (i32, [vp, f64, f64, f64, f64, f64, f64, f64, i32, i32]),
    "tc_ros2_advance": (i32, [vp, i32, i32]),
    "tc_row3_begin": (i32, [vp, f64, f64, f64, f64, f64, f64, f64, i32, i32]),
    "tc_row3_advance": (i32, [vp, i32, i32]),
    "tc_set_time": (i32, [vp, f64]),
    "tc_timing": (i32, [vp, i32, C.POINTER(f64), C.POINTER(i64), C.POINTER(i64)]),
    "tc_pcg_profile": (i32, [vp, C.POINTER(C.c_uint64)]),
    "tc_pcg_status": (i32, [vp, C.POINTER(i64)]),
    "tc_pcg_stats": (i32, [vp, C.POINTER(i64)]),
    "tc_set_block_precond": (i32, [vp, i32, vp, f64]),
    "tc_sell_spmv": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, f64, i32]),
    "tc_vm_eval": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]),
    "tc_pcg_solve": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, f64, f64, i32, C.POINTER(i32)]),
    "tc_pcg_solve_bj": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, f64, f64, i32, C.POINTER(i32), vp]),
    "tc_bar_bench": (i32, [vp, i32, i32, i32, i32, i32, C.POINTER(f64)]),
    "tc_dist_create": (i32, [C.POINTER(vp), vp, i32, i32, i64, i64]),
    "tc_dist_ipc_handle": (i32, [vp, vp]),
    "tc_dist_open_peers": (i32, [vp, vp, C.POINTER(i64)]),
    "tc_dist_open_peers_local": (i32, [vp, C.POINTER(vp), C.POINTER(i64)]),
    "tc_dist_set_halo": (i32, [vp, i64, vp, vp, vp, i32, vp]),
    "tc_dist_set_operator": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]),
    "tc_dist_set_load": (i32, [vp, vp]),
    "tc_dist_set_state_ptr": (i32, [vp, vp]),
    "tc_dist_mean_exchange": (i32, [vp, vp, vp, vp, i32, i32]),
    "tc_dist_local_rows": (i64, [vp]),
    "tc_dist_residual": (i32, [vp, vp, vp, vp]),
    "tc_dist_sum_exchange": (i32, [vp, vp, i32, vp]),
    "tc_dist_rank": (i32, [vp]),
    "tc_dist_solve": (i32, [vp, f64, vp, vp, vp, f64, i32, vp, vp]),
    "tc_dist_max_exchange": (i32, [vp, vp, i32, vp]),
    "tc_set_dist_global_rows": (i32, [vp, i64]),
    "tc_dist_stream": (vp, [vp]),
    "tc_attach_dist": (i32, [vp, i32, vp, i32, i32]),
    "tc_dist_set_state": (i32, [vp, vp]),
    "tc_dist_get_state": (i32, [vp, vp]),
    "tc_dist_copy_state_async": (i32, [vp, vp, i32]),
    "tc_dist_theta_steps": (i32, [vp, i32, f64, f64, f64, i32]),
    "tc_dist_set_pcg": (i32, [vp, i32]),
    "tc_dist_get_pcg": (i32, [vp]),
    "tc_dist_check": (i32, [vp]),
    "tc_dist_reset_error": (i32, [vp]),
    "tc_dist_stats": (i32, [vp, C.POINTER(i64)]),
    "tc_dist_profile": (i32, [vp, C.POINTER(C.c_uint64)]),
    "tc_dist_block_profile": (i32, [vp, C.POINTER(C.c_uint64), i32]),
    "tc_dist_close_peers": (i32, [vp]),
    "tc_dist_destroy": (i32, [vp]),
    "tc_mor_batch_steps": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                 f64, f64, i32, i32]),
    "tc_mor_batch_work_doubles": (i64, [i32, i32, i32, i32]),
    "tc_mor_fused_supported": (i32, [i32, i32, i32, i32]),
    "tc_mor_fused_packed_doubles": (i64, [i32, i32, i32]),
    "tc_mor_fused_work_bytes": (i64, [i32, i32, i32, i32]),
    "tc_mor_fused_pack": (i32, [vp, i32, i32, i32, vp, vp, vp]),
    "tc_mor_fused_steps": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, f64, i32]),
    "tc_synth_cuboid_counts": (i32, [i32, i32, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "tc_synth_cuboid": (i32, [vp, i32, i32, i32, f64, f64, f64, f64, f64, f64, i64, i64, vp, vp, vp, vp, vp,
                              vp, vp, vp]),
    "tc_p1_tet_contributions": (i32, [vp, i64, i64, vp, vp, vp, vp, vp, vp]),
    "tc_p1_tri_loads": (i32, [vp, i64, vp, vp, vp, vp]),
    "tc_segment_sum": (i32, [vp, i64, vp, vp, vp]),
    "tc_csr_to_sell": (i32, [vp, i32, i32, vp, vp, vp, vp, f64, vp, vp, vp, vp, vp]),
    "tc_pcg_multi_work_doubles": (i64, [i64, i32]),
    "tc_pcg_solve_multi": (i32, [vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, f64, i32, vp]),
    "tc_krylov_work_doubles": (i64, [i32]),
    "tc_krylov_orthogonalize": (i32, [vp, i64, i32, vp, i64, vp, i32, vp, vp, vp]),
    "tc_krylov_normalize": (i32, [vp, i64, vp, vp, vp]),
}

_lib = None


def load_library():
    """Load libthermca_b200.so and declare every export of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m thermca_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)         # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TcError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load_library().tc_last_error()
        raise TcError(f"thermca_b200 native call failed ({rc}): {msg.decode() if msg else '?'}")
