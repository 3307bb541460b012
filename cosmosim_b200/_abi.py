"""ctypes binding of include/cosmosim_b200.h and include/cosmosim_b200_state3d.h.  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

ABI_VERSION = 2
PAD = 1024
STATUS_LEN = 16

# flag / option / status constants (mirror the header)
F_IMMOBILE, F_MASS_INT = 1, 2
OPT_COLLISIONS, OPT_BOUNDED, OPT_FP32, OPT_STORE_ACC, OPT_F32_SYM, OPT_DEFER_INTEGRATE, OPT_ENERGY, OPT_F64_SYM, OPT_PEER_EXCHANGE = 1, 2, 4, 8, 32, 64, 128, 256, 512
ST_N_ACTIVE, ST_STEP, ST_N_EVENTS, ST_N_PAIRS, ST_N_DEAD, ST_ERROR, ST_N_ESCAPED, ST_N_CAND, ST_N_GIANTS, ST_PMAX2, ST_N_COMPLEX, ST_SLOT_LO, ST_SLOT_HI, ST_EV_BASE, ST_N_HEAVY = range(15)
SCAN_BLOCKS = 8192
ERR_PAIR_OVERFLOW, ERR_EVENT_OVERFLOW, ERR_NAN, ERR_GIANT_OVERFLOW = 1, 2, 4, 8

_vp = C.c_void_p


class State(C.Structure):
    _fields_ = [("cap", C.c_int32), ("n_total", C.c_int32),
                ("pos", _vp), ("vel", _vp), ("mass", _vp), ("mass_hi", _vp), ("mass_lo", _vp),
                ("radius", _vp), ("volume", _vp), ("eaten", _vp), ("id", _vp), ("flags", _vp),
                ("status", _vp), ("energy", _vp), ("totals", _vp),
                ("rec_mass", _vp), ("rec_mass_hi", _vp), ("rec_mass_lo", _vp), ("rec_radius", _vp),
                ("rec_volume", _vp), ("rec_eaten", _vp), ("rec_death", _vp), ("rec_flags", _vp),
                ("events", _vp), ("event_cap", C.c_int64)]


class Scratch(C.Structure):
    _fields_ = [("h", _vp), ("pk_x", _vp), ("pk_y", _vp), ("pk_m", _vp), ("acc", _vp),
                ("pos_new", _vp), ("vel_new", _vp), ("xx_new", _vp), ("rinf", _vp),
                ("partial", _vp), ("jsplit_max", C.c_int32), ("colbuf", _vp), ("colbuf_rows", C.c_int64), ("tile_ctr", _vp),
                ("alive", _vp), ("escaped", _vp), ("pairs", _vp), ("pairs_sorted", _vp),
                ("pair_cap", C.c_int64), ("cur_pos", _vp), ("cur_vel", _vp), ("mod_row", _vp),
                ("mod_step", _vp), ("chunk_alive", _vp),
                ("row_count", _vp), ("row_start", _vp), ("pmin", _vp), ("pmax", _vp), ("scan_blk", _vp),
                ("pair_flag", _vp), ("cplx_list", _vp), ("scan_a", _vp), ("ev_tmp", _vp), ("label", _vp), ("comp_items", _vp), ("comp_root", _vp),
                ("t_pos", _vp), ("t_vel", _vp), ("t_mass", _vp), ("t_mass_hi", _vp),
                ("t_mass_lo", _vp), ("t_radius", _vp), ("t_volume", _vp), ("t_eaten", _vp),
                ("t_id", _vp), ("t_flags", _vp), ("t_energy", _vp),
                ("cell_of", _vp), ("cell_count", _vp), ("cell_start", _vp), ("cell_items", _vp),
                ("giants", _vp), ("grid_cells", C.c_int32), ("giant_cap", C.c_int32),
                ("peer_acc", _vp * 16), ("n_peers", C.c_int32), ("reserved_scratch", C.c_int32),
                ("grid_dev", _vp), ("tune_hist", _vp), ("corr", _vp)]


class Grid(C.Structure):
    """cosmo_grid: the device-resident broad-phase grid of the current frame."""
    _fields_ = [("h", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("rg", C.c_double),
                ("ncell", C.c_int64), ("c", C.c_int32), ("w", C.c_int32), ("source", C.c_int32),
                ("reserved", C.c_int32), ("box", C.c_double), ("h_tuned", C.c_double),
                ("far2", C.c_double), ("rho", C.c_double), ("cand_total", C.c_int64), ("n_counted", C.c_int64)]


TUNE_HIST_LEN = 2 * 16384 + 4096 + 64


class HostFrame(C.Structure):
    """cosmo_host_frame: one frame's host buffers for cosmo_step_host."""
    _fields_ = [("n", C.c_int32), ("reserved", C.c_int32), ("frame", C.c_int64),
                ("pos", _vp), ("vel", _vp), ("mass", _vp), ("mass_hi", _vp), ("mass_lo", _vp), ("radius", _vp),
                ("volume", _vp), ("eaten", _vp), ("id", _vp), ("flags", _vp),
                ("pos_out", _vp), ("vel_out", _vp), ("mass_out", _vp), ("mass_hi_out", _vp), ("mass_lo_out", _vp),
                ("radius_out", _vp), ("volume_out", _vp), ("eaten_out", _vp), ("id_out", _vp), ("flags_out", _vp),
                ("status_out", _vp), ("events_out", _vp), ("events_cap", C.c_int64)]


class StepParams(C.Structure):
    _fields_ = [("G", C.c_double), ("dt", C.c_double), ("r_max", C.c_double),
                ("opts", C.c_uint32), ("n_upper", C.c_int32), ("i_begin", C.c_int32),
                ("i_end", C.c_int32), ("pos_exp", C.c_int32), ("mass_exp", C.c_int32),
                ("jsplit", C.c_int32), ("detect_mode", C.c_int32), ("resolve_mode", C.c_int32),
                ("small_tiles", C.c_int32), ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("sym_chunk", C.c_int32), ("near_mode", C.c_int32), ("sym_rows", C.c_int32), ("reserved1", C.c_int32),
                ("cell_size", C.c_double),
                ("grid_origin_x", C.c_double), ("grid_origin_y", C.c_double), ("cells_per_side", C.c_int32), ("jsplit_detect", C.c_int32),
                ("giant_radius", C.c_double)]


# include/cosmosim_b200_state3d.h
F3_MASS_INT, F3_DENS_INT = 1, 2
OPT3_COLLISIONS, OPT3_FP32, OPT3_STORE_ACC, OPT3_F64_SYM, OPT3_DEFER_INTEGRATE, OPT3_F32_SYM = 1, 4, 8, 16, 32, 64
ERR3_INT_OVERFLOW = 16


class State3(C.Structure):
    _fields_ = [("cap", C.c_int32), ("n_total", C.c_int32), ("pos", _vp), ("vel", _vp), ("mass", _vp),
                ("mass_hi", _vp), ("mass_lo", _vp), ("density", _vp), ("radius", _vp), ("id", _vp),
                ("flags", _vp), ("status", _vp), ("events", _vp), ("event_cap", C.c_int64)]


class Scratch3(C.Structure):
    _fields_ = [("h", _vp), ("pk", _vp), ("acc", _vp), ("pos_new", _vp), ("vel_new", _vp), ("xx_new", _vp),
                ("rinf", _vp), ("partial", _vp), ("jsplit_max", C.c_int32), ("reserved0", C.c_int32),
                ("tile_ctr", _vp), ("pairs", _vp), ("pairs_sorted", _vp), ("pair_cap", C.c_int64),
                ("exists", _vp), ("cur_pos", _vp), ("cur_vel", _vp), ("touched", _vp), ("scan_in", _vp),
                ("scan_out", _vp), ("scan_blk", _vp),
                ("t_pos", _vp), ("t_vel", _vp), ("t_mass", _vp), ("t_mass_hi", _vp), ("t_mass_lo", _vp),
                ("t_density", _vp), ("t_radius", _vp), ("t_id", _vp), ("t_flags", _vp),
                ("cell_of", _vp), ("cell_count", _vp), ("cell_start", _vp), ("cell_items", _vp), ("giants", _vp),
                ("chunk_bounds", _vp), ("heavy_rows", _vp), ("grid_cells", C.c_int32), ("giant_cap", C.c_int32),
                ("colbuf", _vp), ("colbuf_rows", C.c_int64), ("grid_dev", _vp), ("tune_hist", _vp),
                ("row_count", _vp), ("row_start", _vp), ("label", _vp), ("comp_root", _vp), ("comp_items", _vp),
                ("pair_flag", _vp), ("scan_a", _vp), ("corr", _vp)]


class StepParams3(C.Structure):
    _fields_ = [("G", C.c_double), ("dt", C.c_double), ("opts", C.c_uint32), ("n_upper", C.c_int32),
                ("i_begin", C.c_int32), ("i_end", C.c_int32), ("pos_exp", C.c_int32), ("mass_exp", C.c_int32),
                ("jsplit", C.c_int32), ("jsplit_detect", C.c_int32),
                ("detect_mode", C.c_int32), ("resolve_mode", C.c_int32), ("cells_per_side", C.c_int32),
                ("sym_chunk", C.c_int32), ("cell_size", C.c_double), ("grid_origin_x", C.c_double),
                ("grid_origin_y", C.c_double), ("giant_radius", C.c_double),
                ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("near_mode", C.c_int32),
                ("reserved2", C.c_int32)]


_S3 = [C.POINTER(State3), C.POINTER(Scratch3), C.POINTER(StepParams3), _vp]

# name -> (restype, argtypes); every symbol the headers under include/ declare
SYMBOLS = {
    "cosmo3_accel_integrate": (C.c_int, _S3),
    "cosmo3_detect_pairs": (C.c_int, _S3),
    "cosmo3_resolve_commit": (C.c_int, _S3),
    "cosmo3_step": (C.c_int, _S3),
    "cosmo3_integrate": (C.c_int, _S3),
    "cosmo3_seal_pair_row": (C.c_int, [C.POINTER(State3), _vp, _vp]),
    "cosmo3_append_pair_rows": (C.c_int, [C.POINTER(State3), C.POINTER(Scratch3), _vp, C.c_int32, C.c_int64, _vp]),
    "cosmo3_suggest_jsplit": (C.c_int, [C.c_int32, C.c_int32, C.c_int, C.c_int32]),
    "cosmo3_host_truediv": (C.c_double, [C.c_uint64] * 4),
    "cosmo3_host_radius": (C.c_double, [C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_double]),
    "cosmo_abi_version": (C.c_int, []),
    "cosmo_launch_count": (C.c_longlong, []),
    "cosmo_last_error": (C.c_char_p, []),
    "cosmo_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "cosmo_suggest_jsplit": (C.c_int, [C.c_int32, C.c_int32, C.c_int, C.c_int32]),
    "cosmo_begin_frame": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_accel_integrate": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_energies": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_integrate": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_detect_pairs": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_resolve_commit": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_step": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams), _vp]),
    "cosmo_append_pairs": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), _vp, _vp, C.c_int64, _vp]),
    "cosmo_append_pair_rows": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), _vp, C.c_int32, C.c_int64, _vp]),
    "cosmo_seal_pair_row": (C.c_int, [C.POINTER(State), _vp, _vp]),
    "cosmo_workspace_bytes": (C.c_int64, [C.c_int32, C.c_uint32]),
    "cosmo_workspace_init": (C.c_int, [_vp, C.c_int64, C.c_int32, C.c_uint32, C.POINTER(State), C.POINTER(Scratch), _vp]),
    "cosmo_params_default": (C.c_int, [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_double,
                                       C.c_double, C.POINTER(Scratch), C.POINTER(StepParams)]),
    "cosmo_step_host": (C.c_int, [C.POINTER(State), C.POINTER(Scratch), C.POINTER(StepParams),
                                  C.POINTER(HostFrame), _vp]),
    "cosmo_accel_workspace_bytes": (C.c_int64, [C.c_int32]),
    "cosmo_accel": (C.c_int, [_vp, _vp, C.c_int32, C.c_double, C.c_int, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "cosmo_accel_host": (C.c_int, [_vp, _vp, C.c_int32, C.c_double, C.c_int, C.c_int32, C.c_int32, _vp,
                                   _vp, C.c_int64, _vp]),
    "cosmo_sym_tuned_available": (C.c_int, []),
    "cosmo_microbench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_int64), _vp]),
    "cosmo_host_pow_third": (C.c_double, [C.c_double]),
    "cosmo_host_u128_to_double": (C.c_double, [C.c_uint64, C.c_uint64]),
    "cosmo_host_mass_ge": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_double,
                                     C.c_int, C.c_uint64, C.c_uint64, C.c_double]),
}

_lib = None


class CosmoError(RuntimeError):
    pass


def lib():
    """Load libcosmosim_b200.so (building it in-tree if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.is_stale():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(path):
                raise CosmoError(
                    "libcosmosim_b200.so is missing and cannot be built (%s); there is no "
                    "CPU or PyTorch fallback for the physics update" % e)
    L = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if L.cosmo_abi_version() != ABI_VERSION:
        raise CosmoError("ABI version mismatch: library %d, binding %d" % (L.cosmo_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().cosmo_last_error().decode("utf-8", "replace")
        raise CosmoError("%s failed (%d): %s" % (what or "cosmosim_b200 call", rc, msg))
