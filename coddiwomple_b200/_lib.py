"""ctypes binding of libcdw_b200.so (include/cdw.h).  There is NO CPU fallback: if the CUDA library cannot be loaded
or a call fails, the caller gets an exception."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CDW_LIB") or os.path.join(HERE, "libcdw_b200.so")  # CDW_LIB: a tuning variant built by _build.build(out=...)

CDW_FLAG_RANDOMIZE_VELOCITIES = 1
CDW_FLAG_TRACK_WORKS = 2
CDW_FLAG_ULA = 4
CDW_FLAG_TWIST_REFERENCE_NORMALIZER = 8
CDW_STEP_V, CDW_STEP_R, CDW_STEP_O, CDW_STEP_OPEN, CDW_STEP_CLOSE = 1, 2, 3, 4, 5
CDW_RESAMPLE_MULTINOMIAL = 0
CDW_RESAMPLE_SYSTEMATIC = 1
STAT_WMIN, STAT_SUM_E, STAT_SUM_E2, STAT_NESS, STAT_MEAN, STAT_RESAMPLE, STAT_TOTAL, STAT_NNAN = range(8)
STATS_LEN = 8

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class SystemDesc(C.Structure):
    """struct cdw_system_desc (include/cdw.h)."""
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_lambda", C.c_int32), ("masses", _f64p),
        ("n_bonds", C.c_int32), ("bond_atoms", _i32p), ("bond_slot", _i32p), ("bond_par", _f64p),
        ("n_angles", C.c_int32), ("angle_atoms", _i32p), ("angle_slot", _i32p), ("angle_par", _f64p),
        ("n_torsions", C.c_int32), ("torsion_atoms", _i32p), ("torsion_slot", _i32p), ("torsion_n", _i32p), ("torsion_par", _f64p),
        ("nb_atom_par", _f64p), ("n_atom_offsets", C.c_int32), ("atom_off_idx", _i32p), ("atom_off_par", _f64p),
        ("n_exceptions", C.c_int32), ("exc_par", _f64p), ("n_exc_offsets", C.c_int32), ("exc_off_idx", _i32p), ("exc_off_par", _f64p),
        ("n_pairs", C.c_int32), ("pair_atoms", _i32p), ("pair_nb", _i32p), ("pair_sc_class", _i32p), ("pair_sc_par", _f64p),
        ("slot_sterics_core", C.c_int32), ("slot_sterics_insert", C.c_int32), ("slot_sterics_delete", C.c_int32), ("slot_softcore_alpha", C.c_int32),
        ("n_ext", C.c_int32), ("ext_atom", _i32p), ("ext_slot", _i32p), ("ext_par", _f64p),
        ("n_constraints", C.c_int32), ("cons_atoms", _i32p), ("cons_dist", _f64p),
    ]


class LangevinParams(C.Structure):
    """struct cdw_langevin_params (include/cdw.h)."""
    _fields_ = [("dt", C.c_double), ("gamma", C.c_double), ("kT", C.c_double), ("constraint_tol", C.c_double),
                ("n_steps", C.c_int32), ("flags", C.c_uint32), ("splitting", C.c_uint64), ("trial_counts", C.c_void_p), ("twist", C.c_void_p), ("dense_twist", C.c_void_p)]


class DenseTwist(C.Structure):
    """struct cdw_dense_twist (include/cdw.h): one iteration's dense quadratic twist, device pointers."""
    _fields_ = [("theta", C.c_void_p), ("chol", C.c_void_p), ("A", C.c_void_p), ("b", C.c_void_p), ("beta", C.c_void_p),
                ("logk0", C.c_double), ("cd", C.c_double), ("scratch", C.c_void_p)]


class StaticCache(C.Structure):
    """struct cdw_static_cache (include/cdw.h): static-term forces / energies at the input and output rows."""
    _fields_ = [("force_in", C.c_void_p), ("energy_in", C.c_void_p), ("force_out", C.c_void_p), ("energy_out", C.c_void_p),
                ("peer_force_in", C.c_void_p), ("peer_energy_in", C.c_void_p)]


def _addr(t):
    return None if t is None else (t.data_ptr() if hasattr(t, "data_ptr") else int(t))


def static_cache(force_in, energy_in, force_out, energy_out, peer_force_in=None, peer_energy_in=None):
    return StaticCache(_addr(force_in), _addr(energy_in), _addr(force_out), _addr(energy_out), _addr(peer_force_in), _addr(peer_energy_in))


def make_desc(spec):
    """Build a SystemDesc whose pointers reference `spec`'s numpy arrays (kept alive through the returned tuple)."""
    keep = []

    def ip(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(_i32p)

    def dp(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(_f64p)

    d = SystemDesc()
    d.n_atoms, d.n_lambda, d.masses = spec.n_atoms, spec.n_lambda, dp(spec.masses)
    d.n_bonds, d.bond_atoms, d.bond_slot, d.bond_par = len(spec.bond_slot), ip(spec.bond_atoms), ip(spec.bond_slot), dp(spec.bond_par)
    d.n_angles, d.angle_atoms, d.angle_slot, d.angle_par = len(spec.angle_slot), ip(spec.angle_atoms), ip(spec.angle_slot), dp(spec.angle_par)
    d.n_torsions, d.torsion_atoms, d.torsion_slot = len(spec.torsion_slot), ip(spec.torsion_atoms), ip(spec.torsion_slot)
    d.torsion_n, d.torsion_par = ip(spec.torsion_n), dp(spec.torsion_par)
    d.nb_atom_par = dp(spec.nb_atom_par)
    d.n_atom_offsets, d.atom_off_idx, d.atom_off_par = len(spec.atom_off_par), ip(spec.atom_off_idx), dp(spec.atom_off_par)
    d.n_exceptions, d.exc_par = len(spec.exc_par), dp(spec.exc_par)
    d.n_exc_offsets, d.exc_off_idx, d.exc_off_par = len(spec.exc_off_par), ip(spec.exc_off_idx), dp(spec.exc_off_par)
    d.n_pairs, d.pair_atoms, d.pair_nb = len(spec.pair_nb), ip(spec.pair_atoms), ip(spec.pair_nb)
    d.pair_sc_class, d.pair_sc_par = ip(spec.pair_sc_class), dp(spec.pair_sc_par)
    d.slot_sterics_core, d.slot_sterics_insert, d.slot_sterics_delete, d.slot_softcore_alpha = (int(s) for s in spec.sc_slots)
    d.n_ext, d.ext_atom, d.ext_slot, d.ext_par = len(spec.ext_atom), ip(spec.ext_atom), ip(spec.ext_slot), dp(spec.ext_par)
    d.n_constraints, d.cons_atoms, d.cons_dist = len(spec.cons_dist), ip(spec.cons_atoms), dp(spec.cons_dist)
    return d, keep


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes) — keep in the order of include/cdw.h
    "cdw_version": (C.c_int, []),
    "cdw_last_error": (C.c_char_p, []),
    "cdw_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "cdw_system_create": (C.c_int, [C.POINTER(SystemDesc), C.POINTER(_P)]),
    "cdw_system_destroy": (C.c_int, [_P]),
    "cdw_system_register_protocol": (C.c_int, [_P, _P, C.c_int32, C.c_double, _P]),
    "cdw_system_clear_protocol": (C.c_int, [_P]),
    "cdw_system_set_team": (C.c_int, [_P, C.c_int32]),
    "cdw_system_team": (C.c_int, [_P]),
    "cdw_system_smem_bytes": (C.c_size_t, [_P, C.c_int32]),
    "cdw_system_n_atoms": (C.c_int, [_P]),
    "cdw_system_n_lambda": (C.c_int, [_P]),
    "cdw_reduced_potential": (C.c_int, [_P, _P, C.c_int64, _P, C.c_double, _P, _P]),
    "cdw_energy_forces": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P]),
    "cdw_smc_propagate": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(LangevinParams), C.c_uint64, C.c_uint64,
                                    C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cdw_smc_propagate_cached": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(LangevinParams), C.c_uint64, C.c_uint64,
                                           C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(StaticCache), _P]),
    "cdw_smc_lookahead": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, C.POINTER(LangevinParams), C.c_uint64, C.c_uint64, C.c_int64,
                                    _P, _P, _P, _P, _P]),
    "cdw_smc_lookahead_sharded": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int64, _P, _P, _P, C.c_int64, _P, _P, _P, C.POINTER(LangevinParams),
                                            C.c_uint64, C.c_uint64, C.c_int64, _P, _P, _P, _P, _P]),
    "cdw_smc_propagate_sharded_cached": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int64, _P, _P, C.c_int64, _P, _P, _P, C.POINTER(LangevinParams),
                                                   C.c_uint64, C.c_uint64, C.c_int64, _P, _P, _P, _P, _P, _P, C.POINTER(StaticCache), _P]),
    "cdw_smc_propagate_sharded": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int64, _P, _P, C.c_int64, _P, _P, _P, C.POINTER(LangevinParams), C.c_uint64,
                                            C.c_uint64, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "cdw_baoab": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.POINTER(LangevinParams), C.c_uint64, C.c_uint64, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "cdw_weight_update": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "cdw_weight_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "cdw_weight_stats": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P, _P, C.c_size_t, _P]),
    "cdw_resample_indices": (C.c_int, [C.c_int, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "cdw_smc_resample": (C.c_int, [C.c_int, _P, C.c_int64, C.c_double, _P, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "cdw_smc_weigh_resample": (C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_double, _P, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "cdw_gather_particles": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "cdw_gather_particles_peer": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int64, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "cdw_record_lineage": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P]),
    "cdw_peer_barrier": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "cdw_fill_u64": (C.c_int, [_P, C.c_uint64, C.c_int64, _P]),
    "cdw_philox_uniforms": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, _P, _P]),
    "cdw_ipc_get_handle": (C.c_int, [_P, _P]),
    "cdw_ipc_open_handle": (C.c_int, [_P, C.POINTER(_P)]),
    "cdw_ipc_close_handle": (C.c_int, [_P]),
    "cdw_device_malloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "cdw_device_free": (C.c_int, [_P]),
    "cdw_fma_peak": (C.c_int, [C.c_int, C.c_int, _f64p, _f64p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class CdwError(RuntimeError):
    pass


def load():
    """Load libcdw_b200.so (building is the job of __graft_entry__.build / coddiwomple_b200._build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CdwError(f"{LIB_PATH} is missing: run `python -m coddiwomple_b200._build` (nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cdw_last_error()
        raise CdwError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
