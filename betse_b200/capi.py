"""ctypes binding of include/betse_b200.h (the C-ABI of libbetse_b200.so).

The structures mirror the header field for field; tests/test_capi.py checks that the
library exports every symbol the header declares.  There is no CPU fallback: if the
library is missing or no CUDA device is usable, calls raise ``BetseB200Error``.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbetse_b200.so")

MAX_IONS = 8
ABI_VERSION = 3
NKERNELS = 8

STATUS_NAN_VM = 1
STATUS_NAN_CONC = 2
STATUS_NEG_CLAMP = 4
STATUS_XCHG_TIMEOUT = 8
STATUS_NEG_NET = 16
STEP_DIAG = 1
XCHG_X1, XCHG_X2 = 0, 1
XCHG_PUSH, XCHG_WAIT = 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_d8 = C.c_double * MAX_IONS


class BetseB200Error(RuntimeError):
    pass


class Mesh(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("n_mems", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
        ("mem_to_cells", _ip), ("cell_mem_ptr", _ip), ("nn_i", _ip), ("map_mem2ecm", _ip),
        ("bflags_mems", _bp),
        ("mem_sa", _dp), ("mem_nx", _dp), ("mem_ny", _dp), ("R_rads", _dp),
        ("cell_vol", _dp), ("cell_sa", _dp), ("diviterm", _dp), ("num_mems", _dp),
        ("memSa_per_envSquare", _dp), ("gj_default_weights", _dp),
        ("delta", C.c_double), ("gj_len", C.c_double), ("ecm_vol", C.c_double),
        ("memsa_mean", C.c_double),
        ("n_cells_owned", C.c_int32), ("n_mems_owned", C.c_int32), ("n_flux_slots", C.c_int32),
        ("y0", C.c_int32), ("ny_global", C.c_int32), ("y_own0", C.c_int32), ("y_own1", C.c_int32),
        ("ecm_slot_ptr", _ip), ("ecm_slot_idx", _ip),
    ]


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_ions", C.c_int32),
        ("iNa", C.c_int32), ("iK", C.c_int32), ("iCa", C.c_int32), ("iP", C.c_int32),
        ("z", _d8), ("D_free", _d8), ("D_gj", _d8), ("c_env_bound", _d8), ("cenv_uniform", _d8),
        ("F", C.c_double), ("R", C.c_double), ("q", C.c_double), ("kb", C.c_double),
        ("eo", C.c_double), ("er", C.c_double), ("cm", C.c_double), ("tm", C.c_double),
        ("NAv", C.c_double), ("mu", C.c_double),
        ("T_sim", C.c_double), ("T_p", C.c_double), ("dt", C.c_double),
        ("alpha_NaK", C.c_double), ("alpha_Ca", C.c_double), ("KmNK_Na", C.c_double),
        ("KmNK_K", C.c_double), ("KmNK_ATP", C.c_double), ("KmCa_Ca", C.c_double),
        ("KmCa_ATP", C.c_double),
        ("cATP", C.c_double), ("cADP", C.c_double), ("cPi", C.c_double), ("deltaGATP", C.c_double),
        ("gj_surface", C.c_double), ("gj_vthresh", C.c_double), ("gj_min", C.c_double),
        ("rho_pump", C.c_double), ("rho_channel", C.c_double),
        ("cell_height", C.c_double), ("vol_env", C.c_double), ("cell_radius", C.c_double),
        ("true_cell_size", C.c_double),
        ("ko_env", C.c_double), ("sharpness", C.c_double), ("cell_polarizability", C.c_double),
        ("smooth_cells", C.c_double),
        ("bound_V", C.c_double * 4), ("gauss_w", C.c_double * 5),
        ("NaKATP_block_scalar", C.c_double), ("gj_block_scalar", C.c_double),
        ("is_ecm", C.c_int32), ("v_sensitive_gj", C.c_int32), ("cluster_open", C.c_int32),
        ("fast_update_ecm", C.c_int32),
        ("sigma_env", C.c_double), ("reserved_d", C.c_double),
    ]


STATE_FIELDS = [
    "cc_cells", "cc_at_mem_cell", "cc_env", "vm", "gjopen", "Dm_cells", "D_env_eff", "E_env_x",
    "E_env_y", "v_env", "rho_env", "rho_cells", "vm_ave", "Phi_b", "extra_rho_cells",
    "extra_rho_env", "extra_J_mem", "NaKATP_block", "gj_block",
    "fluxes_mem", "fluxes_gj", "fluxes_env_x", "fluxes_env_y", "rate_NaKATP",
    "Jmem", "Jgj", "Jn", "I_mem", "Jc", "Emc", "dvm",
    "J_cell_x", "J_cell_y", "E_cell_x", "E_cell_y", "sigma_cell", "E_gj_x", "E_gj_y",
    "J_env_x", "J_env_y", "B_field", "Jtx", "Jty", "cenv_uniform", "vm_cell",
    "D_env_weight",
]


class StateHost(C.Structure):
    _fields_ = [(f, _dp) for f in STATE_FIELDS]


class FastHost(C.Structure):
    """betse_fast_host (fast / equivalent-circuit solver)."""
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "vm_ave", "gjopen", "G_Leak", "E_Leak", "G_gj", "sigma_cell", "extra_J_mem",
        "vgj", "Jn", "Emx", "Emy", "J_cell_x", "J_cell_y", "E_cell_x", "E_cell_y")]


class GateTerm(C.Structure):
    _fields_ = [("type", C.c_int32), ("pad", C.c_int32), ("p", C.c_double * 4)]


class Channel(C.Structure):
    _fields_ = [
        ("ion", C.c_int32), ("mpower", C.c_int32), ("hpower", C.c_int32), ("kind", C.c_int32 * 4),
        ("handler", C.c_int32), ("mod_prog", C.c_int32), ("same_gates", C.c_int32),
        ("a", GateTerm * 4), ("b", GateTerm * 4),
        ("time_unit", C.c_double), ("max_Dm", C.c_double), ("rel_perm", C.c_double), ("v_shift", C.c_double),
        ("target_mask", _bp), ("m0", _dp), ("h0", _dp),
    ]


class Modulator(C.Structure):
    _fields_ = [("target", C.c_int32), ("prog", C.c_int32), ("max_val", C.c_double), ("ion", C.c_int32), ("pad", C.c_int32)]


class LigandGate(C.Structure):
    _fields_ = [("species", C.c_int32), ("ion", C.c_int32), ("extracell", C.c_int32), ("pad", C.c_int32),
                ("K", C.c_double), ("n", C.c_double), ("max_val", C.c_double), ("mod", C.c_double)]


class SubstancePump(C.Structure):
    _fields_ = [("species", C.c_int32), ("into_cell", C.c_int32), ("uses_ATP", C.c_int32), ("pad", C.c_int32),
                ("max_val", C.c_double), ("Km", C.c_double)]


TR_MAX_TERMS = 12


class TransporterTerm(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_int32), ("sign", C.c_int32), ("pad", C.c_int32), ("coeff", C.c_double)]


class Transporter(C.Structure):
    _fields_ = [("prog", C.c_int32), ("n_terms", C.c_int32), ("net_z", C.c_double), ("cell_mask", _bp), ("env_mask", _bp),
                ("mem_mask", _bp), ("terms", TransporterTerm * TR_MAX_TERMS)]


class Network(C.Structure):
    _fields_ = [
        ("n_species", C.c_int32), ("n_rates", C.c_int32), ("n_programs", C.c_int32),
        ("n_consts", C.c_int32), ("n_cell_arrays", C.c_int32), ("n_mem_arrays", C.c_int32),
        ("c_cells", _dp), ("code", _ip), ("prog_ptr", _ip), ("consts", _dp), ("cell_arrays", _dp),
        ("mem_arrays", _dp), ("growth_mask", _bp), ("stoich", _dp), ("Dgj", _dp), ("z", _dp),
        ("time_factor", _dp),
        ("env_on", _bp), ("Dm", _dp), ("c_bound", _dp), ("c_env", _dp), ("D_env", _dp),
        ("scale_factor", _dp), ("affect_charge", C.c_int32), ("n_modulators", C.c_int32),
        ("modulators", C.POINTER(Modulator)),
        ("ligand_gates", C.POINTER(LigandGate)), ("n_ligand_gates", C.c_int32), ("n_pumps", C.c_int32),
        ("pumps", C.POINTER(SubstancePump)),
        ("transporters", C.POINTER(Transporter)), ("n_transporters", C.c_int32), ("reserved", C.c_int32),
        ("mem_sa_over_vol", _dp),
        ("intra_on", _bp), ("Do", _dp), ("c_mems", _dp), ("R_rads", _dp), ("map_cell2ecm", _ip),
        ("mu_mem", _dp), ("Emc", _dp),
        ("tj_targets", _ip), ("n_tj", C.c_int32), ("reserved2", C.c_int32), ("D_env_raw", _dp), ("TJ_modulator", _dp),
        ("env_rx_prog", _ip), ("n_env_rx", C.c_int32), ("reserved3", C.c_int32), ("stoich_env", _dp),
    ]


class WindowInfo(C.Structure):
    _fields_ = [
        ("base", C.c_void_p), ("bytes", C.c_uint64), ("ipc_handle", C.c_uint8 * 64),
        ("off_cc_mid", C.c_uint64 * 2), ("off_vm_cell", C.c_uint64 * 2), ("off_flux", C.c_uint64),
        ("off_cc_env", C.c_uint64 * 2), ("off_v_raw", C.c_uint64), ("off_flags", C.c_uint64),
        ("n_cells", C.c_int32), ("n_env", C.c_int32), ("nx", C.c_int32), ("n_ions", C.c_int32),
        ("n_flux_slots", C.c_int32), ("reserved", C.c_int32),
    ]


class Neighbor(C.Structure):
    _fields_ = [
        ("side", C.c_int32), ("same_process", C.c_int32), ("info", WindowInfo),
        ("n_send_cells", C.c_int32), ("send_cells", _ip), ("recv_cell0", C.c_int32),
        ("n_send_flux", C.c_int32), ("send_flux", _ip), ("recv_slot0", C.c_int32),
        ("cc_rows", C.c_int32), ("cc_src_row0", C.c_int32), ("cc_dst_row0", C.c_int32),
        ("v_rows", C.c_int32), ("v_src_row0", C.c_int32), ("v_dst_row0", C.c_int32),
    ]


# Every symbol include/betse_b200.h declares (checked against the header in the tests).
SYMBOLS = [
    "betse_abi_version", "betse_device_count", "betse_create", "betse_destroy", "betse_last_error",
    "betse_create_error", "betse_upload_state", "betse_set_schedule", "betse_step",
    "betse_step_profile", "betse_ensemble_step", "betse_fast_setup", "betse_fast_step", "betse_fast_download", "betse_fast_set_channels", "betse_kernel_name", "betse_download_sample",
    "betse_step_phase", "betse_stream", "betse_sync", "betse_update_v", "betse_update_v_phase",
    "betse_set_row_ranges", "betse_window", "betse_attach_neighbor", "betse_exchange",
    "betse_set_channels", "betse_channel_state", "betse_set_network", "betse_network_state",
    "betse_network_env_state", "betse_network_mem_state", "betse_network_tj_modulator", "betse_network_set_events", "betse_set_noise_flux",
    "betse_host_alloc", "betse_host_alloc_on", "betse_host_copy", "betse_host_expand", "betse_host_free",
]

_lib = None


def load(build_if_missing=True):
    """Load libbetse_b200.so (building it in-tree with nvcc if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise BetseB200Error("libbetse_b200.so not built; run `python -m betse_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(os.environ.get("BETSE_B200_LIB") or LIB_PATH)      # (the override is for A/B timing of two builds on one box)
    vp = C.c_void_p
    lib.betse_abi_version.restype = C.c_int
    lib.betse_device_count.restype = C.c_int
    lib.betse_create.argtypes = [C.POINTER(vp), C.POINTER(Mesh), C.POINTER(Params), C.c_int]
    lib.betse_destroy.argtypes = [vp]
    lib.betse_destroy.restype = None
    lib.betse_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.betse_create_error.argtypes = [C.c_char_p, C.c_size_t]
    lib.betse_upload_state.argtypes = [vp, C.POINTER(StateHost)]
    lib.betse_set_schedule.argtypes = [vp, C.POINTER(Params)]
    lib.betse_step.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    lib.betse_step_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_float),
                                       C.POINTER(C.c_float * NKERNELS), C.POINTER(C.c_int * NKERNELS)]
    lib.betse_ensemble_step.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    lib.betse_fast_setup.argtypes = [vp, C.POINTER(FastHost)]
    lib.betse_fast_step.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    lib.betse_fast_download.argtypes = [vp, C.POINTER(FastHost)]
    lib.betse_fast_set_channels.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.betse_kernel_name.argtypes = [C.c_int]
    lib.betse_kernel_name.restype = C.c_char_p
    lib.betse_download_sample.argtypes = [vp, C.POINTER(StateHost)]
    lib.betse_update_v_phase.argtypes = [vp, C.c_int]
    lib.betse_set_row_ranges.argtypes = [vp] + [C.c_int] * 6
    lib.betse_window.argtypes = [vp, C.POINTER(WindowInfo)]
    lib.betse_attach_neighbor.argtypes = [vp, C.POINTER(Neighbor)]
    lib.betse_exchange.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.betse_set_channels.argtypes = [vp, C.c_int, C.POINTER(Channel), C.c_int]
    lib.betse_channel_state.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp, _dp]
    lib.betse_set_network.argtypes = [vp, C.c_int, C.POINTER(Network)]
    lib.betse_network_state.argtypes = [vp, C.c_int, _dp, _dp]
    lib.betse_network_env_state.argtypes = [vp, C.c_int, _dp]
    lib.betse_network_mem_state.argtypes = [vp, C.c_int, _dp]
    lib.betse_network_tj_modulator.argtypes = [vp, _dp]
    lib.betse_network_set_events.argtypes = [vp, C.c_int, _dp, _dp]
    lib.betse_set_noise_flux.argtypes = [vp, C.c_int, _dp]
    lib.betse_step_phase.argtypes = [vp, C.c_int, C.c_int]
    lib.betse_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.betse_host_alloc_on.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    lib.betse_host_copy.argtypes = [vp, vp, C.c_size_t]
    lib.betse_host_copy.restype = None
    lib.betse_host_expand.argtypes = [_dp, _dp, _ip, C.c_int, C.c_size_t, C.c_size_t]
    lib.betse_host_expand.restype = None
    lib.betse_host_free.argtypes = [vp]
    lib.betse_host_free.restype = None
    lib.betse_stream.argtypes = [vp, C.POINTER(vp)]
    lib.betse_sync.argtypes = [vp, C.POINTER(C.c_uint32)]
    lib.betse_update_v.argtypes = [vp]
    if lib.betse_abi_version() != ABI_VERSION:
        raise BetseB200Error("libbetse_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr_f64(a):
    return a.ctypes.data_as(_dp)


def ptr_i32(a):
    return a.ctypes.data_as(_ip)
