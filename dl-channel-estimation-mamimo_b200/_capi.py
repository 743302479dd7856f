"""ctypes binding of include/mamimo.h.  No math here: every call crosses the C ABI.

The library is the product; there is NO CPU fallback.  If libmamimo_b200.so is missing the
import of this module raises (build it with build.py / __graft_entry__.build()).
"""
import ctypes as C
import os

from . import build as _build

MAX_HIDDEN = 8
ABI_VERSION = 2

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED, ERR_RANGE, ERR_TIMEOUT = range(8)
C64, C128 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
PRECISIONS = {"fp32_simt": 0, "tf32x3": 1, "fp16x3": 2, "bf16x1": 3}
INPUT_MODES = {"ls": 0, "planes": 1, "time_p": 2}


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("n_tx", C.c_int32), ("n_rx", C.c_int32), ("n_ltf", C.c_int32), ("n_sc", C.c_int32), ("n_ps", C.c_int32),
        ("input_mode", C.c_int32), ("precision", C.c_int32),
        ("d_in", C.c_int32), ("d_out", C.c_int32), ("n_hidden", C.c_int32),
        ("hidden", C.c_int32 * MAX_HIDDEN),
        ("len_ltf", C.c_int32), ("max_pkts", C.c_int32), ("act_scale_log2", C.c_int32),
        ("kb_per_chunk", C.c_int32), ("host_chunk_pkts", C.c_int32),
        ("fc_single_cta", C.c_int32),
        ("fc_sm_reserve", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("last_device_flags", C.c_uint32), ("graph_launches", C.c_uint32)]


class Profile(C.Structure):
    _fields_ = [("ls_ms", C.c_double), ("fc_ms", C.c_double), ("stage_ms", C.c_double),
                ("ls_launches", C.c_uint64), ("fc_launches", C.c_uint64), ("stage_launches", C.c_uint64),
                ("lmmse_ms", C.c_double), ("lmmse_launches", C.c_uint64)]


# every symbol include/mamimo.h declares (tests/test_capi_symbols.py checks the header against this)
SYMBOLS = [
    "mamimo_abi_version", "mamimo_status_string", "mamimo_last_error", "mamimo_config_init",
    "mamimo_vht_ltf256", "mamimo_carriers_locations", "mamimo_default_p", "mamimo_pair_row",
    "mamimo_create", "mamimo_destroy", "mamimo_set_pilots", "mamimo_set_pilots_f64", "mamimo_load_layer", "mamimo_finalize_weights",
    "mamimo_ls_estimate", "mamimo_estimate", "mamimo_estimate_stages", "mamimo_predict_planes", "mamimo_predict_time",
    "mamimo_synchronize", "mamimo_poll_flags", "mamimo_get_stats", "mamimo_host_alloc", "mamimo_host_free",
    "mamimo_profile_begin", "mamimo_profile_end", "mamimo_get_debug_counters",
    "mamimo_set_ofdm", "mamimo_ofdm_demod", "mamimo_estimate_time", "mamimo_lmmse", "mamimo_tau_rms", "mamimo_svd",
    "mamimo_gather_create", "mamimo_gather_connect", "mamimo_gather_attach", "mamimo_set_steering_dictionary", "mamimo_omp", "mamimo_ipc_export", "mamimo_ipc_open", "mamimo_ipc_close",
]


def _load():
    # MAMIMO_LIB: an alternative build of the same library (e.g. the -DMAMIMO_FC_DEBUG_COUNTERS one tools/fc_power_probe.py uses)
    path = os.environ.get("MAMIMO_LIB") or _build.LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: the CUDA library is the product and there is no CPU fallback. "
            "Run __graft_entry__.build() (needs nvcc)." % path)
    lib = C.CDLL(path)
    vp, i32, i64, fp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float)
    sig = {
        "mamimo_abi_version": (i32, []),
        "mamimo_status_string": (C.c_char_p, [i32]),
        "mamimo_last_error": (C.c_char_p, [vp]),
        "mamimo_config_init": (None, [C.POINTER(Config)]),
        "mamimo_vht_ltf256": (None, [C.POINTER(C.c_int8)]),
        "mamimo_carriers_locations": (i32, [C.POINTER(i32), i32]),
        "mamimo_default_p": (i32, [i32, fp]),
        "mamimo_pair_row": (i64, [i64, i32, i32, i32, i32]),
        "mamimo_create": (i32, [C.POINTER(Config), C.POINTER(vp)]),
        "mamimo_destroy": (None, [vp]),
        "mamimo_set_pilots": (i32, [vp, vp, vp]),
        "mamimo_set_pilots_f64": (i32, [vp, vp, vp]),
        "mamimo_load_layer": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp]),
        "mamimo_finalize_weights": (i32, [vp]),
        "mamimo_ls_estimate": (i32, [vp, vp, i32, i32, i64, vp, i32, i32, vp]),
        "mamimo_estimate": (i32, [vp, vp, i32, i64, vp, vp, vp, i32, vp]),
        "mamimo_estimate_stages": (i32, [vp, vp, i32, i64, vp, vp, vp, i32, vp, C.c_uint32]),
        "mamimo_predict_planes": (i32, [vp, vp, vp, i64, vp, vp, i32, vp]),
        "mamimo_predict_time": (i32, [vp, vp, vp, i64, vp, vp, i32, vp]),
        "mamimo_synchronize": (i32, [vp]),
        "mamimo_poll_flags": (i32, [vp, vp]),
        "mamimo_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "mamimo_host_alloc": (vp, [C.c_size_t]),
        "mamimo_host_free": (None, [vp]),
        "mamimo_set_ofdm": (i32, [vp, i32, i32, i32, C.POINTER(i32)]),
        "mamimo_ofdm_demod": (i32, [vp, vp, i32, i64, vp, i32, vp]),
        "mamimo_estimate_time": (i32, [vp, vp, i32, i64, vp, vp, vp, i32, vp]),
        "mamimo_lmmse": (i32, [vp, vp, i32, i64, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, i32, i32, vp]),
        "mamimo_tau_rms": (C.c_double, [C.POINTER(C.c_double), i32, i32]),
        "mamimo_svd": (i32, [vp, vp, i32, i64, vp, vp, i32, i32, vp]),
        "mamimo_set_steering_dictionary": (i32, [vp, C.POINTER(C.c_double), i32]),
        "mamimo_omp": (i32, [vp, vp, i32, i32, i64, i32, i32, vp, vp, vp, i32, i32, vp]),
        "mamimo_gather_create": (i32, [vp, i32, i32, i64, C.POINTER(vp), C.POINTER(vp)]),
        "mamimo_gather_connect": (i32, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "mamimo_gather_attach": (i32, [vp, i32, i32, i64, C.POINTER(vp), C.POINTER(vp), vp, vp]),
        "mamimo_ipc_export": (i32, [vp, C.c_char_p]),
        "mamimo_ipc_open": (i32, [C.c_char_p, C.POINTER(vp)]),
        "mamimo_ipc_close": (i32, [vp]),
        "mamimo_get_debug_counters": (i32, [vp, C.POINTER(C.c_uint64), i32]),
        "mamimo_profile_begin": (i32, [vp]),
        "mamimo_profile_end": (i32, [vp, C.POINTER(Profile)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here == missing export: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.mamimo_abi_version() != ABI_VERSION:
        raise ImportError("libmamimo_b200.so ABI version mismatch")
    return lib


lib = _load()


class MamimoError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("mamimo status %d (%s): %s" % (status, lib.mamimo_status_string(status).decode(), msg))
        self.status = status


def check(status, handle=None):
    if status != OK:
        msg = lib.mamimo_last_error(handle)
        raise MamimoError(status, msg.decode() if msg else "")
