"""ctypes binding of ``libsmcb200.so`` (C ABI in ``include/smcb200.h``).

The shared library is built IN-TREE by ``build_library()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``) so that it
travels with the repository snapshot to the GPU box.  There is no CPU fallback: every compute entry point raises when the
library or a CUDA device is missing.
"""
import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmcb200.so")
_SOURCES = ["smcb_api.cu", "resample.cuh", "step.cuh", "operators.cuh", "common.cuh", "models.h", "philox.h", "scan_tile.h",
            "exact_scan.h", "column.cuh", "move.cuh", "plugin.cuh", "gpf.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
              "-diag-suppress", "128"]


ABI_VERSION = 200  # SMCB_VERSION of include/smcb200.h


class SmcbError(RuntimeError):
    pass


class smcb_config(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("proposal", C.c_int32), ("algorithm", C.c_int32), ("resampler", C.c_int32),
        ("particles", C.c_int64), ("batch", C.c_int32), ("n_raw_params", C.c_int32),
        ("params_host", C.POINTER(C.c_float)), ("param_cols", C.c_int32), ("ess_threshold", C.c_float),
        ("seed", C.c_uint64), ("history_rows", C.c_int32), ("fold_lookahead", C.c_int32), ("exact_weights", C.c_int32),
        ("column_offset", C.c_int32), ("lin_steps", C.c_int32), ("lin_alpha", C.c_float), ("lin_second_order", C.c_int32),
        ("nested_samples", C.c_int32),
    ]


class smcb_info(C.Structure):
    _fields_ = [
        ("particles", C.c_int64), ("ld", C.c_int64), ("batch", C.c_int32), ("state_dim", C.c_int32), ("obs_dim", C.c_int32),
        ("t", C.c_int32), ("history_rows", C.c_int32), ("slow_tiles", C.c_int32), ("kernel_launches", C.c_int64),
        ("lb_windows", C.c_int64), ("lb_fail", C.c_int32), ("reserved", C.c_int32),
    ]


# every symbol include/smcb200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("smcb_version", C.c_int, []),
    ("smcb_abi_signature", C.c_int, []),
    ("smcb_last_error", C.c_char_p, []),
    ("smcb_device_count", C.c_int, []),
    ("smcb_filter_create", C.c_int, [C.POINTER(smcb_config), C.POINTER(_P)]),
    ("smcb_filter_destroy", C.c_int, [_P]),
    ("smcb_filter_set_params", C.c_int, [_P, C.POINTER(C.c_float), C.c_int32, C.c_int32, _P]),
    ("smcb_filter_info", C.c_int, [_P, C.POINTER(smcb_info)]),
    ("smcb_filter_initialize", C.c_int, [_P, _P]),
    ("smcb_filter_refresh_state", C.c_int, [_P, C.c_int32, _P]),
    ("smcb_filter_set_observations", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P]),
    ("smcb_filter_run", C.c_int, [_P, C.c_int32, _P]),
    ("smcb_filter_run_stepwise", C.c_int, [_P, C.c_int32, _P]),
    ("smcb_filter_profile", C.c_int, [_P, C.c_int32, _P, _P]),
    ("smcb_filter_batch_filter_host", C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P]),
    ("smcb_filter_attach_exchange", C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("smcb_filter_exchange_wait", C.c_int, [_P, C.POINTER(_P), _P]),
    ("smcb_filter_set_noise", C.c_int, [_P, _P, _P, _P]),
    ("smcb_filter_dump_noise", C.c_int, [_P, _P, _P, _P]),
    ("smcb_filter_set_nested_noise", C.c_int, [_P, _P, _P]),
    ("smcb_filter_ptr", C.c_int, [_P, C.c_int32, C.POINTER(_P)]),
    ("smcb_filter_sync_stats", C.c_int, [_P, _P]),
    ("smcb_normalize", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, _P, _P]),
    ("smcb_get_ess", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    ("smcb_systematic", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, C.c_uint64, _P, C.c_int64,
                                  C.c_int64, _P]),
    ("smcb_multinomial", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, C.c_uint64, _P, C.c_int64,
                                   C.c_int64, _P]),
    ("smcb_filter_pre_weight", C.c_int, [_P, _P, _P, _P, _P]),
    ("smcb_filter_sample_and_weight", C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, _P]),
    ("smcb_filter_set_ess_threshold", C.c_int, [_P, C.c_float]),
    ("smcb_filter_predict_path", C.c_int, [_P, C.c_int32, _P, _P, _P, _P]),
    ("smcb_filter_resample_columns", C.c_int, [_P, _P, C.c_int32, _P]),
    ("smcb_filter_exchange_columns", C.c_int, [_P, _P, _P, _P]),
    ("smcb_residual", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _P, C.c_uint64, _P, C.c_int64, C.c_int64, _P]),
    ("smcb_batched_gather", C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    ("smcb_filter_set_seed", C.c_int, [_P, C.c_uint64]),
    ("smcb_filter_column_record_elems", C.c_int64, [_P]),
    ("smcb_filter_export_columns", C.c_int, [_P, _P, _P]),
    ("smcb_filter_import_columns", C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P]),
    ("smcb_filter_ffbs_step", C.c_int, [_P, _P, _P, _P, _P, C.c_uint64, C.c_int32, _P, _P, _P]),
]

PTR_X, PTR_LOGW, PTR_PREV_INDS, PTR_MEAN, PTR_VAR, PTR_LL, PTR_LL_TOTAL, PTR_HIST_MEAN, PTR_HIST_VAR, PTR_HIST_LL, PTR_ESS, \
    PTR_X_OTHER, PTR_RESAMPLE_LOGW = range(13)

_lib = None
_lock = threading.Lock()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(_HERE, "csrc", s) for s in _SOURCES] + [os.path.join(os.path.dirname(_HERE), "include", "smcb200.h")]
    return any(os.path.exists(s) and os.path.getmtime(s) > built for s in srcs)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compiles ``csrc/smcb_api.cu`` for sm_100a into ``pyfilter_b200/libsmcb200.so`` (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(_HERE, "csrc", "smcb_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SmcbError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


_user_libs = {}


def _bind(lib, lib_path, stale_note=None):
    for name, restype, argtypes in SYMBOLS:
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise SmcbError(f"{lib_path} does not export {name}: the binary does not match include/smcb200.h"
                            + (f" ({stale_note})" if stale_note else ""))
        fn.restype = restype
        fn.argtypes = argtypes
    sig = (C.sizeof(smcb_config) << 16) | C.sizeof(smcb_info)
    if lib.smcb_version() != ABI_VERSION or lib.smcb_abi_signature() != sig:
        raise SmcbError(f"{lib_path} was built from another version of include/smcb200.h (version {lib.smcb_version()} vs {ABI_VERSION}, "
                        f"struct signature {lib.smcb_abi_signature():#x} vs {sig:#x}); rebuild it with pyfilter_b200._lib.build_library(force=True)")
    return lib


def build_user_library(header_path: str, force: bool = False) -> str:
    """A build of the library that carries a user-supplied model (csrc/models.h: ``SMCB_USER_MODEL_HEADER``): the whole C ABI compiled
    once more with the user's header, into ``pyfilter_b200/_user/<hash>/libsmcb200_user.so`` (in-tree, so that it travels with the
    repository snapshot; keyed by the hash of the header and of the library's own sources)."""
    import hashlib

    h = hashlib.sha256()
    h.update(open(header_path, "rb").read())
    for src in _SOURCES:
        h.update(open(os.path.join(_HERE, "csrc", src), "rb").read())
    h.update(open(os.path.join(os.path.dirname(_HERE), "include", "smcb200.h"), "rb").read())
    out_dir = os.path.join(_HERE, "_user", h.hexdigest()[:16])
    out = os.path.join(out_dir, "libsmcb200_user.so")
    if os.path.exists(out) and not force:
        return out
    os.makedirs(out_dir, exist_ok=True)
    import shutil

    local = os.path.join(out_dir, "user_model.h")
    if os.path.abspath(header_path) != os.path.abspath(local):
        shutil.copyfile(header_path, local)
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f'-DSMCB_USER_MODEL_HEADER="{local}"', "-o", out, os.path.join(_HERE, "csrc", "smcb_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SmcbError(f"nvcc failed on the user model:\n{res.stdout[-2000:]}\n{res.stderr[-4000:]}")
    return out


def load_user_library(so_path: str):
    with _lock:
        lib = _user_libs.get(so_path)
        if lib is None:
            lib = _bind(C.CDLL(so_path), so_path)
            _user_libs[so_path] = lib
        return lib


def load_library():
    """Loads the shared library (building it when nvcc is available and the sources are newer) and binds every symbol."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        alt = os.environ.get("SMCB_LIB_PATH")  # diagnostics: an alternative build of the same sources (tools/variants.py)
        stale_note = None
        if alt:
            lib_path = alt
        elif _stale():
            lib_path = LIB_PATH
            try:
                build_library()
            except (SmcbError, FileNotFoundError) as e:
                if not os.path.exists(LIB_PATH):
                    raise SmcbError(f"libsmcb200.so is missing and could not be built ({e}); pyfilter_b200 has no CPU fallback")
                stale_note = f"libsmcb200.so is OLDER than its sources and could not be rebuilt ({str(e)[:200]})"
        else:
            lib_path = LIB_PATH
        lib = C.CDLL(lib_path)
        for name, restype, argtypes in SYMBOLS:
            try:
                fn = getattr(lib, name)
            except AttributeError:
                raise SmcbError(f"{lib_path} does not export {name}: the binary does not match include/smcb200.h"
                                + (f" ({stale_note})" if stale_note else ""))
            fn.restype = restype
            fn.argtypes = argtypes
        # the binary must have been built from THIS interface: version and struct layout
        sig = (C.sizeof(smcb_config) << 16) | C.sizeof(smcb_info)
        if lib.smcb_version() != ABI_VERSION or lib.smcb_abi_signature() != sig:
            raise SmcbError(f"{lib_path} was built from another version of include/smcb200.h (version {lib.smcb_version()} vs {ABI_VERSION}, "
                            f"struct signature {lib.smcb_abi_signature():#x} vs {sig:#x}); rebuild it with pyfilter_b200._lib.build_library(force=True)")
        if stale_note:
            import warnings

            warnings.warn(stale_note + "; its interface matches, its kernels may not", RuntimeWarning)
        _lib = lib
        return lib


def check(rc: int, lib=None):
    if rc != 0:
        msg = (lib or load_library()).smcb_last_error().decode()
        if rc == -4:
            raise NotImplementedError(msg)
        if rc == -1:
            raise ValueError(msg)
        raise SmcbError(f"libsmcb200 error {rc}: {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available() or load_library().smcb_device_count() < 1:
        raise SmcbError("pyfilter_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


class DeviceView:
    """Zero-copy ``__cuda_array_interface__`` wrapper of a borrowed device pointer; ``owner`` keeps the handle alive."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


def as_tensor(ptr: int, shape, typestr: str, owner):
    import torch

    t = torch.as_tensor(DeviceView(ptr, shape, typestr, owner), device="cuda")
    t._smcb_owner = owner
    return t
