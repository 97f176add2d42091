"""ctypes binding of libplnerf_b200.so (the C ABI in include/plnerf_b200.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a only).  There is no fallback of any kind:
if the shared object is missing or fails to load, importing the compute entry points raises.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libplnerf_b200.so")
# Developer library: the same sources with -DPLNERF_DEBUG (bring-up GEMMs, MMA-rate microbenchmarks, timeline tracing,
# environment knobs).  PLNERF_DEBUG_LIB=1 makes lib() load it instead of the product library -- a developer switch for
# the scripts under tests/gpu_*.py, never a fallback: the product library has no debug entry points and reads no
# environment variables.
DEBUG_LIB_PATH = os.path.join(HERE, "libplnerf_b200_debug.so")
USE_DEBUG_LIB = os.environ.get("PLNERF_DEBUG_LIB", "") not in ("", "0")
SOURCES = ["ops.cu", "mlp_fwd.cu", "api.cu"]
HEADERS = ["common.cuh", "ops.cuh", "umma.cuh", "mlp_fwd3.cuh", "debug_kernels.cuh", os.path.join("..", "..", "include", "plnerf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

MAX_DEPTH = 16
ABI_VERSION = 2
MODE_CONSTANT, MODE_LINEAR = 0, 1
COLOR_MIDPOINT, COLOR_LEFT = 0, 1
PREC_BF16, PREC_BF16X3 = 0, 1

fptr = C.POINTER(C.c_float)


class NetDesc(C.Structure):
    _fields_ = [("D", C.c_int32), ("W", C.c_int32), ("input_ch", C.c_int32), ("input_ch_views", C.c_int32),
                ("output_ch", C.c_int32), ("use_viewdirs", C.c_int32), ("n_skips", C.c_int32),
                ("skips", C.c_int32 * MAX_DEPTH)]


class NetParams(C.Structure):
    _fields_ = [("pts_w", C.c_void_p * MAX_DEPTH), ("pts_b", C.c_void_p * MAX_DEPTH),
                ("views_w", C.c_void_p), ("views_b", C.c_void_p), ("feature_w", C.c_void_p),
                ("feature_b", C.c_void_p), ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p),
                ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p), ("output_w", C.c_void_p), ("output_b", C.c_void_p)]


class NetGrads(C.Structure):
    _fields_ = NetParams._fields_


class RenderCfg(C.Structure):
    _fields_ = [("N_samples", C.c_int32), ("N_importance", C.c_int32), ("mode", C.c_int32),
                ("color_mode", C.c_int32), ("white_bkgd", C.c_int32), ("lindisp", C.c_int32),
                ("farcolorfix", C.c_int32), ("perturb", C.c_int32), ("raw_noise_std", C.c_float),
                ("zero_tol", C.c_float), ("epsilon", C.c_float), ("multires", C.c_int32),
                ("multires_views", C.c_int32), ("precision", C.c_int32), ("seed", C.c_uint64),
                ("ray_id_offset", C.c_uint64)]


class RenderOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "raw", "rgb0", "disp0",
                                          "acc0", "depth0", "z_std", "z_vals", "inds")]


class RenderGrads(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("g_rgb_map", "g_disp_map", "g_acc_map", "g_depth_map", "g_rgb0", "g_disp0",
                                          "g_acc0", "g_depth0")]


def needs_build(path=None):
    path = path or LIB_PATH
    if not os.path.isfile(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, debug=False):
    """Compile csrc/*.cu into libplnerf_b200.so for sm_100a (nvcc cross-compiles without a GPU).
    debug=True builds the developer library libplnerf_b200_debug.so (-DPLNERF_DEBUG -DPLNERF_ENABLE_TRACE) instead."""
    path = DEBUG_LIB_PATH if debug else LIB_PATH
    if not force and not needs_build(path):
        return path
    cmd = ["nvcc"] + NVCC_FLAGS + (["-DPLNERF_DEBUG", "-DPLNERF_ENABLE_TRACE"] if debug else []) + ["-o", path] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return path


_lib = None

_SIGS = {
    "plnerf_last_error": (C.c_char_p, []),
    "plnerf_abi_version": (C.c_int, []),
    "plnerf_launch_count": (C.c_uint64, []),
    "plnerf_encode": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "plnerf_pack_rays": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "plnerf_pack_pixel_rays": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "plnerf_stratified_z": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "plnerf_packed_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int]),
    "plnerf_pack_weights": (C.c_int, [C.POINTER(NetDesc), C.POINTER(NetParams), C.c_int, C.c_void_p, C.c_void_p]),
    "plnerf_query_workspace_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int64]),
    "plnerf_network_query": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "plnerf_train_stash_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int64, C.c_int]),
    "plnerf_network_query_train": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                             C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "plnerf_packed_bwd_bytes": (C.c_size_t, [C.POINTER(NetDesc)]),
    "plnerf_pack_weights_bwd": (C.c_int, [C.POINTER(NetDesc), C.POINTER(NetParams), C.c_void_p, C.c_void_p]),
    "plnerf_pack_weights_train": (C.c_int, [C.c_int, C.POINTER(C.POINTER(NetDesc)), C.POINTER(C.POINTER(NetParams)),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "plnerf_network_query_bwd": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_size_t, C.POINTER(NetGrads), C.c_void_p]),
    "plnerf_mlp_forward": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "plnerf_raw2outputs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_raw2outputs_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_sample_pdf_pl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_sample_pdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                    C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_sample_pdf_pl_return_u": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_float,
                                                C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "plnerf_sample_pdf_return_u": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                             C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_sample_pdf_pl_return_u_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_sample_pdf_return_u_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_merge_samples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_render_workspace_bytes": (C.c_size_t, [C.POINTER(RenderCfg), C.POINTER(NetDesc), C.c_int64]),
    "plnerf_render_rays_fwd": (C.c_int, [C.POINTER(RenderCfg), C.POINTER(NetDesc), C.c_void_p, C.POINTER(NetDesc),
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.POINTER(RenderOut), C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    "plnerf_render_train_workspace_bytes": (C.c_size_t, [C.POINTER(RenderCfg), C.POINTER(NetDesc), C.POINTER(NetDesc), C.c_int64]),
    "plnerf_render_rays_fwd_train": (C.c_int, [C.POINTER(RenderCfg), C.POINTER(NetDesc), C.c_void_p, C.POINTER(NetDesc),
                                               C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.POINTER(RenderOut), C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
    "plnerf_render_rays_bwd": (C.c_int, [C.POINTER(RenderCfg), C.POINTER(NetDesc), C.c_void_p, C.c_void_p,
                                         C.POINTER(NetDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                         C.c_void_p, C.c_void_p, C.POINTER(RenderGrads), C.POINTER(NetGrads),
                                         C.POINTER(NetGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plnerf_mse_loss_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "plnerf_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_int64, C.c_int, C.c_void_p]),
    "plnerf_train_rays_mse": (C.c_int, [C.POINTER(RenderCfg), C.POINTER(NetDesc), C.c_void_p, C.c_void_p,
                                        C.POINTER(NetDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_void_p, C.POINTER(RenderOut), C.POINTER(NetGrads), C.POINTER(NetGrads),
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "plnerf_profile_enable": (C.c_int, [C.c_int]),
    "plnerf_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}

# developer library only (libplnerf_b200_debug.so); not part of the ABI
_DEBUG_SIGS = {
    "plnerf_debug_umma_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "plnerf_debug_umma_gemm_mn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p,
                                            C.c_void_p]),
    "plnerf_debug_set_trace": (C.c_int, [C.c_void_p]),
    "plnerf_debug_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "plnerf_debug_umma_gemm_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                            C.c_uint32, C.c_void_p, C.c_void_p]),
}

# symbols that include/plnerf_b200.h declares (checked by tests/test_abi.py)
PUBLIC_SYMBOLS = list(_SIGS)


def lib():
    """The loaded library (loads on first use).  Raises if it is not built / not loadable."""
    global _lib
    if _lib is None:
        path = DEBUG_LIB_PATH if USE_DEBUG_LIB else LIB_PATH
        if not os.path.isfile(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(plnerf_b200 has no fallback path)")
        _lib = _load(path, dict(_SIGS, **_DEBUG_SIGS) if USE_DEBUG_LIB else _SIGS)
    return _lib


def _load(path, sigs):
    L = C.CDLL(path)
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.plnerf_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{os.path.basename(path)}: ABI version mismatch")
    return L


_dlib = None


def debug_lib():
    """The developer library (bring-up GEMMs, microbenchmarks, tracing) -- for tests/gpu_*.py and the descriptor-pinning
    test only.  Separate handle from lib(): the product path never touches it."""
    global _dlib
    if _dlib is None:
        if USE_DEBUG_LIB:
            _dlib = lib()
        else:
            if not os.path.isfile(DEBUG_LIB_PATH):
                raise RuntimeError(f"{DEBUG_LIB_PATH} is missing: build it with plnerf_b200._lib.build(debug=True)")
            _dlib = _load(DEBUG_LIB_PATH, dict(_SIGS, **_DEBUG_SIGS))
    return _dlib


def check(rc):
    if rc != 0:
        msg = lib().plnerf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"plnerf_b200 error {rc}: {msg}")
