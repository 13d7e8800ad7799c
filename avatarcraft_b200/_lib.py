"""ctypes binding of libavatarcraft_b200.so (the C ABI declared in include/avatarcraft_b200.h).

The library is the product: there is NO fallback.  If it is missing or fails to load, every
operator raises -- a CPU or eager-torch path here would void the parity claims.
"""
import ctypes
import os
import subprocess

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("AC_LIB_PATH") or os.path.join(_PKG, "libavatarcraft_b200.so")   # AC_LIB_PATH: tuning variants
SOURCES = ["api_common.cu", "encoder_ops.cu", "nsr_kernels.cu", "nsr_render_tc.cu", "warp_ops.cu", "sh_ops.cu", "raymarch_ops.cu", "frame_ops.cu", "sd_ops.cu", "nsr_train_tc.cu", "nsr_shade_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]

AC_OK, AC_E_INVALID_ARG, AC_E_UNSUPPORTED, AC_E_CUDA, AC_E_WORKSPACE = 0, -1, -2, -3, -4
MLP_BLOB_FLOATS = 10848


def build_variant(path: str, defines) -> str:
    """Compile a tuning variant of the library (extra -D switches) to `path`."""
    srcs = [os.path.join(_PKG, "csrc", s) for s in SOURCES]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", path] + srcs)
    return path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a with nvcc (cross-compiles without a GPU): one object per translation unit
    (in parallel, only the stale ones), then one link."""
    if os.environ.get("AC_LIB_PATH"):
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    csrc = os.path.join(_PKG, "csrc")
    hdrs = [os.path.join(csrc, h) for h in os.listdir(csrc) if h.endswith(".cuh")] + [os.path.join(_ROOT, "include", "avatarcraft_b200.h")]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    objdir = os.path.join(_PKG, "_build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(name):
        src, obj = os.path.join(csrc, name), os.path.join(objdir, name[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            subprocess.check_call([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
            return obj, True
        return obj, False

    with ThreadPoolExecutor(max(1, min(len(SOURCES), os.cpu_count() or 1))) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB_PATH) or any(os.path.getmtime(LIB_PATH) < os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs)
    return LIB_PATH


class NsrModel(ctypes.Structure):
    _fields_ = [("embeddings", ctypes.c_void_p), ("offsets", ctypes.c_void_p), ("mlp_blob", ctypes.c_void_p),
                ("variance", ctypes.c_void_p), ("log2_per_level_scale", ctypes.c_float),
                ("base_resolution", ctypes.c_uint32)]


class NsrRenderArgs(ctypes.Structure):
    _fields_ = [("rays_o", ctypes.c_void_p), ("rays_d", ctypes.c_void_p),
                ("bg_color", ctypes.c_void_p), ("jitter", ctypes.c_void_p), ("alpha_mask", ctypes.c_void_p),
                ("z_in", ctypes.c_void_p), ("pts_in", ctypes.c_void_p), ("near_far_in", ctypes.c_void_p),
                ("n_rays", ctypes.c_uint32), ("num_steps", ctypes.c_uint32), ("upsample_steps", ctypes.c_uint32),
                ("eikonal_segment", ctypes.c_uint32),
                ("bound", ctypes.c_float), ("cos_anneal_ratio", ctypes.c_float), ("normal_epsilon_ratio", ctypes.c_float),
                ("rgb", ctypes.c_void_p), ("depth", ctypes.c_void_p), ("weight_sum", ctypes.c_void_p),
                ("normal", ctypes.c_void_p),
                ("weights", ctypes.c_void_p), ("pts_color", ctypes.c_void_p), ("pts_alpha", ctypes.c_void_p),
                ("z_vals", ctypes.c_void_p),
                ("eikonal", ctypes.c_void_p), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_uint64),
                ("c0_ray_bias", ctypes.c_void_p), ("opacity_only", ctypes.c_uint32), ("skip_masked", ctypes.c_uint32)]


class NsrShadeArgs(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("rays_o", "rays_d", "z_vals", "points", "centre", "fd", "bg_color")] + \
               [("n_rays", ctypes.c_uint32), ("n_samples", ctypes.c_uint32), ("num_steps", ctypes.c_uint32),
                ("bound", ctypes.c_float), ("eps", ctypes.c_float), ("cos_anneal_ratio", ctypes.c_float)] + \
               [(k, ctypes.c_void_p) for k in ("rgb", "depth", "weight_sum", "normal", "eik_partial", "weights", "pts_color", "pts_alpha")]


class NsrShadeGrads(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("g_rgb", "g_weight_sum", "g_normal", "g_depth", "g_eikonal", "wsum_gt")] + \
               [("opacity_weight", ctypes.c_float)] + \
               [(k, ctypes.c_void_p) for k in ("eik_out", "scale", "g_centre", "g_fd", "g_variance", "terms_a", "terms_b", "g_b1", "opacity_loss")] + \
               [("terms_ld", ctypes.c_uint64)]


class WeightNormLayer(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("dW", "v", "g", "dv", "dg")] + \
               [("rows", ctypes.c_int), ("cols", ctypes.c_int), ("ldw", ctypes.c_int), ("scale", ctypes.c_void_p), ("db", ctypes.c_void_p),
                ("db_col", ctypes.c_int)]


_V, _U32, _F, _I, _D, _I64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float, ctypes.c_int, ctypes.c_double, ctypes.c_int64
_SIGNATURES = {
    "ac_version": (ctypes.c_char_p, []),
    "ac_last_cuda_error": (ctypes.c_char_p, []),
    "ac_launch_count": (ctypes.c_uint64, []),
    "ac_launch_count_add": (None, [ctypes.c_uint64]),
    "ac_hash_encode_forward": (_I, [_V, _V, _V, _V, _U32, _U32, _U32, _U32, _F, _U32, _I, _V, _V, _V]),
    "ac_hash_encode_backward": (_I, [_V, _V, _V, _V, _V, _U32, _U32, _U32, _U32, _F, _U32, _I, _V, _V, _V]),
    "ac_sh_encode_forward": (_I, [_V, _V, _U32, _U32, _U32, _I, _V, _V]),
    "ac_sh_encode_backward": (_I, [_V, _V, _U32, _U32, _U32, _V, _V, _V]),
    "ac_hash_encode_forward_pm": (_I, [_V, _V, _V, _V, _U32, _U32, _U32, _U32, _F, _U32, _I, _V, _V]),
    "ac_hash_encode_backward_pm": (_I, [_V, _V, _V, _V, _U32, _U32, _U32, _U32, _F, _U32, _I, _V, _V, _V]),
    "ac_hash_level_scales": (_I, [_V, _U32, _F, _U32, _V]),
    "ac_nsr_pack_mlp": (_I, [_V] * 13 + [_V]),
    "ac_nsr_forward_sdf": (_I, [ctypes.POINTER(NsrModel), _V, _V, _U32, _F, _V]),
    "ac_nsr_sdf_backward": (_I, [ctypes.POINTER(NsrModel), _V, _V, _U32, _F, _V, _V, _V, _V, _V]),
    "ac_nsr_sdf_backward_fused": (_I, [ctypes.POINTER(NsrModel), _V, _V, _U32, _F, _V, _V, _V, _V, _V]),
    "ac_nsr_forward_sdf_stencil": (_I, [ctypes.POINTER(NsrModel), _V, _U32, _F, _F, _V, _V, _V]),
    "ac_nsr_sdf_backward_stencil": (_I, [ctypes.POINTER(NsrModel), _V, _U32, _F, _F, _V, _V, _V, _V, _V, _V, _V]),
    "ac_nsr_sdf_backward_workspace_bytes": (ctypes.c_uint64, [_U32]),
    "ac_nsr_sdf_backward_fused_ws": (_I, [ctypes.POINTER(NsrModel), _V, _V, _U32, _F, _V, _V, _V, _V, _V, ctypes.c_uint64, _V]),
    "ac_nsr_sdf_backward_stencil_ws": (_I, [ctypes.POINTER(NsrModel), _V, _U32, _F, _F, _V, _V, _V, _V, _V, _V, _V, ctypes.c_uint64, _V, _V]),
    "ac_nsr_sdf_feature_cache_bytes": (ctypes.c_uint64, [_U32]),
    "ac_nsr_forward_sdf_stencil_cache": (_I, [ctypes.POINTER(NsrModel), _V, _U32, _F, _F, _V, _V, _V, ctypes.c_uint64, _V]),
    "ac_nsr_shade_forward": (_I, [ctypes.POINTER(NsrModel), ctypes.POINTER(NsrShadeArgs), _V, _V]),
    "ac_nsr_shade_backward": (_I, [ctypes.POINTER(NsrModel), ctypes.POINTER(NsrShadeArgs), ctypes.POINTER(NsrShadeGrads), _V]),
    "ac_absmax_scale": (_I, [_V, _U32, _F, _V, _V]),
    "ac_nsr_sdf_backward_scales": (_I, [ctypes.POINTER(NsrModel), _V, _V, _U32, _V, _V]),
    "ac_fill_uniform": (_I, [_V, ctypes.c_uint64, ctypes.c_uint64, _V]),
    "ac_zero": (_I, [_V, ctypes.c_uint64, _V]),
    "ac_nsr_section_points": (_I, [_V, _V, _V, _U32, _U32, _F, _V, _V]),
    "ac_nsr_ray_points": (_I, [_V, _V, _V, _V, _U32, _U32, _F, _V, _V, _V]),
    "ac_nsr_merge_gather": (_I, [_V, _V, _V, _U32, _U32, _V, _V]),
    "ac_clamp_inplace": (_I, [_V, ctypes.c_uint64, _F, _V]),
    "ac_nsr_take_sdf": (_I, [_V, ctypes.c_uint64, _V, _V]),
    "ac_nsr_weight_norm_backward": (_I, [ctypes.POINTER(WeightNormLayer), _U32, _V]),
    "ac_nsr_forward_color": (_I, [ctypes.POINTER(NsrModel), _V, _V, _V, _V, _U32, _V]),
    "ac_nsr_forward_color_bias": (_I, [ctypes.POINTER(NsrModel), _V, _V, _V, _V, _V, _U32, _V]),
    "ac_nsr_viewdir_bias": (_I, [_V, _V, _U32, _V, _V]),
    "ac_nsr_fd_gradient": (_I, [ctypes.POINTER(NsrModel), _V, _V, _U32, _F, _F, _V]),
    "ac_nsr_render_workspace_bytes": (ctypes.c_uint64, [_U32]),
    "ac_nsr_render": (_I, [ctypes.POINTER(NsrModel), ctypes.POINTER(NsrRenderArgs), _V]),
    "ac_warp_mesh_bytes": (ctypes.c_uint64, [_U32]),
    "ac_warp_prepare_mesh": (_I, [_V, _V, _U32, _U32, _V, _V]),
    "ac_warp_samples_to_canonical": (_I, [_V, _U32, _V, _U32, _V, _F, _V, _V, _V, _V, _V, _V]),
    "ac_warp_samples_to_canonical_rays": (_I, [_V, _U32, _U32, _V, _U32, _V, _F, _V, _V, _V, _V, _V, _V]),
    "ac_warp_samples_to_canonical_ordered": (_I, [_V, _V, _U32, _V, _U32, _V, _F, _V, _V, _V, _V, _V, _V]),
    "ac_warp_samples_to_canonical_masked": (_I, [_V, _V, _U32, _V, _U32, _V, _F, _V, _V, _V]),
    "ac_warp_query_keys": (_I, [_V, _U32, _V, _U32, _F, _V, _V]),
    "ac_mesh_guided_near_far": (_I, [_V, _V, _U32, _V, _U32, _F, _F, _V, _V]),
    "ac_march_rays_train": (_I, [_V, _V, _V, _F, _I, _F, _U32, _U32, _U32, _V, _V, _V, _V, _V, _U32, _V]),
    "ac_composite_rays_train_forward": (_I, [_V, _V, _V, _V, _F, _U32, _U32, _V, _V, _V]),
    "ac_composite_rays_train_backward": (_I, [_V, _V, _V, _V, _V, _V, _V, _V, _F, _U32, _U32, _V, _V, _V]),
    "ac_march_rays": (_I, [_U32, _U32, _V, _V, _V, _V, _F, _U32, _V, _F, _V, _V, _V, _V, _V, _U32, _V]),
    "ac_composite_rays": (_I, [_U32, _U32, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    "ac_compact_rays": (_I, [_U32, _V, _V, _V, _V, _V, _V]),
    "ac_gen_rays": (_I, [_V, _D, _D, _D, _D, _U32, _U32, _D, _D, _D, _D, _I, _V, _V, _V]),
    "ac_select_background": (_I, [_I, _U32, ctypes.c_uint64, _F, _V, _V]),
    "ac_adam_step": (_I, [_V, _V, _V, _V, ctypes.c_uint64, _F, _F, _F, _F, _U32, _F, _V]),
    "ac_sdf_grid_points": (_I, [_V, _V, _U32, _U32, _U32, _V, _V]),
    "ac_iso_surface": (_I, [_V, _V, _V, _U32, _F, _V, _V, ctypes.c_uint64, _V, _V]),
    "ac_sd_gemm_f16": (_I, [_V, _V, _V, _V, _I, _V, _V, _I, _I, _I, _I, _I64, _I64, _I64, _I64, _I, _I, _I64, _I64, _I64, _I64, _I64, _I64, _V]),
    "ac_sd_flash_attention_f16": (_I, [_V, _V, _V, _V, _I, _I, _I, _I, _I, _I64, _I64, _I64, _I64, _F, _V]),
    "ac_sd_conv3x3_f16": (_I, [_V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _V]),
    "ac_sd_group_norm_stats": (_I, [_V, _I, _I, _I, _I, _F, _V, _V, _V]),
    "ac_sd_im2col_f16": (_I, [_V, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _V, _V, _V, _I, _I, _V, _V]),
    "ac_sd_layer_norm_f16": (_I, [_V, _I, _I, _V, _V, _F, _V, _V]),
    "ac_sd_geglu_f16": (_I, [_V, _I64, _I, _V, _V]),
    "ac_sd_softmax_f16": (_I, [_V, _I64, _I, _I64, _I64, _F, _V, _V]),
    "ac_sd_cast_f16": (_I, [_V, _I64, _V, _V]),
    "ac_sd_group_norm_backward": (_I, [_V, _V, _I, _I, _I, _I, _V, _V, _V, _I, _V, _V, _V, _V, _V]),
    "ac_sd_softmax_backward_f16": (_I, [_V, _V, _I64, _I, _I64, _F, _V, _V]),
    "ac_sd_transpose_f16": (_I, [_V, _I, _I, _I64, _V, _I64, _V]),
    "ac_sd_conv_s2_dgrad_operand_f16": (_I, [_V, _I, _I, _I, _I, _I, _I, _V, _V]),
    "ac_nsr_debug_tc_layer": (_I, [_V, _V, _V, _V]),
    "ac_nsr_upsample_round": (_I, [_V, _V, _V, _V, _U32, _U32, _F, _V, _V, _V, _V, _V]),
    "ac_nsr_debug_upsample": (_I, [_V, _V, _V, _V, _U32, _U32, _F, _V, _V, _V, _V, _V, _V, _V]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU or eager fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = the .so is stale w.r.t. the header
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    """Error convention of the reference ops: failures surface as RuntimeError
    (TORCH_CHECK / std::runtime_error, encoder/hashencoder/src/hashencoder.cu:16-19,349,364)."""
    if rc == AC_OK:
        return
    reason = {AC_E_INVALID_ARG: "invalid argument", AC_E_UNSUPPORTED: "unsupported D/C combination",
              AC_E_CUDA: "CUDA error: " + lib().ac_last_cuda_error().decode(),
              AC_E_WORKSPACE: "workspace too small"}.get(rc, f"error {rc}")
    raise RuntimeError(f"{what}: {reason}")


def ptr(t):
    """Device pointer of a CUDA fp32/int32 tensor (None -> NULL), with the reference's checks
    (CHECK_CUDA / CHECK_CONTIGUOUS, hashencoder.cu:414-430)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("tensor must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
