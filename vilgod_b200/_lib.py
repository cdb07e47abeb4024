"""ctypes binding of libvilgod_b200.so (C ABI declared in include/vilgod_b200.h).

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
import of anything that needs it raises, and every entry point needs an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libvilgod_b200.so")
LIB_PATHS = {"f16": LIB_PATH, "bf16": os.path.join(HERE, "lib", "libvilgod_b200_bf16.so")}

VG_ABI_VERSION = 2
VG_MAX_VIEWS = 16
VG_VIT_LAYERS = 12
VG_TILE_ELEMS = 196 * 256

VG_OK, VG_EINVAL, VG_ESHAPE, VG_EWORKSPACE, VG_EDEGENERATE, VG_ECUDA, VG_ESTATE = 0, -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "VG_OK", -1: "VG_EINVAL", -2: "VG_ESHAPE", -3: "VG_EWORKSPACE",
                -4: "VG_EDEGENERATE", -5: "VG_ECUDA", -6: "VG_ESTATE"}
VG_ROTATE_TORCH_CPU, VG_ROTATE_FUSED, VG_ROTATE_UNFUSED = 0, 1, 2
VG_DIV_TRUE, VG_DIV_RECIPROCAL = 0, 1
VG_EPI_BIAS_BF16, VG_EPI_BIAS_QGELU_BF16, VG_EPI_BIAS_RESID_F32 = 0, 1, 2

fp = C.POINTER(C.c_float)


class VgConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("resolution", C.c_int32), ("depth", C.c_int32),
                ("image_size", C.c_int32), ("num_views", C.c_int32), ("rotate_mode", C.c_int32),
                ("div_mode", C.c_int32), ("pool_kernel", C.c_int32), ("pool_pad", C.c_int32),
                ("obj_ratio", C.c_double), ("depth_bias", C.c_double), ("logit_scale", C.c_double),
                ("rot", (C.c_float * 9) * VG_MAX_VIEWS), ("gauss", C.c_float * 9)]


class VgVitLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln_1_weight", "ln_1_bias", "attn_in_proj_weight", "attn_in_proj_bias",
        "attn_out_proj_weight", "attn_out_proj_bias", "ln_2_weight", "ln_2_bias",
        "mlp_c_fc_weight", "mlp_c_fc_bias", "mlp_c_proj_weight", "mlp_c_proj_bias")]


class VgVitWeights(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("conv1_weight", "class_embedding",
                                           "positional_embedding", "ln_pre_weight", "ln_pre_bias")]
                + [("layers", VgVitLayerWeights * VG_VIT_LAYERS)]
                + [(n, C.c_void_p) for n in ("ln_post_weight", "ln_post_bias", "proj")])


class VgProjectDebug(C.Structure):
    _fields_ = [("d_grid", C.c_void_p), ("d_densified", C.c_void_p)]


VG_K_NAMES = ["projection", "gemm_patch", "gemm_qkv", "gemm_out", "gemm_fc", "gemm_proj",
              "attention", "layernorm", "ln_pre", "head", "vote"]


class VgKernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * len(VG_K_NAMES)), ("launches", C.c_int64 * len(VG_K_NAMES)),
                ("work", C.c_double * len(VG_K_NAMES))]


class VgVitDebug(C.Structure):
    _fields_ = [("stop_after_layer", C.c_int32), ("d_x", C.c_void_p)]


# every symbol include/vilgod_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vg_abi_version": (C.c_int, []),
    "vg_operand_dtype": (C.c_int, []),
    "vg_create": (C.c_int, [C.POINTER(VgConfig), C.POINTER(C.c_void_p)]),
    "vg_destroy": (None, [C.c_void_p]),
    "vg_last_error": (C.c_char_p, [C.c_void_p]),
    "vg_launch_count": (C.c_int64, [C.c_void_p]),
    "vg_profile_begin": (C.c_int, [C.c_void_p]),
    "vg_profile_end": (C.c_int, [C.c_void_p, C.POINTER(VgKernelTimes)]),
    "vg_load_vit_weights": (C.c_int, [C.c_void_p, C.POINTER(VgVitWeights), C.c_void_p]),
    "vg_set_text_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                                       C.c_int32, C.c_void_p]),
    "vg_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "vg_canonicalise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "vg_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.POINTER(VgProjectDebug), C.c_void_p]),
    "vg_encode_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.POINTER(VgVitDebug), C.c_void_p]),
    "vg_vote": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                          C.c_void_p]),
    "vg_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_size_t, C.c_void_p]),
    "vg_test_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vg_test_gemm_lnf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p]),
    "vg_test_gemm_patch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p]),
    "vg_test_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vg_test_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
}

_libs = {}


def load(operand_dtype="f16"):
    """dlopen the fp16- (default) or bf16-operand build and bind every declared symbol (raises if missing)."""
    if operand_dtype not in LIB_PATHS:
        raise ValueError(f"operand_dtype must be one of {sorted(LIB_PATHS)}")
    if operand_dtype not in _libs:
        path = LIB_PATHS[operand_dtype]
        if operand_dtype == "f16" and os.environ.get("VG_LIB_PATH"):
            path = os.environ["VG_LIB_PATH"]      # instrumented debug build (python -m vilgod_b200.build --trace)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not built: run `python -m vilgod_b200.build` (needs nvcc, sm_100a). "
                "vilgod_b200 has no CPU fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.vg_abi_version() != VG_ABI_VERSION:
            raise RuntimeError(f"{os.path.basename(path)} ABI version mismatch; rebuild")
        if lib.vg_operand_dtype() != (1 if operand_dtype == "f16" else 0):
            raise RuntimeError(f"{os.path.basename(path)} was built for another operand dtype")
        _libs[operand_dtype] = lib
    return _libs[operand_dtype]


class VilgodError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {text}")
        self.code = code
