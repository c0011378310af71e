"""ctypes binding of libmsq_b200.so (the C ABI declared in include/msq_b200.h).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked without a
GPU), but every compute call needs a CUDA device and raises RuntimeError otherwise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsq_b200.so")


class MsqConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden", "layers", "heads", "inter", "vocab", "max_pos", "type_vocab", "vit_width", "vit_layers",
        "vit_patch", "vit_res", "para_heads", "para_ff", "para_layers", "precise", "reserved", "rn_width",
        "rn_blocks0", "rn_blocks1", "rn_blocks2", "rn_blocks3", "rn_embed", "reserved2", "reserved3")]


class MsqEncodeOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "sents", "para", "h0", "key", "cls", "cls_mat", "cls_score", "score_mat", "his1", "his2", "top_vec")]


_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "msq_last_error": (C.c_char_p, []),
    "msq_version": (C.c_int, []),
    "msq_launch_count": (_I64, []),
    "msq_tc_available": (C.c_int, []),
    "msq_profile_enable": (C.c_int, [_I32]),
    "msq_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I64)]),
    "msq_model_create": (C.c_int, [C.POINTER(MsqConfig), C.POINTER(_P)]),
    "msq_model_destroy": (None, [_P]),
    "msq_model_set_weight": (C.c_int, [_P, C.c_char_p, _P, _I64, _P]),
    "msq_model_pack": (C.c_int, [_P, _P]),
    "msq_model_update_weight": (C.c_int, [_P, C.c_char_p, _P, _I64, _P]),
    "msq_model_refresh": (C.c_int, [_P, _P]),
    "msq_vit_forward": (C.c_int, [_P, _P, _I64, _P, _I64, _P, _P]),
    "msq_inner_forward": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _I64, _P, _P, _P, _P, _P]),
    "msq_encode": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, C.POINTER(MsqEncodeOut), _P]),
    "msq_training_loss": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _P, _P, _F, _P, _P, _P]),
    "msq_beam_search": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "msq_decode_step": (C.c_int, [_P] * 12 + [_I32, _I32, _P, _P, _P, _P]),
    "msq_pointer_p1": (C.c_int, [_P] * 10 + [_I64, _I32, _I32, _I32, _P, _P, _P, _P]),
    "msq_order_manuals_dev": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _I32, _P, _P]),
    "msq_order_manuals_host": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _I32, _P, _P]),
    "msq_scan_steps": (C.c_int, [_P, _I64, _I32, _I32, _I64, _I64, _P, _P, _P, C.POINTER(_I32), _P]),
    "msq_expand_pairs": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "msq_order_manuals_raw_host": (C.c_int, [_P, _P, _I64, _I32, _I32, _I64, _I64, _I64, _P, _I32, _P, _P]),
    "msq_train_param_count": (_I64, [_P, _P]),
    "msq_train_grad_numel": (_I64, [_P, _P]),
    "msq_train_param_info": (C.c_int, [_P, _I64, C.POINTER(C.c_char_p), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)]),
    "msq_train_read_param": (C.c_int, [_P, C.c_char_p, _P, _I64, _P]),
    "msq_inner_forward_train": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P, _I64, _P, _P, _P, _P]),
    "msq_inner_backward": (C.c_int, [_P, _P, _P, _P, _P]),
    "msq_train_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _P, _P, _F, _P, _P, _P]),
    "msq_train_set_dropout": (C.c_int, [_P, _F, _F, _F, C.c_uint32, _P]),
    "msq_train_set_bn_mode": (C.c_int, [_P, _I32, _P]),
    "msq_train_set_triplets": (C.c_int, [_P, _P, _I64, _F, _P]),
    "msq_train_set_multimodal_loss": (C.c_int, [_P, _I32, _P]),
    "msq_train_dropout_step": (_I64, [_P]),
    "msq_train_ready_count": (_I64, [_P]),
    "msq_train_ready_info": (C.c_int, [_P, _I64, C.POINTER(_I64), C.POINTER(_I64)]),
    "msq_train_ready_wait": (C.c_int, [_P, _I64, _P]),
    "msq_adamw_step": (C.c_int, [_P, _P, _F, _F, _F, _F, _F, _F, _F, _P, _P]),
    "msq_gemm": (C.c_int, [_I32, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P]),
    "msq_gemm_deferred_ln": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _F, _P, _P, _P, _I64, _I32, _I32, _I32, _P]),
    "msq_gemm_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _I64, _I32, _I32, _I32, _P]),
    "msq_layernorm": (C.c_int, [_I32, _P, _I64, _I32, _P, _P, _F, _P, _P]),
    "msq_attention": (C.c_int, [_I32, _P, _I64, _I32, _I32, _F, _P, _I32, _P, _P]),
    "msq_f32_to_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "msq_f32_to_bf16_split": (C.c_int, [_P, _P, _I64, _I32, _P]),
}

_lib = None


def load():
    """Load the shared library (building it first if a toolchain is present and sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.environ.get("MSQ_REBUILD") == "1":
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmsq_b200.so is missing: run `python -m multimodal_sequencing_b200.build` "
                           "(the CUDA extension is mandatory; there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("msq_b200: " + load().msq_last_error().decode(errors="replace"))
