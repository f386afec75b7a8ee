"""ctypes binding of libmsd_b200.so (include/msd.h).  There is no fallback: if the library is missing or the
device is not sm_100, every product call raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmsd_b200.so"

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t
c_u64, c_u32, c_i64 = C.c_uint64, C.c_uint32, C.c_int64


class MsdError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [
        ("in_channels", c_int), ("image_size", c_int), ("patch_size", c_int), ("embed_dim", c_int),
        ("depth", c_int), ("num_heads", c_int), ("mlp_hidden", c_int), ("dec_channels", c_int * 4),
        ("out_channels", c_int), ("ln_eps_block", c_float), ("ln_eps_outer", c_float), ("gemm_backend", c_int),
    ]


class BlockWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
        "w_qkv_ln", "s_qkv", "b_qkv_ln", "w_fc1_ln", "s_fc1", "b_fc1_ln")]


class StageWeights(C.Structure):
    _fields_ = [("w_up", c_void_p), ("b_up", c_void_p), ("w_conv", c_void_p * 4), ("b_conv", c_void_p * 4)]


class ModelWeights(C.Structure):
    _fields_ = [
        ("w_pe", c_void_p), ("b_pe", c_void_p), ("pe_ln_g", c_void_p), ("pe_ln_b", c_void_p), ("pos_embed", c_void_p),
        ("blocks", C.POINTER(BlockWeights)), ("norm_g", c_void_p), ("norm_b", c_void_p),
        ("stages", StageWeights * 4), ("w_head3", c_void_p), ("b_head3", c_void_p), ("w_head1", c_void_p),
        ("b_head1", c_void_p),
    ]


class BnParams(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("gamma", "beta", "running_mean", "running_var")]


class TrainBlock(C.Structure):
    _fields_ = [("wT_qkv", c_void_p), ("wT_proj", c_void_p), ("wT_fc1", c_void_p), ("wT_fc2", c_void_p), ("trainable", c_int)]


class TrainStage(C.Structure):
    _fields_ = [("w_up", c_void_p), ("b_up", c_void_p), ("wT_up", c_void_p), ("bn_up", BnParams), ("w_conv", c_void_p * 4),
                ("wd_conv", c_void_p * 4), ("bn_conv", BnParams * 4)]


class TrainWeights(C.Structure):
    _fields_ = [("enc", ModelWeights), ("blocks_t", C.POINTER(TrainBlock)), ("stages", TrainStage * 4), ("w_head3", c_void_p),
                ("b_head3", c_void_p), ("wd_head3", c_void_p), ("bn_head", BnParams), ("w_head1", c_void_p), ("b_head1", c_void_p),
                ("bn_eps", c_float), ("bn_momentum", c_float)]


class BlockGrads(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


class StageGrads(C.Structure):
    _fields_ = [("w_up", c_void_p), ("bn_up_g", c_void_p), ("bn_up_b", c_void_p), ("w_conv", c_void_p * 4), ("bn_g", c_void_p * 4),
                ("bn_b", c_void_p * 4)]


class ModelGrads(C.Structure):
    _fields_ = [("w_pe", c_void_p), ("b_pe", c_void_p), ("pe_ln_g", c_void_p), ("pe_ln_b", c_void_p), ("pos_embed", c_void_p),
                ("blocks", C.POINTER(BlockGrads)), ("norm_g", c_void_p), ("norm_b", c_void_p), ("stages", StageGrads * 4),
                ("w_head3", c_void_p), ("bn_head_g", c_void_p), ("bn_head_b", c_void_p), ("w_head1", c_void_p), ("b_head1", c_void_p)]


class AdamWTensor(C.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("numel", C.c_longlong),
                ("lr", c_float), ("weight_decay", c_float), ("step", c_int)]


class RepackOp(C.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("kind", c_int), ("d0", c_int), ("d1", c_int), ("kpad", c_int), ("flag", c_int)]


REPACK_CAST, REPACK_TRANSPOSE, REPACK_CONV3, REPACK_CONV3_DGRAD, REPACK_CONVT, REPACK_CONVT_T, REPACK_REPEAT_F32 = range(7)

# name -> (restype, argtypes); must list every symbol include/msd.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "msd_version": (c_int, []),
    "msd_last_error": (C.c_char_p, []),
    "msd_check_device": (c_int, []),
    "msd_philox_offset_increment": (c_u64, [c_i64, c_u32]),
    "msd_philox_normal_f32": (c_int, [c_void_p, c_i64, c_u64, c_u64, c_u32, c_void_p]),
    "msd_philox_uniform_f32": (c_int, [c_void_p, c_i64, c_u64, c_u64, c_u32, c_void_p]),
    "msd_corrupt_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_float, c_float, c_float, c_float, c_int, c_u64, c_u64, c_u32, c_u32, c_u32,
                                C.POINTER(c_u64), c_void_p]),
    "msd_metrics_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "msd_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "msd_loss_fwd_bwd_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "msd_model_create": (c_int, [C.POINTER(ModelDesc), C.POINTER(c_void_p)]),
    "msd_model_set_weights": (c_int, [c_void_p, C.POINTER(ModelWeights)]),
    "msd_model_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "msd_model_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "msd_model_launch_count": (c_int, [c_void_p]),
    "msd_model_profile_enable": (c_int, [c_void_p, c_int]),
    "msd_model_profile_read": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "msd_model_destroy": (None, [c_void_p]),
    "msd_linear_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_void_p]),
    "msd_conv3x3_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "msd_conv3x3_head_nchw_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                          c_int, c_void_p]),
    "msd_convt2x2_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_void_p]),
    "msd_layernorm_f32_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "msd_layernorm_pos_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msd_ln_stats_parts": (c_int, [c_int, c_int]),
    "msd_layernorm_pos_stats_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                            c_int, c_int, c_void_p]),
    "msd_linear_resid_stats_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                           c_void_p]),
    "msd_linear_lnfold_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_int,
                                       c_int, c_int, c_void_p]),
    "msd_linear_gelu_pre_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msd_linear_dgelu_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msd_attention_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msd_attention_tc5_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msd_patchify_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_head1x1_nchw_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    "msd_wgrad_linear_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, C.c_longlong, c_int, c_int,
                                      c_void_p]),
    "msd_colsum_bf16": (c_int, [c_void_p, C.c_longlong, c_int, c_void_p, c_void_p]),
    "msd_f64_to_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "msd_bn_stats_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "msd_bn_sums_bf16": (c_int, [c_void_p, C.c_longlong, c_int, c_void_p, c_void_p]),
    "msd_bn_finalize": (c_int, [c_void_p, C.c_double, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "msd_conv3x3_stats_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                            c_int, C.POINTER(c_int), c_void_p]),
    "msd_bn_act_fwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_bn_act_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_gelu_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, C.c_longlong, c_void_p]),
    "msd_gelu_bwd_colsum_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "msd_layernorm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, c_void_p]),
    "msd_layernorm_bwd_colsum": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "msd_convt_unshuffle_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_head1x1_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_batch_sum_f32": (c_int, [c_void_p, c_void_p, c_int, C.c_longlong, c_void_p]),
    "msd_attention_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msd_gelu_fwd_bf16": (c_int, [c_void_p, c_void_p, C.c_longlong, c_void_p]),
    "msd_model_set_train_weights": (c_int, [c_void_p, C.POINTER(TrainWeights)]),
    "msd_model_train_layout": (c_int, [c_void_p, c_int, C.POINTER(c_size_t), c_int]),
    "msd_model_train_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "msd_model_forward_train": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "msd_model_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, C.POINTER(ModelGrads), c_void_p]),
    "msd_model_grad_groups": (c_int, [c_void_p]),
    "msd_model_grad_group_wait": (c_int, [c_void_p, c_int, c_void_p]),
    "msd_ingest_u16_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_export_u16": (c_int, [c_void_p, c_void_p, C.c_longlong, c_void_p]),
    "msd_tile_extract_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msd_tile_blend_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p]),
    "msd_grad_sqnorm_f32": (c_int, [C.POINTER(AdamWTensor), c_int, c_void_p, c_void_p]),
    "msd_adamw_step_f32": (c_int, [C.POINTER(AdamWTensor), c_int, c_float, c_float, c_float, c_int, c_float, c_void_p, c_void_p]),
    "msd_repack_weights": (c_int, [C.POINTER(RepackOp), c_int, c_void_p]),
    "msd_wgrad_conv3x3_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (built by build.py).  Raises MsdError when it is absent -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MsdError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(nvcc, sm_100a).  libmsd_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().msd_last_error().decode("utf-8", "replace")
        raise MsdError(f"libmsd_b200 error {rc}: {msg}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise MsdError("libmsd_b200 operates on CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise MsdError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


_device_ok: set[int] = set()


def require_device(device) -> None:
    """Check once per device that it is sm_100."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _device_ok:
        return
    with torch.cuda.device(idx):
        check(load().msd_check_device())
    _device_ok.add(idx)
