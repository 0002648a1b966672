"""ctypes binding of libwae_b200.so (include/wae_b200.h).

There is deliberately no fallback here: if the shared library is missing, or a call reports an
error (including "device is not sm_100"), a WaeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libwae_b200.so"

WAE_MAX_LAYERS = 64

AR_SAMPLE_CATEGORICAL = 0
AR_SAMPLE_NONE = 1
AR_SAMPLE_MOL = 2
AR_SAMPLE_GAUSS = 3


class WaeError(RuntimeError):
    pass


class StackDims(C.Structure):
    _fields_ = [
        ("layers", C.c_int32), ("kernel_size", C.c_int32), ("R", C.c_int32), ("G", C.c_int32),
        ("S", C.c_int32), ("C", C.c_int32), ("Gi", C.c_int32), ("O", C.c_int32), ("Oin", C.c_int32),
        ("dilation", C.c_int32 * WAE_MAX_LAYERS),
    ]


class StackF32(C.Structure):
    _fields_ = [("d", StackDims)] + [(n, C.c_void_p) for n in
                                    ("wf", "bf", "w1", "b1", "wg", "w2", "b2", "w3", "b3", "w4", "b4")]


class StackBF16(C.Structure):
    _fields_ = [("d", StackDims)] + [(n, C.c_void_p) for n in
                                    ("w1", "wo", "ws", "w3", "w4", "b1", "wg", "bo", "bs_sum", "b3", "b4", "wf", "bf", "wfb")]


class StackSaved(C.Structure):
    _fields_ = [("x_all", C.c_void_p), ("h_all", C.c_void_p), ("c_cl", C.c_void_p), ("r1", C.c_void_p), ("r2", C.c_void_p),
                ("gate", C.c_void_p)]


class StackBwd(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wdh", "wdx", "wct", "w4t", "w3t", "x_all", "h_all", "c_cl", "r1", "r2", "gemb",
                                           "dw1", "dwo", "dws", "dw3", "dw4", "dgb", "dbo", "dbs", "db3", "db4", "dc", "dx0", "dy", "gate")]


class CondFrontend(C.Structure):
    _fields_ = [("conv_in_w_t", C.c_void_p), ("n_stages", C.c_int32), ("scale", C.c_int32 * 8), ("filter", C.c_void_p * 8),
                ("speaker_ids", C.c_void_p), ("speaker_table", C.c_void_p), ("n_speakers", C.c_int32), ("coef", C.c_void_p)]


class EncLayer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p)] + [(n, C.c_int32) for n in ("cin", "cout", "k", "stride", "relu", "residual")]


class Encoder(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("layer", EncLayer * 16), ("lin_w_t", C.c_void_p), ("lin_b", C.c_void_p),
                ("hid", C.c_int32), ("D", C.c_int32)]


class VqSlice(C.Structure):
    _fields_ = [("codebook", C.c_void_p), ("K", C.c_int32), ("d0", C.c_int32), ("sub_d", C.c_int32),
                ("idx_out", C.c_void_p), ("counts_out", C.c_void_p), ("sqerr_out", C.c_void_p)]


class ArPost(C.Structure):
    _fields_ = [("table", C.c_void_p), ("mu", C.c_int32), ("scalar_is_mulaw", C.c_int32), ("preemphasis_coef", C.c_float),
                ("gain", C.c_float), ("out_wave", C.c_void_p)]


class ArWeights(C.Structure):
    _fields_ = [
        ("d", StackDims),
        ("wtype", C.c_int32), ("cluster", C.c_int32), ("utts_per_cluster", C.c_int32), ("reserved", C.c_int32),
        ("blob", C.c_void_p), ("layer_off", C.c_void_p),
    ] + [(n, C.c_void_p) for n in ("b1", "wg", "bo", "bs", "b3", "b4", "wf", "bf")]


_lib = None

# name -> (restype, argtypes); also the list of symbols tests check against include/wae_b200.h
SIGNATURES = {
    "wae_version": (C.c_int, []),
    "wae_last_error": (C.c_char_p, []),
    "wae_device_check": (C.c_int, [C.c_int]),
    "wae_launch_count": (C.c_int64, []),
    "wae_vq_search": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wae_vq_set_variant": (C.c_int, [C.c_int]),
    "wae_synth_postprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "wae_vq_ema_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "wae_vq_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]),
    "wae_upsample_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wae_upsample_stage_backward_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "wae_upsample_stage_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_enc_layer_forward_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_enc_layer_backward_input": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_enc_layer_backward_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wae_conv1d_relu_res": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_stack_workspace_f32": (C.c_size_t, [C.POINTER(StackDims), C.c_int, C.c_int]),
    "wae_stack_forward_f32": (C.c_int, [C.POINTER(StackF32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_stack_workspace_bf16": (C.c_size_t, [C.POINTER(StackDims), C.c_int, C.c_int]),
    "wae_stack_forward_bf16": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_stack_forward_bf16_up": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_stack_forward_bf16_save": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                              C.c_void_p, C.POINTER(StackSaved), C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_train_transpose_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_train_im2col": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_train_gate_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "wae_train_dx_accum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "wae_stack_forward_bf16_idx": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_stack_nll_bf16_idx": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_stack_forward_bf16_lat": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(CondFrontend),
                                             C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p]),
    "wae_encoder_vq_supported": (C.c_int, [C.POINTER(Encoder)]),
    "wae_encoder_vq_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "wae_encoder_vq_profile": (C.c_int, [C.c_void_p]),
    "wae_encoder_vq_forward": (C.c_int, [C.POINTER(Encoder), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(VqSlice), C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_dump_text": (C.c_int, [C.c_char_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int]),
    "wae_nll_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_sumsq": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "wae_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "wae_adam_step_dlr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "wae_set_layer_cluster": (C.c_int, [C.c_int]),
    "wae_layer_kernel_name": (C.c_char_p, []),
    "wae_set_head_pair": (C.c_int, [C.c_int]),
    "wae_layer_set_profile_buffer": (None, [C.c_void_p]),
    "wae_profile_enable": (None, [C.c_int]),
    "wae_profile_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "wae_gemm_bf16_tn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wae_stack_gate_save_supported": (C.c_int, [C.POINTER(StackDims)]),
    "wae_stack_forward_bf16_save_idx": (C.c_int, [C.POINTER(StackBF16), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                  C.POINTER(StackSaved), C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_onehot_bf16": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_stack_backward_workspace_bf16": (C.c_size_t, [C.POINTER(StackDims), C.c_int, C.c_int]),
    "wae_stack_backward_bf16": (C.c_int, [C.POINTER(StackBF16), C.POINTER(StackBwd), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "wae_stack_backward_workspace_bf16_2s": (C.c_size_t, [C.POINTER(StackDims), C.c_int, C.c_int]),
    "wae_stack_backward_bf16_2s": (C.c_int, [C.POINTER(StackBF16), C.POINTER(StackBwd), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "wae_train_ce_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "wae_colsum_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wae_gemm_bf16_nt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wae_gemm_bf16_tn_bf16out": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "wae_ar_workspace": (C.c_size_t, [C.POINTER(ArWeights), C.c_int, C.c_int]),
    "wae_ar_set_profile_buffer": (None, [C.c_void_p]),
    "wae_ar_generate_wave": (C.c_int, [C.POINTER(ArWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(ArPost),
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    "wae_ar_generate": (C.c_int, [C.POINTER(ArWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load libwae_b200.so (once). Raises WaeError if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise WaeError(
                f"{LIB_PATH} not found: build it with `python -m wavenet_autoencoders_b200.build` "
                "(this package has no CPU / PyTorch fallback for its CUDA kernels)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().wae_last_error().decode(errors="replace")
        raise WaeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL). The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "libwae_b200 takes contiguous tensors"
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib().wae_launch_count())
