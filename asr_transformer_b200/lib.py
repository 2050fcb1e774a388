"""ctypes binding of the C-ABI library (include/asr_b200.h).

There is no CPU path and no library fallback: if ``libasr_b200.so`` is missing or a tensor is not a CUDA tensor
the call fails loudly.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASR_B200_LIB: developer override (tools/variants.py builds experiment variants of the same library)
LIB_PATH = os.environ.get("ASR_B200_LIB") or os.path.join(_HERE, "libasr_b200.so")

c_void_p, c_int, c_float, c_size_t, c_longlong = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong


class AsrConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "vocab_size", "input_dim", "embedding_dim", "decoder_seq_len", "encoder_seq_len", "encoder_num_layers",
        "decoder_num_layers", "num_heads", "ff_dim", "pad_token_id", "eos_token_id", "bos_token_id")]


class AsrMhaWeights(C.Structure):
    _fields_ = [("w_qkv", c_void_p), ("b_qkv", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p)]


class AsrNormWeights(C.Structure):
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p)]


class AsrFfnWeights(C.Structure):
    _fields_ = [("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p)]


class AsrEncoderLayerWeights(C.Structure):
    _fields_ = [("norm1", AsrNormWeights), ("attn", AsrMhaWeights), ("norm2", AsrNormWeights), ("ffn", AsrFfnWeights)]


class AsrDecoderLayerWeights(C.Structure):
    _fields_ = [("norm1", AsrNormWeights), ("self_attn", AsrMhaWeights), ("norm2", AsrNormWeights),
                ("cross_attn", AsrMhaWeights), ("norm3", AsrNormWeights), ("ffn", AsrFfnWeights)]


class AsrWeights(C.Structure):
    _fields_ = [("conv1_w", c_void_p), ("conv1_b", c_void_p), ("conv2_wfrag", c_void_p), ("conv2_b", c_void_p),
                ("lin_in_w", c_void_p), ("lin_in_b", c_void_p), ("enc_pe", c_void_p),
                ("enc_layers", C.POINTER(AsrEncoderLayerWeights)), ("enc_norm_out", AsrNormWeights),
                ("embedding", c_void_p), ("dec_pe", c_void_p),
                ("dec_layers", C.POINTER(AsrDecoderLayerWeights)), ("dec_norm", AsrNormWeights),
                ("classifier_w", c_void_p), ("dec_image", c_void_p),
                ("dec_image_bytes", c_size_t)]


# name -> (restype, argtypes); mirrors include/asr_b200.h one to one (tests/test_host.py checks the symbol list against the header)
_SIGNATURES = {
    "asr_last_error": (C.c_char_p, []),
    "asr_version": (c_int, []),
    "asr_create": (c_int, [C.POINTER(AsrConfig), C.POINTER(c_void_p)]),
    "asr_destroy": (None, [c_void_p]),
    "asr_load_weights": (c_int, [c_void_p, C.POINTER(AsrWeights)]),
    "asr_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(c_size_t)]),
    "asr_encode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "asr_encoder_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "asr_decoder_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                    c_void_p, c_void_p]),
    "asr_beam_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, C.POINTER(c_size_t)]),
    "asr_decode_beam": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                c_void_p]),
    "asr_decode_greedy": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "asr_decode_prepare": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "asr_decode_run": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "asr_decode_profile": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "asr_launch_count": (C.c_ulonglong, []),
    "asr_split_operands": (c_int, []),
    "asr_decoder_image_bytes": (c_size_t, [C.POINTER(AsrConfig)]),
    "asr_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "asr_f32_to_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "asr_gemm_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_int, c_void_p]),
    "asr_gemm_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "asr_gemm_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                            c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "asr_attention": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong,
                              c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "asr_mha_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "asr_mha": (c_int, [c_void_p, c_void_p, C.POINTER(AsrMhaWeights), c_int, c_int, c_int, c_int, c_int, c_int,
                        c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "asr_ffn_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "asr_ffn": (c_int, [c_void_p, C.POINTER(AsrFfnWeights), c_int, c_int, c_int, c_void_p, c_size_t, c_void_p,
                        c_void_p]),
    "asr_conv_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "asr_conv_frontend": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                  c_size_t, c_void_p, c_void_p]),
    "asr_spectrogram": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "asr_embed_pe": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "asr_dec_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                               c_int, c_void_p, c_void_p]),
    "asr_dec_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_float,
                                  c_void_p, c_void_p]),
    "asr_umma_probe": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libasr_b200.so (built in-tree by ``python -m asr_transformer_b200.build``). Never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA library with `python -m asr_transformer_b200.build` "
                "(nvcc, sm_100a). There is no CPU or library fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().asr_last_error()
        raise RuntimeError(f"asr_b200 {what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None) -> Optional[int]:
    """Device pointer of a CUDA tensor (None -> NULL). CPU tensors are an error by design; ``dtype`` (when given) must
    match: the kernels read raw memory, a half-precision module would otherwise be read as fp32 silently."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("asr_b200 has no CPU path: expected a CUDA tensor, got device %s" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("asr_b200 expects contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"asr_b200 expects a {dtype} tensor here, got {t.dtype}")
    return t.data_ptr()


def on(device) -> "torch.cuda.device":
    """Context manager that makes ``device`` (a torch.device or a CUDA tensor) the current CUDA device for the C calls
    inside it: kernels are launched on torch's current stream OF THAT DEVICE, and the library's per-device kernel
    configuration applies to the device the pointers live on (a model on cuda:1 works while cuda:0 is current)."""
    if isinstance(device, torch.Tensor):
        device = device.device
    if device.type != "cuda":
        raise RuntimeError("asr_b200 has no CPU path: expected a CUDA device, got %s" % device)
    return torch.cuda.device(device)


def stream() -> int:
    """torch's current stream on the current device (call inside ``on(device)``)."""
    return torch.cuda.current_stream().cuda_stream


_workspaces = {}


def workspace(nbytes: int, device: torch.device, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream, tag); contents are undefined between calls.  Keyed by the stream the
    caller is on, so two streams never share scratch; a buffer that is outgrown stays alive until the work already
    queued on that stream has finished with it (record_stream)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = torch.cuda.current_stream(idx)
    key = (idx, st.cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            buf.record_stream(st)
        buf = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=torch.device("cuda", idx))
        _workspaces[key] = buf
    return buf
