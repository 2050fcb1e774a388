"""asr_transformer_b200: B200-native (sm_100a) Speech-Transformer forward + greedy decode.

Drop-in for ``modules/Transformer`` of shockless/asr-transformer: same module classes, constructor
signatures, ``state_dict`` keys and ``forward`` / ``evaluate`` results; the arithmetic runs in hand-written
CUDA kernels (tcgen05/TMEM/TMA GEMM and flash attention, warp-shuffle LayerNorm, KV-cached decode step)
behind the C-ABI declared in ``include/asr_b200.h``.  No CPU path, no Triton, no library fallback.
"""
from .layers import MHA, MHAHead, FeedForward, TrainablePositionalEncoding, LayerNorm
from .model import Transformer, Encoder, Decoder, EncoderLayer, DecoderLayer, ConvFrontEnd
from .text import Detokenizer
from .frontend import Spectrogram

__all__ = ["Transformer", "Encoder", "Decoder", "EncoderLayer", "DecoderLayer", "ConvFrontEnd", "MHA", "MHAHead",
           "FeedForward", "TrainablePositionalEncoding", "LayerNorm", "Detokenizer", "Spectrogram"]
