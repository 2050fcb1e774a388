"""CPU checks of the cluster decoder's packed weight image (engine.pack_dec_image / csrc/decode_cluster.cu).

The kernel cannot run here, so its index arithmetic is restated lane by lane in numpy: the fragment-major image is read
exactly as mm_stream reads it (one 16-byte word per lane and plane, K permutation shared by the A and B operands), fed
through the PTX-documented mma.m16n8k16 register layout, and the result tiles are interpreted as the epilogues do.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from asr_transformer_b200 import engine as E
from oracle import speech_transformer as O
from tests.util import build_model


def mma_m16n8k16(a_regs, b_regs):
    """a_regs [32 lanes][4 regs][2], b_regs [32][2][2] -> c [32][4] following the PTX fragment layout."""
    A = np.zeros((16, 16))
    Bm = np.zeros((16, 8))
    for lane in range(32):
        g, tg = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * tg + e] = a_regs[lane, 0, e]
            A[g + 8, 2 * tg + e] = a_regs[lane, 1, e]
            A[g, 2 * tg + 8 + e] = a_regs[lane, 2, e]
            A[g + 8, 2 * tg + 8 + e] = a_regs[lane, 3, e]
            Bm[2 * tg + e, g] = b_regs[lane, 0, e]
            Bm[2 * tg + 8 + e, g] = b_regs[lane, 1, e]
    Cm = A @ Bm
    c = np.zeros((32, 4))
    for lane in range(32):
        g, tg = lane >> 2, lane & 3
        c[lane] = [Cm[g, 2 * tg], Cm[g, 2 * tg + 1], Cm[g + 8, 2 * tg], Cm[g + 8, 2 * tg + 1]]
    return c


def emulate_mm_stream(image, R, K, x):
    """image: flat float array in pack_mma_a order; x [8, K] -> out [R, 8] computed the way the kernel does."""
    MT, KB = R // 16, K // 32
    out = np.zeros((R, 8))
    for mt in range(MT):
        acc = np.zeros((32, 4))
        for kb in range(KB):
            blk = image[(kb * MT + mt) * 512:(kb * MT + mt + 1) * 512].reshape(2, 32, 8)   # k-tile s, lane, 8 elements
            a_regs = np.zeros((2, 32, 4, 2))
            b_regs = np.zeros((2, 32, 2, 2))
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                xw = x[g, kb * 32 + tg * 8: kb * 32 + tg * 8 + 8]          # one 16-byte word of the bf16 x row
                for s in range(2):
                    wd = blk[s, lane]                                       # one LDS.128 = {a0, a1, a2, a3}
                    a_regs[s, lane] = [wd[0:2], wd[2:4], wd[4:6], wd[6:8]]
                    b_regs[s, lane] = [xw[4 * s:4 * s + 2], xw[4 * s + 2:4 * s + 4]]   # (b.x, b.y) / (b.z, b.w)
            acc += mma_m16n8k16(a_regs[0], b_regs[0]) + mma_m16n8k16(a_regs[1], b_regs[1])
        for lane in range(32):                                             # epilogue interpretation of the float4 tile
            g, tg = lane >> 2, lane & 3
            out[16 * mt + 2 * g, 2 * tg] = acc[lane, 0]                      # MMA row g     = matrix row 2g of the tile
            out[16 * mt + 2 * g, 2 * tg + 1] = acc[lane, 1]
            out[16 * mt + 2 * g + 1, 2 * tg] = acc[lane, 2]                  # MMA row g + 8 = matrix row 2g + 1
            out[16 * mt + 2 * g + 1, 2 * tg + 1] = acc[lane, 3]
    return out


@pytest.mark.parametrize("R,K", [(192, 128), (64, 256), (32, 64), (128, 32)])
def test_fragment_order_matches_kernel_indexing(R, K):
    g = torch.Generator().manual_seed(R * 1000 + K)
    w = torch.randn(R, K, generator=g).to(torch.float16)
    x = torch.randn(8, K, generator=g).double().numpy()
    img = E.pack_mma_a(w).double().numpy()
    got = emulate_mm_stream(img, R, K, x)
    ref = w.double().numpy() @ x.T
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


def unpack_mma_a(flat, R, K):
    t = flat.reshape(K // 32, R // 16, 2, 8, 4, 2, 2, 2)      # kb, mt, s, g, tg, pair, p, e
    return t.permute(1, 3, 6, 0, 4, 2, 5, 7).reshape(R, K)    # mt, g, p, kb, tg, s, pair, e


@pytest.mark.parametrize("name", ["T0", "C2", "C5"])
def test_image_layout_and_contents(name):
    cfg = O.CONFIGS[name]
    m = build_model(cfg)
    dec = m.decoder
    D, H, FF, V, nd = cfg.embedding_dim, cfg.num_heads, cfg.ff_dim, cfg.vocab_size, cfg.decoder_num_layers
    lay = E.dec_image_layout(D, H, FF, V, nd)
    img = E.pack_dec_image(dec)
    assert img.dtype == torch.uint8 and img.numel() == lay["total_bytes"]
    FFS, VS = lay["FFS"], lay["VS"]

    def mat(r, l, key, R, K):
        off = r * lay["rank_bytes"] + (l * lay["layer_bytes"] + lay["off_" + key] if l >= 0 else lay["off_cls"])
        return unpack_mma_a(img[off:off + R * K * 2].view(torch.float16), R, K)

    for r in (0, H - 1):
        for l in (0, nd - 1):
            layer = dec._layers[l]
            sa = E.pack_mha(layer._mask_attention)
            ca = E.pack_mha(layer._cross_attention)
            ff = E.pack_ffn(layer._feedforward)
            rows = torch.cat([torch.arange(r * 64, (r + 1) * 64) + k * D for k in range(3)])
            assert torch.equal(mat(r, l, "qkv", 192, D), sa["w_qkv"][rows])
            assert torch.equal(mat(r, l, "wo", D, 64), sa["w_out"][:, r * 64:(r + 1) * 64])
            assert torch.equal(mat(r, l, "wqc", 64, D), ca["w_qkv"][r * 64:(r + 1) * 64])
            assert torch.equal(mat(r, l, "woc", D, 64), ca["w_out"][:, r * 64:(r + 1) * 64])
            assert torch.equal(mat(r, l, "w1", FFS, D), ff["w1"][r * FFS:(r + 1) * FFS])
            assert torch.equal(mat(r, l, "w2", D, FFS), ff["w2"][:, r * FFS:(r + 1) * FFS])
            off = r * lay["rank_bytes"] + l * lay["layer_bytes"] + lay["off_small"]
            small = img[off:off + lay["small_floats"] * 4].view(torch.float32)
            assert torch.equal(small[:192], sa["b_qkv"][rows])
            assert torch.equal(small[192:256], ca["b_qkv"][r * 64:(r + 1) * 64])
            assert torch.equal(small[256:256 + FFS], ff["b1"][r * FFS:(r + 1) * FFS])
            assert torch.equal(small[256 + FFS:256 + FFS + D], sa["b_out"])
            assert torch.equal(small[256 + FFS + 2 * D:256 + FFS + 3 * D], ff["b2"])
            assert torch.equal(small[256 + FFS + 7 * D:256 + FFS + 8 * D], layer._norm3.weight.detach())
            nxt = small[256 + FFS + 9 * D:256 + FFS + 11 * D]          # norm1 of the next layer rides along
            if l + 1 < nd:
                assert torch.equal(nxt[:D], dec._layers[l + 1]._norm1.weight.detach())
                assert torch.equal(nxt[D:], dec._layers[l + 1]._norm1.bias.detach())
            else:
                assert not nxt.any()
        cls = torch.zeros(H * VS, D, dtype=torch.float16)
        cls[:V] = dec._classifier.weight.detach().to(torch.float16)
        assert torch.equal(mat(r, -1, "cls", VS, D), cls[r * VS:(r + 1) * VS])


def test_layout_agrees_with_the_library():
    from asr_transformer_b200 import lib as L
    lib = L.load()
    for name in ("T0", "C2", "C4", "C5"):
        cfg = O.CONFIGS[name]
        c = L.AsrConfig(vocab_size=cfg.vocab_size, input_dim=cfg.input_dim, embedding_dim=cfg.embedding_dim,
                        decoder_seq_len=cfg.decoder_seq_len, encoder_seq_len=cfg.encoder_seq_len,
                        encoder_num_layers=cfg.encoder_num_layers, decoder_num_layers=cfg.decoder_num_layers,
                        num_heads=cfg.num_heads, ff_dim=cfg.ff_dim, pad_token_id=4, eos_token_id=2, bos_token_id=1)
        lay = E.dec_image_layout(cfg.embedding_dim, cfg.num_heads, cfg.ff_dim, cfg.vocab_size, cfg.decoder_num_layers)
        assert lib.asr_decoder_image_bytes(C.byref(c)) == lay["total_bytes"]
    # unsupported: 3 heads
    c = L.AsrConfig(vocab_size=250, input_dim=80, embedding_dim=192, decoder_seq_len=8, encoder_seq_len=8,
                    encoder_num_layers=1, decoder_num_layers=1, num_heads=3, ff_dim=96)
    assert lib.asr_decoder_image_bytes(C.byref(c)) == 0 and E.dec_image_layout(192, 3, 96, 250, 1) is None
