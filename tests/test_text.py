"""Detokeniser parity: our BPE-suffix decoder against the `tokenizers` library's decode of the
reference's tokenizer.json (golden fixture tests/golden/text_detok.json; live check when the reference is mounted)."""
import json
import os

import pytest
import torch

from asr_transformer_b200.text import Detokenizer

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, "golden", "text_detok.json"), encoding="utf-8"))


def make():
    return Detokenizer(FX["vocab"], FX["special_ids"], FX["suffix"])


def test_golden_cases():
    d = make()
    assert len(FX["vocab"]) == 250 and FX["suffix"] == "[EOF]"
    for c in FX["cases"]:
        assert d.decode(c["ids"]) == c["text"], c["ids"]
        assert d.decode(c["ids"], skip_special_tokens=False) == c["text_with_special"], c["ids"]


def test_decode_batch_cuts_at_eos_and_lengths():
    d = make()
    rows = torch.tensor([[1, 244, 247, 2, 9, 9], [1, 5, 249, 6, 7, 8]], dtype=torch.int32)
    n = torch.tensor([4, 6], dtype=torch.int32)
    assert d.decode_batch(rows, n) == [d.decode([1, 244, 247, 2]), d.decode([1, 5, 249, 6, 7, 8])]
    assert d.decode_batch(rows) == [d.decode([1, 244, 247, 2]), d.decode([1, 5, 249, 6, 7, 8])]   # EOS found / absent
    assert d.decode([1, 244, 2]) == "что" and d.decode([244, 247]) == "что это"


@pytest.mark.skipif(not os.path.exists("/root/reference/tokenizer.json"), reason="reference not mounted")
def test_live_against_tokenizers_library():
    tokenizers = pytest.importorskip("tokenizers")
    tok = tokenizers.Tokenizer.from_file("/root/reference/tokenizer.json")
    d = Detokenizer.from_tokenizer_json("/root/reference/tokenizer.json")
    g = torch.Generator().manual_seed(3)
    for n in (0, 1, 3, 40, 129):
        ids = torch.randint(0, 250, (n,), generator=g).tolist()
        assert d.decode(ids) == tok.decode(ids) and d.decode(ids, False) == tok.decode(ids, skip_special_tokens=False)
