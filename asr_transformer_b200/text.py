"""Token ids -> text: the step after the hot path.

The reference trains a `tokenizers` BPE model with ``end_of_word_suffix="[EOF]"`` and decodes with
``BPEDecoder(suffix="[EOF]")`` (reference modules/tokenizer.py:9-19, tokenizer.json `decoder`).  That decoder is a pure
string rule, restated here without the `tokenizers` dependency: every token's ``[EOF]`` suffix becomes a space, except
in the last token of the sequence where it is dropped; special tokens ([MASK] [SOS] [EOS] [UNK] [PAD]) are skipped.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Sequence, Union


class Detokenizer:
    def __init__(self, vocab: Dict[str, int], special_ids: Iterable[int], suffix: str = "[EOF]"):
        self.id_to_token: Dict[int, str] = {int(i): t for t, i in vocab.items()}
        self.special = {int(i) for i in special_ids}
        self.suffix = suffix

    @classmethod
    def from_tokenizer_json(cls, path_or_dict: Union[str, dict]) -> "Detokenizer":
        """Build from a `tokenizers` JSON file (the reference's tokenizer.json) or its parsed dict."""
        d = path_or_dict if isinstance(path_or_dict, dict) else json.load(open(path_or_dict, encoding="utf-8"))
        model = d["model"]
        if model.get("type") != "BPE":
            raise ValueError("expected a BPE tokenizer.json, got %r" % model.get("type"))
        dec = d.get("decoder") or {}
        suffix = dec.get("suffix") or model.get("end_of_word_suffix") or ""
        special = [t["id"] for t in d.get("added_tokens", []) if t.get("special")]
        return cls(model["vocab"], special, suffix)

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True) -> str:
        toks: List[str] = []
        for i in ids:
            i = int(i)
            if skip_special_tokens and i in self.special:
                continue
            t = self.id_to_token.get(i)
            if t is not None:            # ids outside the vocabulary are dropped, as `tokenizers` does
                toks.append(t)
        if not self.suffix:
            return "".join(toks)
        n = len(toks)
        return "".join(t.replace(self.suffix, "" if k == n - 1 else " ") for k, t in enumerate(toks))

    def decode_batch(self, tokens, n_tokens=None, skip_special_tokens: bool = True) -> List[str]:
        """tokens: (B, L+1) ids as produced by ``Transformer.greedy_decode`` (column 0 = BOS); n_tokens: (B,) number of
        tokens up to and including the first EOS (``None``: cut at the first EOS found, else use the whole row)."""
        rows = tokens.tolist() if hasattr(tokens, "tolist") else [list(r) for r in tokens]
        lens: Optional[List[int]] = None if n_tokens is None else (
            n_tokens.tolist() if hasattr(n_tokens, "tolist") else list(n_tokens))
        out = []
        for b, row in enumerate(rows):
            if lens is not None:
                row = row[:int(lens[b])]
            else:
                eos = next((k for k, t in enumerate(row) if k > 0 and self.id_to_token.get(int(t)) == "[EOS]"), None)
                if eos is not None:
                    row = row[:eos + 1]
            out.append(self.decode(row, skip_special_tokens))
        return out
