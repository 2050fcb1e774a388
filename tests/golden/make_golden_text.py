#!/usr/bin/env python
"""Golden vectors for the detokeniser: the reference's tokenizer.json decoded by the `tokenizers` library itself.

Run in the build container (needs /root/reference and `tokenizers`): writes tests/golden/text_detok.json with the
vocabulary, the special ids and (ids -> text) pairs for random and hand-picked sequences."""
import json
import os
import random

from tokenizers import Tokenizer

REF = "/root/reference/tokenizer.json"
tok = Tokenizer.from_file(REF)
d = json.load(open(REF, encoding="utf-8"))
vocab = d["model"]["vocab"]
special = [t["id"] for t in d["added_tokens"] if t["special"]]
rng = random.Random(0)
cases = []
seqs = [[1, 2], [1], [], [1, 244, 2], [1, 5, 6, 249, 2, 4, 4], [244, 247], [3, 0, 4]]
for n in (1, 2, 5, 17, 64, 129):
    for _ in range(6):
        seqs.append([rng.randrange(0, len(vocab)) for _ in range(n)])
for _ in range(10):                      # realistic rows: BOS, subwords, EOS, padding
    body = [rng.randrange(5, len(vocab)) for _ in range(rng.randrange(1, 40))]
    seqs.append([1] + body + [2] + [4] * rng.randrange(0, 8))
for ids in seqs:
    cases.append({"ids": ids, "text": tok.decode(ids, skip_special_tokens=True),
                  "text_with_special": tok.decode(ids, skip_special_tokens=False)})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "text_detok.json")
json.dump({"vocab": vocab, "special_ids": special, "suffix": d["decoder"]["suffix"], "cases": cases},
          open(out, "w", encoding="utf-8"), ensure_ascii=False)
print("wrote", out, len(cases), "cases")
