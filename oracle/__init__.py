"""CPU oracle for the Speech-Transformer forward / greedy-decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline, never as the thing shipped.  The product path
(``asr_transformer_b200``) never imports this package and fails loudly when the
CUDA library is missing.

Parity status: PINNED by execution.  The reference ships no tests or golden
vectors, so the oracle is pinned against the reference's
own PyTorch modules executed in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference`` and writes the
fixtures; ``tests/test_oracle.py`` checks the oracle against them, and, when
``/root/reference`` is importable, against the live reference as well).
"""
