"""CPU oracle for the BioReason hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import anything from this package, and only as the checker / timed CPU baseline.
The product (`bioreason_b200/`) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the UNMODIFIED reference code itself, executed in the build container by
`tests/golden/make_golden.py` (fixtures committed under `tests/golden/`); see
`tests/test_oracle_golden.py`.  The NT-v2 encoder deltas (gated-SiLU FFN, no FFN bias) are restated
from the public model card because the hub's remote code is unreachable offline -- that single item
is "parity unpinned" (DESIGN.md §3).
"""
