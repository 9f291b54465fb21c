"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's MC-dropout GA-MIL head
(`/root/reference/model.py:256-328`, `MultiHeadGatedAttentionMIL.mc_inference`).

Nothing in the shipped package (`montecarlo-gated-mil_b200/`) imports this
directory.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may use it, and only as the checker or
the timed CPU baseline — never as the product path.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so the oracle is pinned against the reference *executed live*
in the build container (`tests/golden/make_golden.py` imports
`/root/reference/model.py`, injects dropout masks and stores the reference's
outputs under `tests/golden/`).  `tests/test_oracle.py` checks the oracle
against those stored reference outputs.
"""
