"""CPU oracle for the neosr training-step hot path.

TEST INFRASTRUCTURE ONLY. This package is a plain PyTorch-CPU (fp32/fp64)
restatement of the reference algorithm for `image.feed_data` +
`image.optimize_parameters` and the modules underneath it. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it; the product path (`neosr_b200/`) never does.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against the *live reference modules* imported from
/root/reference in the build container (`oracle/ref_shim.py`,
`oracle/make_golden.py`) and against the fixtures those scripts wrote into
`tests/golden/`.  `tests/test_oracle_golden.py` re-checks the oracle against
the committed fixtures on every run; `tests/test_oracle_vs_reference.py`
re-checks it against the live reference whenever /root/reference exists.
"""
