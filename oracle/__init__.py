"""ORACLE — test infrastructure only (CPU restatement of the reference's algorithms).

Nothing under ``oracle/`` is on the product path. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / CPU baseline.

Pinning: every function here is checked against outputs of the *live* reference
(imported from /root/reference in the build container by tests/golden/make_golden.py);
the resulting hashes/tensors are committed under tests/golden/ and re-checked by
tests/test_oracle_golden.py. The reference itself ships no tests or golden vectors
(SURVEY.md §4), so those generated fixtures are the pin.
"""
