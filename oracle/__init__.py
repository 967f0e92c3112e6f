"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the fake-quant / observer / QLinear hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and there only
as the checker (or as the timed CPU baseline), never as the shipped path.  The product
package ``outlier_suppression_b200`` never imports this package and fails loudly when the
CUDA extension is missing.

Parity status: PINNED against the reference itself.  The reference repo ships no tests or
golden vectors, so ``tests/golden/gen_golden.py`` imports the unmodified
reference from ``/root/reference`` (via ``oracle/ref_shim.py``) in the build container, runs
it on seeded inputs and commits the outputs under ``tests/golden/*.npz``; ``tests/test_oracle_*``
check this restatement against those vectors bit-for-bit.
"""
