"""CPU oracle for the pyellipticfd hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement (NumPy, and plain C under
``oracle/c``) of the reference algorithms at /root/reference, each function
citing the reference file:line it follows.  It exists to *check* the CUDA path.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package ``pyellipticfd_b200`` never imports it and has no CPU fallback.

Pinning: every restatement here is checked against the *unmodified* reference
(imported through ``oracle/refshim.py`` in the build container) by
``oracle/make_golden.py``, which also writes the committed fixtures in
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the restatements
against those fixtures on any machine (no reference tree needed).
"""
