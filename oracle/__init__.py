"""CPU oracle for the SQ implicit-function losses -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch fp64 on the host) of the hot path of
timoblak/sq-recovery: ``torch/classes.py:109-447`` and
``torch/quaternion.py:19-21,46-67``.  It exists to *check* the CUDA product in
``sq_recovery_b200`` and to serve as the timed CPU baseline of ``bench.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``sq_recovery_b200/`` does;
the product path raises when its CUDA library is missing rather than falling
back to anything in here.

Parity status: PINNED.  ``oracle/make_goldens.py`` imports the unmodified
reference from ``/root/reference/torch`` (dev container only) and freezes its
outputs on the reference's own fixtures (``data/example_imgs`` + ``labels.txt``,
``classes.py:458-461``, ``visu.py:77``) and on seeded random inputs into
``tests/golden/``; ``tests/test_oracle_goldens.py`` checks this restatement
against those vectors on every CPU test run.
"""
