"""sq_recovery_b200 -- B200-native drop-in for the loss hot path of timoblak/sq-recovery.

    from sq_recovery_b200.classes import ExplicitLoss, ImplicitLoss, IoUAccuracy, LeastSquares

replaces ``from classes import ...`` in the reference's ``torch/train.py:8``, ``torch/visu.py:8``,
``torch/test.py:7`` and ``torch/test_random.py:7``.  The classes keep the reference constructors, call
signatures and autograd behaviour (``torch/classes.py:109-447``) and run hand-written CUDA kernels for
sm_100a through the C ABI in ``include/sqloss.h`` (``libsqloss.so``, built by ``python -m sq_recovery_b200.build``).
There is no CPU fallback: constructing a loss on a CPU device or calling it without the library raises.
"""
from .classes import ExplicitLoss, ImplicitLoss, IoUAccuracy, LeastSquares  # noqa: F401
from . import quaternion  # noqa: F401

__all__ = ["ExplicitLoss", "ImplicitLoss", "IoUAccuracy", "LeastSquares", "quaternion"]
