"""focal_b200 -- B200-native (sm_100a) implementation of the FOCAL contrastive-loss hot path.

Public API: :class:`focal_b200.loss.FOCALLoss` (mirror of the reference module) and the lower-level
:class:`focal_b200.engine.FocalEngine`.  The numerical work lives in ``libfocal_b200.so`` (C ABI in
``include/focal_b200.h``); build it with ``python -m focal_b200.build``.
"""
from .engine import FocalEngine, FocalHyper  # noqa: F401
from .loss import FOCALLoss  # noqa: F401

__all__ = ["FOCALLoss", "FocalEngine", "FocalHyper"]
