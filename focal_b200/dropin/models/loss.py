"""Shadow of the reference's ``models/loss.py``.

``/root/reference/src/models`` has no ``__init__.py`` (namespace package), so putting THIS directory's parent
(``focal_b200/dropin``) before ``<reference>/src`` on ``PYTHONPATH`` makes ``from models.loss import FOCALLoss``
(reference ``src/train_utils/model_selection.py:11``) resolve here while every other ``models.*`` module still
comes from the reference.  ``train.py -learn_framework=FOCAL`` then runs unmodified on the B200 kernels.
"""
from focal_b200.loss import FOCALLoss  # noqa: F401
