"""erd_b200: B200-native (sm_100a) GFL + Elastic Response Distillation loss path.

Public surface mirrors the reference's plugin API for this path:
``GFLHeadIncrementERD.loss_by_feat`` / ``loss`` and ``GFLIncrementERD.sel_pos`` / ``loss``.
Importing the package does not need a GPU; calling into it does (there is no CPU path).
"""
from .synth import Batch, make_batch  # noqa: F401

__all__ = ['GFLHeadIncrementERD', 'GFLIncrementERD', 'ErdPath', 'make_batch', 'Batch']


def __getattr__(name):
    if name in ('GFLHeadIncrementERD', 'parse_losses', 'ErsSelection'):
        from . import head
        return getattr(head, name)
    if name == 'GFLIncrementERD':
        from . import detector
        return detector.GFLIncrementERD
    if name in ('ErdPath', 'default_path'):
        from . import ops
        return getattr(ops, name)
    raise AttributeError(name)
