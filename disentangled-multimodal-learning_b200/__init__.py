"""B200-native (sm_100a) implementation of the WSI multimodal-MIL attention hot path.

Drop-in mirrors of the reference operator interface (helenypzhang/Disentangled-Multimodal-Learning):
``NystromAttention``, ``DeformCrossAttention1D``, ``DeformCrossTransMIL``, ``TransLayer``,
``TransMIL``, ``DeformPathomicNet``, ``GatherLayer``.  Every compute call goes through the C-ABI
library built from ``csrc/`` (see ``include/dml_b200.h``); there is no CPU or eager fallback.
Import as ``dml_b200`` (see ``dml_b200/__init__.py``).
"""
__version__ = "0.1.0"
