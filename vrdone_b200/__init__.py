"""vrdone_b200: B200-native (sm_100a) implementation of the VrdONE ``MaskVRD`` inference hot path.

``MaskVRD`` is a drop-in for the reference ``models.maskvrd.MaskVRD`` in eval mode; kernels live in
``csrc/`` behind the C ABI declared in ``include/vrdone_b200.h`` (``libvrdone_b200.so``, built in-tree
by ``vrdone_b200.build``).
"""
from .maskvrd import MaskVRD  # noqa: F401
from .synth import load_config  # noqa: F401

__all__ = ["MaskVRD", "load_config"]
