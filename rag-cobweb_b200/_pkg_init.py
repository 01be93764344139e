"""cobweb-b200: B200-native engine for the Cobweb concept-tree hot path."""
from . import synth  # noqa: F401

__all__ = ["synth"]
