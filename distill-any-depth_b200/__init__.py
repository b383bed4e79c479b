"""B200-native (sm_100a) drop-in for the Distill-Any-Depth hot path."""
from . import synthetic  # noqa: F401
