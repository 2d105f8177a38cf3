"""ipavsr_b200 — B200-native (sm_100a) implementation of ip-avsr's AdeNet/DeltaNet hot path."""
__version__ = '0.1.0'
