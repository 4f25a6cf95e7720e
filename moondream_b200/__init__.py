"""moondream_b200 — B200-native (sm_100a) engine for moondream's batched VLM forward path."""

__version__ = "0.1.0"
