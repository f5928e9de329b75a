"""pq3d_b200 — B200-native promptable query decoder (PQ3D hot path)."""
__version__ = "0.1.0"
