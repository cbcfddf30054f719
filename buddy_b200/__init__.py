"""buddy_b200 — B200-native (sm_100a) hot path for BUDDy reverse-diffusion dereverberation."""
__version__ = "0.1.0"
