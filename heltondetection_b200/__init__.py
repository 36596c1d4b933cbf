"""heltondetection_b200: B200-native post-CNN box pipeline."""
__version__ = "0.1.0"
