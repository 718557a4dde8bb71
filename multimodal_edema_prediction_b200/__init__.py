"""B200-native DuETT hot path (drop-in for lastdancewithyou/multimodal_edema_prediction's DuETT path)."""
__version__ = "0.1.0"
