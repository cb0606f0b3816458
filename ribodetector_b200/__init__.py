"""ribodetector_b200 — B200-native hot path of RiboDetector (encode → BiLSTM → FC → argmax /
pair-combine) behind the reference's plugin seam.  See DESIGN.md."""
__version__ = "0.1.0"
