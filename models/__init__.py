"""Top-level `models` package: the import surface the reference's training scripts use
(`from models import dcgan, dcgan_specnorm, ...`, main_dcgan.py:11, main_sngan.py:11), re-exporting the B200-native
mirrors so those scripts run unchanged from this repository root."""
