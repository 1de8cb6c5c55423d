"""Top-level `utils` package mirroring the reference's import surface (`from utils.criterion import GANLoss`)."""
