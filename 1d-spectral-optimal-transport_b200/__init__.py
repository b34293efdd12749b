"""B200-native Spectral Optimal Transport loss (drop-in for the reference's losses.Wasserstein1D)."""
__version__ = "0.1.0"
