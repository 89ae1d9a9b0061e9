"""B200-native reverse-diffusion sampler hot path for DiffBindFR."""
__version__ = "0.1.0"
