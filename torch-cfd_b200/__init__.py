"""torch-cfd_b200: B200-native (sm_100a) implementation of torch-cfd's spectral hot path."""
__version__ = "0.1.0"
