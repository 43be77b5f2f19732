"""B200-native Gaussian-distance loss."""
