"""ResNet50 / ResNet50-IBN-A trunks: reference-layout parameters + the B200 inference engine."""
