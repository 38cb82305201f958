"""B200-native BP5 matrix-free CG hot path.  Import as `dealceed_b200`."""
