"""transcar_b200 - B200-native (sm_100a) TransCAR fusion-decoder hot path."""
