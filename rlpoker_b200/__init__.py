"""B200-native CFR solver behind the rlpoker Python API."""
