"""easyfea_b200 — B200-native (sm_100a, FP64) element integration + CSR assembly behind EasyFEA's Python API."""
__version__ = "0.1.0"
