"""B200-native batched geochemical reaction path for PFLOTRAN."""
__version__ = "0.1.0"
