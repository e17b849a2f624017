"""Shim: same names as the reference module models/UMNN/made.py, served by umnn_b200."""
from umnn_b200.networks import MaskedLinear, MADE, ConditionnalMADE  # noqa: F401
