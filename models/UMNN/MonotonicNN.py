"""Shim: same names as the reference module models/UMNN/MonotonicNN.py, served by umnn_b200."""
from umnn_b200.networks import IntegrandNN, _flatten  # noqa: F401
from umnn_b200.flow import MonotonicNN  # noqa: F401
from umnn_b200.integral import NeuralIntegral, ParallelNeuralIntegral  # noqa: F401
