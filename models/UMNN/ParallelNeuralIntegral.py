"""Shim: same names as the reference module models/UMNN/ParallelNeuralIntegral.py, served by umnn_b200."""
from umnn_b200.integral import ParallelNeuralIntegral, integrate, computeIntegrand, _flatten  # noqa: F401
from umnn_b200.quadrature import compute_cc_weights, _host_cache as _cc_weights_cache  # noqa: F401
