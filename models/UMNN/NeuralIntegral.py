"""Shim: same names as the reference module models/UMNN/NeuralIntegral.py, served by umnn_b200."""
from umnn_b200.integral import NeuralIntegral, _flatten  # noqa: F401
from umnn_b200.integral import integrate_sequential as integrate  # noqa: F401
from umnn_b200.integral import computeIntegrand_sequential as computeIntegrand  # noqa: F401
from umnn_b200.quadrature import compute_cc_weights, _host_cache as _cc_weights_cache  # noqa: F401
