"""Shim: same names as the reference module models/UMNN/UMNNMAF.py, served by umnn_b200."""
from umnn_b200.networks import (ELUPlus, dict_act_func, _flatten, compute_lipschitz_linear,  # noqa: F401
                                IntegrandNetwork, MADE, ConditionnalMADE)
from umnn_b200.flow import UMNNMAF, EmbeddingNetwork  # noqa: F401
from umnn_b200.integral import NeuralIntegral, ParallelNeuralIntegral  # noqa: F401
from umnn_b200.integral import integrate_sequential as sequential_integrate  # noqa: F401
from umnn_b200.integral import integrate as parallel_integrate  # noqa: F401
