"""Import-path shim for the reference's `models.UMNN` package (models/UMNN/__init__.py:1-6).

As in the reference, `models.UMNN.NeuralIntegral` / `models.UMNN.ParallelNeuralIntegral` resolve to the
autograd Function classes while `from models.UMNN.ParallelNeuralIntegral import integrate` still reaches
the submodule.
"""
from .UMNNMAFFlow import UMNNMAFFlow  # noqa: F401
from .MonotonicNN import MonotonicNN, IntegrandNN  # noqa: F401
from .UMNNMAF import IntegrandNetwork, UMNNMAF  # noqa: F401
from .made import MADE  # noqa: F401
from .NeuralIntegral import NeuralIntegral  # noqa: F401
from .ParallelNeuralIntegral import ParallelNeuralIntegral  # noqa: F401
