"""Import-path shim for the reference's `models.UMNN` package (reference: models/UMNN/__init__.py:1-6).

As there, `models.UMNN.NeuralIntegral` / `models.UMNN.ParallelNeuralIntegral` name the autograd Function
classes (the class import shadows the submodule attribute), while
`from models.UMNN.ParallelNeuralIntegral import integrate` still reaches the submodule through sys.modules.
Everything is served by the `umnn_b200` package.
"""
from .made import MADE
from .MonotonicNN import IntegrandNN, MonotonicNN
from .UMNNMAF import UMNNMAF, IntegrandNetwork
from .UMNNMAFFlow import UMNNMAFFlow
from .NeuralIntegral import NeuralIntegral          # the class, not the module
from .ParallelNeuralIntegral import ParallelNeuralIntegral

__all__ = ["MADE", "IntegrandNN", "MonotonicNN", "UMNNMAF", "IntegrandNetwork", "UMNNMAFFlow", "NeuralIntegral",
           "ParallelNeuralIntegral"]
