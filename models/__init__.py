"""Import-path shim: `from models import UMNNMAFFlow` etc. resolve to umnn_b200 (reference: models/__init__.py:1)."""
from models.UMNN import UMNNMAFFlow, MADE, ParallelNeuralIntegral, NeuralIntegral  # noqa: F401
