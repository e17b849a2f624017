"""Shim: same names as the reference module models/UMNN/UMNNMAFFlow.py, served by umnn_b200."""
from umnn_b200.flow import ListModule, UMNNMAFFlow, EmbeddingNetwork, UMNNMAF  # noqa: F401
