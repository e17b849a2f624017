"""umnn_b200: B200-native Clenshaw-Curtis integration hot path of UMNN behind the reference's
nn.Module / autograd.Function surface.  See DESIGN.md; the C ABI is in include/umnn_b200.h."""
from .flow import EmbeddingNetwork, MonotonicNN, UMNNMAF, UMNNMAFFlow  # noqa: F401
from .graphs import GraphedLogLikelihood  # noqa: F401
from .integral import (NeuralIntegral, ParallelNeuralIntegral, UnsupportedIntegrandError, cc_integrate,  # noqa: F401
                       cc_integrate_host, integrate, integrate_sequential, prepare_integral, torch_route)
from .kernel import invalidate_packed  # noqa: F401
from .networks import (ConditionnalMADE, ContiguousIntegrand, ELUPlus, IntegrandNN, IntegrandNetwork,  # noqa: F401
                       MADE, MaskedLinear)
from .quadrature import compute_cc_weights  # noqa: F401

__version__ = "0.1.0"
