"""metmhn_b200: B200-native (sm_100a CUDA) implementation of metMHN's training hot path --
the per-patient marginal log-likelihood and its exact gradient -- behind the call surface of
the reference's `metmhn.regularized_optimization`."""
from . import regularized_optimization  # noqa: F401
from . import likelihood  # noqa: F401
from ._lib import Handle, MetMHNError, measure_fp64_tflops  # noqa: F401
from .regularized_optimization import (  # noqa: F401
    learn_mhn, score, score_and_grad, score_and_grad_reg, score_reg, symmetric_penal)
