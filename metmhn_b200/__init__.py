"""metmhn_b200: B200-native (sm_100a CUDA) implementation of metMHN's training hot path --
the per-patient marginal log-likelihood and its exact gradient -- behind the call surface of
the reference's `metmhn.regularized_optimization`."""
import os as _os

# An evaluation overlaps independent chunks on up to 32 side streams; the default of 8 hardware work queues
# would serialise them.  Only effective if set before the CUDA context of this process is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import regularized_optimization  # noqa: F401,E402
from . import likelihood  # noqa: F401,E402
from ._lib import Handle, MetMHNError, measure_fp64_tflops  # noqa: F401,E402
from .regularized_optimization import (  # noqa: F401,E402
    learn_mhn, score, score_and_grad, score_and_grad_reg, score_reg, symmetric_penal)
