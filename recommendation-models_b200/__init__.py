"""b200rec: the B200 (sm_100a) hot path of yaochitc/recommendation-models behind the reference's
own module / model surface.  Everything computes in libb200rec.so (hand-written CUDA, C ABI in
include/b200rec.h); this package is the thin host mirror of the reference interface.

The directory name carries a hyphen (it is the project name), so import it through
``__graft_entry__.load_package()`` which registers it as ``recommendation_models_b200``.
"""
from . import _lib
from ._lib import B200RecError, device_count, launch_count, lib
from .models import (InternalDCNModel, InternalDeepFMModel, InternalFMModel, InternalLRModel,
                     InternalPNNModel, InternalXDeepFMModel, make_model)
from .nn import (CINEncoder, CrossEncoder, DotProduct2, DuplicateTable, FirstOrderEncoder, Gather, HigherOrderEncoder,
                 Linear, ProductEncoder, Scatter, SecondOrderEncoder)
from .ps import EmbeddingTable, ParRecModel, distinct, scatter_add
from . import data, metrics, sharded, synth
