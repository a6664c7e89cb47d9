"""poasta_b200 — B200-native gap-affine sequence-to-graph aligner behind POASTA's aligner API."""
from .errors import (CudaError, GraphError, InvalidAlignment, NoAlignment, PoastaError, Unsupported,
                     WeightsUnequalSize)
from .graph import AlignedPair, POAGraph

__all__ = ["POAGraph", "AlignedPair", "PoastaError", "WeightsUnequalSize", "InvalidAlignment", "GraphError",
           "CudaError", "NoAlignment", "Unsupported"]
