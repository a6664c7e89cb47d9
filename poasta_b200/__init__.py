"""poasta_b200 — B200-native gap-affine sequence-to-graph aligner behind POASTA's aligner API."""
from .errors import (CudaError, GraphError, InvalidAlignment, NoAlignment, PoastaError, Unsupported,
                     WeightsUnequalSize)
from .graph import AlignedPair, POAGraph
from .aligner import (AffineDijkstra, AffineMinGapCost, AffinePathAware, AlignmentType, AstarResult, GapAffine,
                      PoastaAligner, TIEBREAK_DP, TIEBREAK_EXACT, align_batch, msa_batch)

__all__ = ["POAGraph", "AlignedPair", "PoastaAligner", "GapAffine", "AlignmentType", "AffineDijkstra", "AffineMinGapCost",
           "AffinePathAware", "AstarResult", "align_batch", "msa_batch", "TIEBREAK_EXACT", "TIEBREAK_DP", "PoastaError", "WeightsUnequalSize", "InvalidAlignment", "GraphError",
           "CudaError", "NoAlignment", "Unsupported"]
