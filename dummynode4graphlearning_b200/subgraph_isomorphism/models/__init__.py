from .basemodel import GraphAdjModel, GraphAdjModelV2  # noqa: F401
from .compgcn import CompGCN, CompGCNLayer  # noqa: F401
from .dmpnn import DMPNN, DMPLayer  # noqa: F401
from .embed import (EquivariantEmbedding, MultihotEmbedding, NormalEmbedding, OrthogonalEmbedding,  # noqa: F401
                    PositionEmbedding, UniformEmbedding)
from .filter import ScalarFilter  # noqa: F401
from .pred import MaxPredictNet, MeanPredictNet, PredictNet, SumPredictNet  # noqa: F401
from .rgin import RGIN, RGINLayer  # noqa: F401
from .rgcn import RGCN, RGCNLayer  # noqa: F401
