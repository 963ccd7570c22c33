"""ggnn_b200 -- B200 (sm_100a) implementation of GGNN's batched query hot path and graph-construction
kernels behind the reference's own API.  `import ggnn_b200 as ggnn` is the drop-in for `import ggnn`."""
from .api import (GGNN, QueryFuture, DistanceMeasure, Evaluation, Evaluator, FloatDataset, Graph, IntDataset, UCharDataset,
                  set_log_level)

__all__ = ["GGNN", "QueryFuture", "DistanceMeasure", "Evaluation", "Evaluator", "FloatDataset", "UCharDataset", "IntDataset",
           "Graph", "set_log_level"]
__version__ = "0.1.0"
