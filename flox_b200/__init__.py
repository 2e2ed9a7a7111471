"""flox_b200 — B200-native grouped reductions behind flox's `groupby_reduce` / `Aggregation` /
engine plug-in contract.  See DESIGN.md; the C ABI is include/flox_b200.h."""
from . import aggregate_b200, parallel  # noqa: F401
from .aggregations import Aggregation, generic_aggregate  # noqa: F401
from .chunked import ChunkedArray  # noqa: F401
from .core import chunk_reduce, groupby_reduce  # noqa: F401
from .reindex import ReindexArrayType, ReindexStrategy, reindex_  # noqa: F401
from .scan import groupby_scan  # noqa: F401

__version__ = "0.1.0"
__all__ = ["groupby_reduce", "groupby_scan", "chunk_reduce", "generic_aggregate", "Aggregation", "ChunkedArray", "ReindexStrategy",
           "ReindexArrayType", "reindex_", "aggregate_b200"]
