"""weatherbenchx_b200 -- B200-native engine behind the WeatherBench-X
Metric / Statistic / Aggregator surface.

Only the statistic + aggregation hot path is provided:
  metrics.base / metrics.deterministic / metrics.probabilistic / metrics.spectral
  aggregation (Aggregator, AggregationState), weighting (GridAreaWeighting),
  binning (Regions), distributed (shard + all-reduce of AggregationStates).
Field arithmetic happens in hand-written sm_100a CUDA kernels behind the C ABI
in include/wbx_b200.h; there is no CPU fallback.
"""

from weatherbenchx_b200.xarray_lite import DataArray, Dataset  # noqa: F401

__version__ = '0.1.0'
