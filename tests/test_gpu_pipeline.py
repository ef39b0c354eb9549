"""GPU tests of the chunk driver with the real Aggregator (fused kernels):
chunked evaluation == one-chunk evaluation == oracle, files included."""

import numpy as np
import pytest

import wbx_oracle as oracle
from test_pipeline import INIT, LAT, LEAD, _datasets
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import io_netcdf
from weatherbenchx_b200 import pipeline
from weatherbenchx_b200 import time_chunks
from weatherbenchx_b200 import weighting
from weatherbenchx_b200.data_loaders import array_loaders
from weatherbenchx_b200.metrics import deterministic

pytestmark = pytest.mark.gpu
RTOL = 1e-5

METRICS = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE(),
           'wind': deterministic.WindVectorRMSE('t', 'z', 'tz')}


def _oracle_values(reduce_dims):
  preds, tgts = _datasets()
  p = array_loaders.PredictionsFromArrays(preds).load_chunk(INIT, LEAD)
  t = array_loaders.TargetsFromArrays(tgts).load_chunk(INIT, LEAD)
  dims = p['t'].dims
  w = [(oracle.grid_area_weights(LAT), ('latitude',))]
  out = {}
  se = {}
  for var in ('t', 'z'):
    pv, tv = p[var].values, t[var].values
    se[var] = oracle.squared_error(pv, tv)
    ws, sw, _ = oracle.aggregate(se[var], dims, reduce_dims, weights=w)
    out[f'rmse.{var}'] = np.sqrt(ws / sw)
    ws, sw, _ = oracle.aggregate(oracle.absolute_error(pv, tv), dims,
                                 reduce_dims, weights=w)
    out[f'mae.{var}'] = ws / sw
  ws, sw, _ = oracle.aggregate(se['t'] + se['z'], dims, reduce_dims, weights=w)
  out['wind.tz'] = np.sqrt(ws / sw)
  return out


@pytest.mark.parametrize('reduce_dims', [
    ['init_time', 'latitude', 'longitude'], ['latitude', 'longitude']])
@pytest.mark.parametrize('chunks', [(1, 2), (4, None), (None, None)])
def test_pipeline_on_gpu_matches_oracle(reduce_dims, chunks, tmp_path):
  preds, tgts = _datasets()
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=chunks[0],
                                 lead_time_chunk_size=chunks[1])
  agg = aggregation.Aggregator(reduce_dims=reduce_dims,
                               weigh_by=[weighting.GridAreaWeighting()])
  path = str(tmp_path / 'metrics.nc')
  out = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds),
      array_loaders.TargetsFromArrays(tgts, add_nan_mask=True), METRICS, agg,
      out_path=path)
  values = out[None][1]
  expected = _oracle_values(reduce_dims)
  assert set(values) == set(expected)
  written = io_netcdf.open_dataset(path)
  for k, e in expected.items():
    np.testing.assert_allclose(values[k].values, e, rtol=RTOL)
    np.testing.assert_array_equal(written[k].values, values[k].values)
  if 'init_time' not in reduce_dims:
    np.testing.assert_array_equal(values['rmse.t'].coords['init_time'].values,
                                  INIT)
