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
@pytest.mark.parametrize('chunks,lanes', [((1, 2), 1), ((1, 2), 3),
                                          ((4, None), 2), ((None, None), 1)])
def test_pipeline_on_gpu_matches_oracle(reduce_dims, chunks, lanes, tmp_path):
  preds, tgts = _datasets()
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=chunks[0],
                                 lead_time_chunk_size=chunks[1])
  agg = aggregation.Aggregator(reduce_dims=reduce_dims,
                               weigh_by=[weighting.GridAreaWeighting()])
  path = str(tmp_path / 'metrics.nc')
  out = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds),
      array_loaders.TargetsFromArrays(tgts, add_nan_mask=True), METRICS, agg,
      out_path=path, lanes=lanes)
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


def test_per_init_time_state_sums_to_the_reduced_one():
  """Keeping init_time (the state statistical_inference consumes) and summing
  it afterwards (AggregationState.sum_along_dims, aggregation.py:150-175)
  equals reducing init_time on the GPU; host fields against a climatology
  that stays on the device."""
  import torch
  from weatherbenchx_b200 import engine
  from weatherbenchx_b200 import xarray_lite as xl
  preds, tgts = _datasets(('t',))
  rng = np.random.default_rng(1)
  grid = {k: preds['t'].coords[k].values for k in ('latitude', 'longitude')}
  clim = {'t': engine.to_device(xl.DataArray(
      rng.normal(size=(366, 4, len(LAT), 12)).astype(np.float32),
      ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords=dict(grid, dayofyear=np.arange(1, 367),
                  hour=np.arange(0, 24, 6)), name='t'))}
  metrics = {'acc': deterministic.ACC(clim), 'rmse': deterministic.RMSE()}
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=2)
  out = {}
  for name, rd in (('kept', ['latitude', 'longitude']),
                   ('reduced', ['init_time', 'latitude', 'longitude'])):
    out[name] = pipeline.run_pipeline(
        times, array_loaders.PredictionsFromArrays(preds),
        array_loaders.TargetsFromArrays(tgts), metrics,
        aggregation.Aggregator(reduce_dims=rd,
                               weigh_by=[weighting.GridAreaWeighting()]),
        require_output=False)[None][0]
  assert torch.cuda.is_available()
  summed = out['kept'].sum_along_dims(['init_time']).metric_values(metrics)
  direct = out['reduced'].metric_values(metrics)
  assert out['kept'].sum_weights['SquaredError']['t'].dims == (
      'init_time', 'lead_time')
  for k in direct:
    np.testing.assert_allclose(summed[k].values, direct[k].values, rtol=1e-9)


def test_device_resident_archive_uses_strided_windows():
  """Forecasts and analyses already on the GPU: the (init, lead) target window
  is a torch.as_strided view of the analysis tensor, results equal the host
  run."""
  import torch
  from weatherbenchx_b200 import engine
  preds, tgts = _datasets(('t',))
  dpreds = {k: engine.to_device(v) for k, v in preds.items()}
  dtgts = {k: engine.to_device(v) for k, v in tgts.items()}
  loader = array_loaders.TargetsFromArrays(dtgts)
  chunk = loader.load_chunk(INIT[1:4], LEAD)['t']
  assert chunk.is_device and chunk.shape[:2] == (3, len(LEAD))
  assert (chunk.data.untyped_storage().data_ptr() ==
          dtgts['t'].data.untyped_storage().data_ptr())
  host = array_loaders.TargetsFromArrays(tgts).load_chunk(INIT[1:4], LEAD)['t']
  np.testing.assert_array_equal(chunk.values, host.values)
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE()}
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=3)
  agg = aggregation.Aggregator(reduce_dims=['init_time', 'latitude',
                                            'longitude'],
                               weigh_by=[weighting.GridAreaWeighting()])
  out_d = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(dpreds), loader, metrics, agg,
      require_output=False)[None][1]
  out_h = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds),
      array_loaders.TargetsFromArrays(tgts), metrics, agg,
      require_output=False)[None][1]
  assert torch.cuda.is_available()
  for k in out_h:
    np.testing.assert_allclose(out_d[k].values, out_h[k].values, rtol=1e-12)


@pytest.mark.parametrize('masked', [False, True])
@pytest.mark.parametrize('lanes', [1, 2])
def test_target_rows_cached_on_the_device(masked, lanes):
  """TargetsFromArrays(device_cache=True): every analysis row is uploaded once
  and stays on the GPU; host forecasts are streamed against device targets
  (WBX_FLAG_TARGET_DEVICE / MASK_DEVICE host-space plans).  Same numbers as
  the all-host run, with and without a NaN mask, ACC climatology on the
  device."""
  from weatherbenchx_b200 import engine
  from weatherbenchx_b200 import xarray_lite as xl
  preds, tgts = _datasets(('t', 'z'))
  if masked:
    tgts = {k: v.copy() for k, v in tgts.items()}
    tgts['t'].data[3, 2:5, 1:7] = np.nan
  rng = np.random.default_rng(4)
  grid = {k: preds['t'].coords[k].values for k in ('latitude', 'longitude')}
  clim = {v: engine.to_device(xl.DataArray(
      rng.normal(size=(366, 4, len(LAT), 12)).astype(np.float32),
      ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords=dict(grid, dayofyear=np.arange(1, 367),
                  hour=np.arange(0, 24, 6)), name=v)) for v in ('t', 'z')}
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE(),
             'acc': deterministic.ACC(clim)}
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=1)
  agg = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()], masked=masked)
  cached = array_loaders.TargetsFromArrays(tgts, device_cache=True,
                                           add_nan_mask=masked)
  chunk = cached.load_chunk(INIT[1:3], LEAD)['t']
  assert chunk.is_device
  host_chunk = array_loaders.TargetsFromArrays(tgts).load_chunk(
      INIT[1:3], LEAD)['t']
  np.testing.assert_array_equal(chunk.values, host_chunk.values)
  out_c = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds), cached, metrics, agg,
      require_output=False, lanes=lanes)[None][1]
  out_h = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds),
      array_loaders.TargetsFromArrays(tgts, add_nan_mask=masked), metrics, agg,
      require_output=False)[None][1]
  assert set(out_c) == set(out_h)
  for k in out_h:
    np.testing.assert_allclose(out_c[k].values, out_h[k].values, rtol=1e-12,
                               err_msg=k)
  # each row went up exactly once
  n_rows = len(np.unique((INIT[:, None] + LEAD[None, :]).ravel()))
  row_bytes = len(LAT) * 12 * 4
  assert cached.uploaded_bytes == 2 * n_rows * row_bytes
