"""The replay recipe of fastpath.py on the CPU: record an evaluation (plans
interpreted by tests/wbx_emulator.py), pack the raw launch results the way a
replay lays them out in its device buffer, and check that ``decode`` returns
exactly what the ordinary path returned -- for plain, binned, outer-binned,
climatology, masked, categorical and ensemble launches.  The launch side
(device buffers, streams, pinned slots) is covered by the `-m gpu` tests."""

import numpy as np
import pytest

import wbx_emulator
from weatherbenchx_b200 import _cabi, aggregation, binning, fastpath, weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import categorical, deterministic
from weatherbenchx_b200.metrics import probabilistic, wrappers

NY, NX = 8, 16
LAT = np.linspace(-87.5, 87.5, NY)
LON = np.linspace(0, 360, NX, endpoint=False)
INIT = (np.datetime64('2020-01-01T00', 'ns') +
        np.arange(4) * np.timedelta64(12, 'h'))
LEAD = (np.arange(3) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
DIMS = ('init_time', 'lead_time', 'latitude', 'longitude')
COORDS = {'init_time': INIT, 'lead_time': LEAD, 'latitude': LAT,
          'longitude': LON}


def _data(seed, names=('a', 'b'), nan=False):
  rng = np.random.default_rng(seed)
  P, T = {}, {}
  for n in names:
    p = rng.normal(280, 5, (4, 3, NY, NX)).astype(np.float32)
    t = (p + rng.normal(0, 2, p.shape)).astype(np.float32)
    if nan:
      t[1, 2, 3, 4] = np.nan
    P[n] = xl.DataArray(p, DIMS, coords=COORDS, name=n)
    T[n] = xl.DataArray(t, DIMS, coords=COORDS, name=n)
  return P, T, rng


def _replay_equals_ordinary(metrics, agg, P, T):
  with fastpath.recording() as rec:
    stats = metrics_base.compute_unique_statistics_for_all_metrics(
        metrics, P, T)
    state = agg.aggregate_statistics(stats)
  want = state.metric_values(metrics)
  assert rec.clean and rec.groups
  chunk = fastpath.CompiledChunk(rec, metrics, [], list(want.keys()),
                                 allocate=False)
  flat = np.full(chunk._n, np.nan)
  for kind, glaunches, _ in chunk.groups:
    for launch, off_ws, off_w, ws_cols, w_cols in glaunches:
      ws, w = launch.plan.run_to_host()
      assert ws.shape == (launch.n_rows, ws_cols)
      flat[off_ws:off_ws + ws.size] = ws.ravel()
      flat[off_w:off_w + w.size] = w.ravel()
  got = chunk.decode(flat)
  assert list(got) == list(want)
  for k in want:
    assert got[k].dims == want[k].dims, k
    np.testing.assert_array_equal(got[k].values, want[k].values, err_msg=k)
    for d in want[k].dims:
      if d in want[k].coords:
        np.testing.assert_array_equal(got[k].coords[d].values,
                                      want[k].coords[d].values)
  return got


def test_plain_and_climatology(monkeypatch):
  wbx_emulator.installed(monkeypatch)
  P, T, rng = _data(0)
  clim = {n: xl.DataArray(
      rng.normal(280, 3, (366, 4, NY, NX)).astype(np.float32),
      ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords={'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
              'latitude': LAT, 'longitude': LON}, name=n) for n in P}
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE(),
             'bias': deterministic.Bias(), 'acc': deterministic.ACC(clim)}
  for rd in (['init_time', 'latitude', 'longitude'],
             ['latitude', 'longitude'],
             ['init_time', 'lead_time', 'latitude', 'longitude']):
    agg = aggregation.Aggregator(reduce_dims=rd,
                                 weigh_by=[weighting.GridAreaWeighting()])
    got = _replay_equals_ordinary(metrics, agg, P, T)
    assert set(got) == {f'{m}.{v}' for m in metrics for v in P}


def test_masked_and_skipna(monkeypatch):
  wbx_emulator.installed(monkeypatch)
  P, T, _ = _data(1, nan=True)
  from weatherbenchx_b200.data_loaders import base as loaders_base
  Tm = loaders_base.add_nan_mask_to_data(T)
  metrics = {'rmse': deterministic.RMSE(), 'bias': deterministic.Bias()}
  rd = ['init_time', 'latitude', 'longitude']
  _replay_equals_ordinary(
      metrics, aggregation.Aggregator(reduce_dims=rd, masked=True), P, Tm)
  _replay_equals_ordinary(
      metrics, aggregation.Aggregator(reduce_dims=rd, skipna=True), P, T)
  got = _replay_equals_ordinary(
      metrics, aggregation.Aggregator(reduce_dims=rd), P, T)
  assert np.isnan(got['rmse.a'].values).any()


def test_bins(monkeypatch):
  wbx_emulator.installed(monkeypatch)
  P, T, rng = _data(2)
  land = xl.DataArray(rng.random((NY, NX)) < 0.4, ('latitude', 'longitude'),
                      coords={'latitude': LAT, 'longitude': LON})
  regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
             'box': ((10, 80), (30, 200))}
  metrics = {'rmse': deterministic.RMSE(), 'mse': deterministic.MSE()}
  rd = ['init_time', 'latitude', 'longitude']
  for bins in ([binning.Regions(regions, land_sea_mask=land)],
               [binning.ByTimeUnit('hour', 'init_time')],
               [binning.ByTimeUnit('hour', 'init_time'),
                binning.Regions(regions)],
               [binning.LatitudeBins(45), binning.LandSea(
                   land.astype(np.float32))]):
    agg = aggregation.Aggregator(
        reduce_dims=rd, weigh_by=[weighting.GridAreaWeighting()], bin_by=bins)
    _replay_equals_ordinary(metrics, agg, P, T)


def test_categorical_and_ensemble(monkeypatch):
  wbx_emulator.installed(monkeypatch)
  P, T, rng = _data(3, names=('a',))
  both = [wrappers.ContinuousToBinary('both', [278.0, 282.0], 'threshold')]
  metrics = {'csi': wrappers.WrappedMetric(categorical.CSI(), both),
             'rmse': deterministic.RMSE()}
  rd = ['init_time', 'latitude', 'longitude']
  agg = aggregation.Aggregator(reduce_dims=rd,
                               weigh_by=[weighting.GridAreaWeighting()])
  _replay_equals_ordinary(metrics, agg, P, T)
  x = rng.normal(size=(4, 6, NY, NX)).astype(np.float32)
  y = rng.normal(size=(4, NY, NX)).astype(np.float32)
  ecoords = {'init_time': INIT, 'number': np.arange(6), 'latitude': LAT,
             'longitude': LON}
  X = {'e': xl.DataArray(x, ('init_time', 'number', 'latitude', 'longitude'),
                         coords=ecoords, name='e')}
  Y = {'e': xl.DataArray(y, ('init_time', 'latitude', 'longitude'),
                         coords={k: ecoords[k] for k in
                                 ('init_time', 'latitude', 'longitude')},
                         name='e')}
  ens = {'crps': probabilistic.CRPSEnsemble(ensemble_dim='number'),
         'ssr': probabilistic.UnbiasedSpreadSkillRatio(ensemble_dim='number')}
  _replay_equals_ordinary(ens, agg, X, Y)


def test_unrecordable_routes_are_not_compiled(monkeypatch):
  """Wind-vector sums, member means and the generic kernel keep taking the
  ordinary path: the recorder is marked dirty."""
  wbx_emulator.installed(monkeypatch)
  P, T, _ = _data(4, names=('u', 'v'))
  metrics = {'wv': deterministic.WindVectorRMSE(['u'], ['v'], ['wind'])}
  agg = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'])
  with fastpath.recording() as rec:
    stats = metrics_base.compute_unique_statistics_for_all_metrics(
        metrics, P, T)
    agg.aggregate_statistics(stats)
  assert not rec.clean
  # host inputs never get a key: nothing to replay
  assert fastpath.chunk_key(metrics, agg, P, T) is None
