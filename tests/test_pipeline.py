"""CPU tests of the on-box chunk driver: TimeChunks, the array loaders, the
NetCDF round trip and run_pipeline (single process and world_size-2 gloo).

The Aggregator is replaced by an oracle-backed stand-in with the same
``aggregate_statistics`` contract (the real one launches CUDA kernels; the GPU
version of these tests is tests/test_gpu_pipeline.py).  The identity under
test is the reference's beam_pipeline_test.py:82-170: chunked evaluation ==
evaluation of everything at once.
"""

import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.dirname(__file__)):
  if p not in sys.path:
    sys.path.insert(0, p)

import wbx_oracle as oracle  # noqa: E402
from weatherbenchx_b200 import aggregation  # noqa: E402
from weatherbenchx_b200 import io_netcdf  # noqa: E402
from weatherbenchx_b200 import pipeline  # noqa: E402
from weatherbenchx_b200 import time_chunks  # noqa: E402
from weatherbenchx_b200 import xarray_lite as xl  # noqa: E402
from weatherbenchx_b200.data_loaders import array_loaders  # noqa: E402
from weatherbenchx_b200.metrics import base as metrics_base  # noqa: E402
from weatherbenchx_b200.metrics import deterministic  # noqa: E402

H = np.timedelta64(1, 'h')


# ---------------------------------------------------------------------------
# TimeChunks (time_chunks.py:37-202; the docstring examples are the known
# answers)
# ---------------------------------------------------------------------------


def _example_times():
  init_times = np.arange('2020-01-01T00', '2020-01-02T00', 6 * H,
                         dtype='datetime64[h]')
  lead_times = np.arange(0, 18, 6, dtype='timedelta64[h]')
  return init_times, lead_times


def test_time_chunks_docstring_example_exact_lead_times():
  init_times, lead_times = _example_times()
  times = time_chunks.TimeChunks(init_times, lead_times,
                                 init_time_chunk_size=2, lead_time_chunk_size=2)
  got = list(times)
  assert len(times) == len(got) == 4
  expect = [(init_times[:2], lead_times[:2]), (init_times[:2], lead_times[2:]),
            (init_times[2:], lead_times[:2]), (init_times[2:], lead_times[2:])]
  for (gi, gl), (ei, el) in zip(got, expect):
    np.testing.assert_array_equal(gi, ei)
    np.testing.assert_array_equal(gl, el)
  assert gi.dtype == np.dtype('datetime64[ns]')
  assert gl.dtype == np.dtype('timedelta64[ns]')
  offsets = [o for o, _ in times.iter_with_chunk_offsets()]
  assert [(o.init_time, o.lead_time) for o in offsets] == [
      (0, 0), (0, 2), (2, 0), (2, 2)]
  np.testing.assert_array_equal(times[3][0], init_times[2:])
  with pytest.raises(IndexError):
    times[4]  # pylint: disable=pointless-statement


def test_time_chunks_lead_slice_and_errors():
  init_times, _ = _example_times()
  lead = slice(np.timedelta64(0, 'h'), np.timedelta64(6, 'h'))
  times = time_chunks.TimeChunks(init_times, lead, init_time_chunk_size=2)
  got = list(times)
  assert len(got) == 2 and got[0][1] == lead and got[1][1] == lead
  assert len(time_chunks.TimeChunks(init_times, lead)) == 1
  with pytest.raises(ValueError, match='not compatible for slice'):
    time_chunks.TimeChunks(init_times, lead, lead_time_chunk_size=1)
  with pytest.raises(ValueError, match='start and stop'):
    time_chunks.TimeChunks(init_times, slice(None, np.timedelta64(6, 'h')))
  with pytest.raises(ValueError, match='step must be None'):
    time_chunks.TimeChunks(init_times, slice(0 * H, 6 * H, 1 * H))
  with pytest.raises(ValueError, match='non-negative'):
    time_chunks.TimeChunks(init_times, lead, init_time_chunk_size=-1)
  with pytest.raises(ValueError, match='np.ndarray or slice'):
    time_chunks.TimeChunks(init_times, [0, 6])


# ---------------------------------------------------------------------------
# synthetic forecast / analysis pair + loaders
# ---------------------------------------------------------------------------

NLAT, NLON = 7, 12
LAT = np.linspace(-75, 75, NLAT)
LON = np.arange(NLON) * 30.0
INIT = np.arange('2020-01-01T00', '2020-01-04T00', 12 * H,
                 dtype='datetime64[ns]')            # 6 init times
LEAD = (np.arange(4) * 12 * H).astype('timedelta64[ns]')
VALID = np.arange('2020-01-01T00', '2020-01-06T00', 12 * H,
                  dtype='datetime64[ns]')


def _datasets(variables=('t', 'z')):
  rng = np.random.default_rng(0)
  preds, tgts = {}, {}
  for var in variables:
    truth = rng.normal(size=(len(VALID), NLAT, NLON)).astype(np.float32)
    tgts[var] = xl.DataArray(
        truth, ('valid_time', 'latitude', 'longitude'),
        coords={'valid_time': VALID, 'latitude': LAT, 'longitude': LON},
        name=var)
    fc = np.empty((len(INIT), len(LEAD), NLAT, NLON), np.float32)
    for i, it in enumerate(INIT):
      for j, lt in enumerate(LEAD):
        k = int(np.nonzero(VALID == it + lt)[0][0])
        fc[i, j] = truth[k] + (j + 1) * 0.1 * rng.normal(size=(NLAT, NLON))
    preds[var] = xl.DataArray(
        fc, ('init_time', 'lead_time', 'latitude', 'longitude'),
        coords={'init_time': INIT, 'lead_time': LEAD, 'latitude': LAT,
                'longitude': LON}, name=var)
  return preds, tgts


def test_loaders_select_like_the_reference():
  preds, tgts = _datasets(('t',))
  pl = array_loaders.PredictionsFromArrays(preds)
  tl = array_loaders.TargetsFromArrays(tgts, add_nan_mask=True)
  p = pl.load_chunk(INIT[2:4], LEAD[1:3])['t']
  t = tl.load_chunk(INIT[2:4], LEAD[1:3])['t']
  assert p.dims == t.dims == ('init_time', 'lead_time', 'latitude', 'longitude')
  assert p.shape == t.shape == (2, 2, NLAT, NLON)
  assert np.shares_memory(p.data, preds['t'].data)       # contiguous: a view
  np.testing.assert_array_equal(t.coords['valid_time'].values,
                                INIT[2:4, None] + LEAD[None, 1:3])
  k = int(np.nonzero(VALID == INIT[3] + LEAD[2])[0][0])
  np.testing.assert_array_equal(t.values[1, 1], tgts['t'].values[k])
  # regular init / lead spacing: the (init, lead) gather is a strided window
  # over the analysis, not a copy; an irregular one is gathered
  assert np.shares_memory(t.data, tgts['t'].data)
  t_irregular = tl.load_chunk(INIT[[0, 1, 3]], LEAD[1:3])['t']
  assert not np.shares_memory(t_irregular.data, tgts['t'].data)
  k = int(np.nonzero(VALID == INIT[3] + LEAD[1])[0][0])
  np.testing.assert_array_equal(t_irregular.values[2, 0], tgts['t'].values[k])
  assert t.coords['mask'].values.all()
  # lead-time interval (inclusive) for predictions, refused for targets
  window = slice(LEAD[1], LEAD[2])
  assert pl.load_chunk(INIT[:1], window)['t'].sizes['lead_time'] == 2
  with pytest.raises(ValueError, match='Lead time slice not supported'):
    tl.load_chunk(INIT[:1], window)
  # no lead times: init times are valid times
  t0 = tl.load_chunk(INIT[:2])['t']
  assert t0.dims == ('init_time', 'latitude', 'longitude')
  with pytest.raises(KeyError):
    pl.load_chunk(np.array(['2021-01-01'], 'datetime64[ns]'), LEAD)
  # non-contiguous selection copies, in the requested order
  p2 = pl.load_chunk(INIT[[4, 1]], LEAD[[3, 0]])['t']
  np.testing.assert_array_equal(p2.values[0, 1], preds['t'].values[4, 0])


# ---------------------------------------------------------------------------
# run_pipeline
# ---------------------------------------------------------------------------


class OracleAggregator:
  """aggregate_statistics with the CPU oracle (test double)."""

  def __init__(self, reduce_dims):
    self.reduce_dims = list(reduce_dims)
    self.calls = 0

  def aggregate_statistics(self, statistics):
    self.calls += 1
    w = oracle.grid_area_weights(LAT)
    sws, sw = {}, {}
    for name, per_var in statistics.items():
      sws[name], sw[name] = {}, {}
      for var, lazy in per_var.items():
        fn = oracle.DETERMINISTIC_STATISTICS[lazy.kind]
        a, b, dims = oracle.aggregate(
            fn(lazy.predictions.values, lazy.targets.values), lazy.dims,
            self.reduce_dims, weights=[(w, ('latitude',))])
        coords = {d: lazy.coords[d] for d in dims if d in lazy.coords}
        sws[name][var] = xl.DataArray(a, dims, coords=coords, name=var)
        sw[name][var] = xl.DataArray(b, dims, coords=coords, name=var)
    return aggregation.AggregationState(sws, sw)


METRICS = {'rmse': deterministic.RMSE(), 'bias': deterministic.Bias()}


def _monolithic(reduce_dims):
  preds, tgts = _datasets()
  p = array_loaders.PredictionsFromArrays(preds).load_chunk(INIT, LEAD)
  t = array_loaders.TargetsFromArrays(tgts).load_chunk(INIT, LEAD)
  stats = metrics_base.compute_unique_statistics_for_all_metrics(METRICS, p, t)
  return OracleAggregator(reduce_dims).aggregate_statistics(
      stats).metric_values(METRICS)


def _run(reduce_dims, init_chunk, lead_chunk, **kw):
  preds, tgts = _datasets()
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=init_chunk,
                                 lead_time_chunk_size=lead_chunk)
  agg = OracleAggregator(reduce_dims)
  out = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds),
      array_loaders.TargetsFromArrays(tgts), METRICS, agg, **kw)
  return out, agg


@pytest.mark.parametrize('reduce_dims', [
    ['init_time', 'latitude', 'longitude'],
    ['latitude', 'longitude'],
    ['init_time', 'lead_time', 'latitude', 'longitude'],
    ['lead_time', 'latitude', 'longitude'],
])
@pytest.mark.parametrize('chunks', [(1, 1), (2, 3), (4, None), (None, 2)])
@pytest.mark.parametrize('prefetch,lanes', [(0, 1), (2, 1), (2, 3)])
def test_chunked_pipeline_equals_monolithic(reduce_dims, chunks, prefetch,
                                            lanes):
  out, agg = _run(reduce_dims, *chunks, prefetch=prefetch, lanes=lanes,
                  require_output=False)
  state, values = out[None]
  mono = _monolithic(reduce_dims)
  assert set(values) == set(mono) == {'rmse.t', 'rmse.z', 'bias.t', 'bias.z'}
  for k in mono:
    assert values[k].dims == mono[k].dims
    np.testing.assert_allclose(values[k].values, mono[k].values, rtol=1e-12)
    for d in values[k].dims:
      np.testing.assert_array_equal(values[k].coords[d].values,
                                    mono[k].coords[d].values)
  n_init = -(-len(INIT) // (chunks[0] or len(INIT)))
  n_lead = -(-len(LEAD) // (chunks[1] or len(LEAD)))
  assert agg.calls == n_init * n_lead
  assert state.sum_weights['SquaredError']['t'].dims == mono['rmse.t'].dims


def test_pipeline_writes_metrics_and_state_files(tmp_path):
  out_path = str(tmp_path / 'metrics.nc')
  state_path = str(tmp_path / 'state.nc')
  preds, tgts = _datasets()
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=2)
  aggregators = {'time_mean': OracleAggregator(['init_time', 'latitude',
                                                'longitude']),
                 'per_init': OracleAggregator(['latitude', 'longitude'])}
  out = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds),
      array_loaders.TargetsFromArrays(tgts), METRICS, aggregators,
      out_path=out_path, aggregation_state_out_path=state_path)
  assert set(out) == {'time_mean', 'per_init'}
  for name in aggregators:
    values = io_netcdf.open_dataset(str(tmp_path / f'metrics_{name}.nc'))
    for k, v in out[name][1].items():
      np.testing.assert_array_equal(values[k].values, v.values)
      assert values[k].dims == v.dims
    state = pipeline.load_aggregation_state(str(tmp_path / f'state_{name}.nc'))
    again = state.metric_values(METRICS)
    for k, v in out[name][1].items():
      np.testing.assert_allclose(again[k].values, v.values, rtol=1e-15)
  per_init = io_netcdf.open_dataset(str(tmp_path / 'metrics_per_init.nc'))
  np.testing.assert_array_equal(
      per_init['rmse.t'].coords['init_time'].values, INIT)
  np.testing.assert_array_equal(
      per_init['rmse.t'].coords['lead_time'].values, LEAD)
  with pytest.raises(ValueError, match='At least one of'):
    pipeline.run_pipeline(times, None, None, METRICS,
                          aggregators['time_mean'])
  with pytest.raises(ValueError, match="don't match aggregator names"):
    pipeline.run_pipeline(times, None, None, METRICS, aggregators,
                          out_path={'other': 'x.nc'})


def test_pipeline_resumes_from_checkpoint(tmp_path):
  ckpt = str(tmp_path / 'ckpt')
  reduce_dims = ['init_time', 'latitude', 'longitude']

  class Crash(Exception):
    pass

  def crash_after_three(done, total):
    del total
    if done == 3:
      raise Crash()

  with pytest.raises(Crash):
    _run(reduce_dims, 1, None, require_output=False, checkpoint_path=ckpt,
         checkpoint_every=2, progress=crash_after_three, prefetch=0)
  assert os.path.exists(ckpt + '.rank0of1.pkl')
  out, agg = _run(reduce_dims, 1, None, require_output=False,
                  checkpoint_path=ckpt, checkpoint_every=2, prefetch=0)
  assert agg.calls == len(INIT) - 2      # chunks 0, 1 came from the checkpoint
  mono = _monolithic(reduce_dims)
  for k in mono:
    np.testing.assert_allclose(out[None][1][k].values, mono[k].values,
                               rtol=1e-12)
  # a finished run leaves a checkpoint that covers everything
  _, agg = _run(reduce_dims, 1, None, require_output=False,
                checkpoint_path=ckpt, checkpoint_every=2, prefetch=0)
  assert agg.calls == 0


def test_lanes_fold_in_chunk_order_and_surface_errors():
  """Concurrent lanes: results are bit-identical to the sequential run (folded
  in chunk order), and an exception in a lane reaches the caller."""
  rd = ['init_time', 'latitude', 'longitude']
  seq, _ = _run(rd, 1, 1, require_output=False, lanes=1)
  par, agg = _run(rd, 1, 1, require_output=False, lanes=4)
  assert agg.calls == len(INIT) * len(LEAD)
  for k, v in seq[None][1].items():
    assert par[None][1][k].values.tobytes() == v.values.tobytes()

  class Exploding(OracleAggregator):
    def aggregate_statistics(self, statistics):
      if self.calls == 3:
        raise RuntimeError('lane on fire')
      return super().aggregate_statistics(statistics)

  preds, tgts = _datasets()
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=1)
  with pytest.raises(RuntimeError, match='lane on fire'):
    pipeline.run_pipeline(
        times, array_loaders.PredictionsFromArrays(preds),
        array_loaders.TargetsFromArrays(tgts), METRICS, Exploding(rd),
        require_output=False, lanes=2)


def test_trace_and_ready_events(tmp_path, monkeypatch):
  """WBX_PIPELINE_TRACE records the phases of every chunk, and the events a
  loader hands out through take_ready_events are waited for by the evaluation
  (the device cache of TargetsFromArrays issues its uploads without waiting
  when the chunk driver asks it to)."""
  import json
  trace = tmp_path / 'trace.jsonl'
  monkeypatch.setenv('WBX_PIPELINE_TRACE', str(trace))
  waited = []

  class Event:
    def __init__(self, tag):
      self.tag = tag

    def synchronize(self):
      waited.append(self.tag)

  class EventfulTargets(array_loaders.TargetsFromArrays):
    def load_chunk(self, init_times, lead_times=None, reference=None):
      assert self.async_uploads        # set by run_pipeline, reset afterwards
      self._chunk_events.append(Event(str(np.asarray(init_times)[0])))
      return super().load_chunk(init_times, lead_times, reference)

  preds, tgts = _datasets()
  rd = ['init_time', 'latitude', 'longitude']
  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=1)
  loader = EventfulTargets(tgts)
  out = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(preds), loader, METRICS,
      OracleAggregator(rd), require_output=False, lanes=2)
  assert loader.async_uploads is False
  assert len(waited) == len(times) and len(set(waited)) == len(INIT)
  rows = [json.loads(line) for line in trace.read_text().splitlines()]
  phases = {r['phase'] for r in rows}
  assert phases == {'load', 'wait', 'statistics', 'aggregate'}
  assert {r['chunk'] for r in rows} == set(range(len(times)))
  assert all(r['end'] >= r['start'] for r in rows)
  mono = _monolithic(rd)
  for k in mono:
    np.testing.assert_allclose(out[None][1][k].values, mono[k].values,
                               rtol=1e-12)


def test_loader_errors_surface_from_the_prefetch_thread():
  class Broken:
    def load_chunk(self, *args):
      raise RuntimeError('disk on fire')

  times = time_chunks.TimeChunks(INIT, LEAD, init_time_chunk_size=2)
  with pytest.raises(RuntimeError, match='disk on fire'):
    pipeline.run_pipeline(times, Broken(), Broken(), METRICS,
                          OracleAggregator(['latitude']), require_output=False)


# ---------------------------------------------------------------------------
# world_size 2 (gloo)
# ---------------------------------------------------------------------------


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _entry(rank, world_size, port, reduce_dims, tmp):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world_size)
  try:
    out, agg = _run(reduce_dims, 1, 2, out_path=os.path.join(tmp, 'm.nc'))
    np.savez(os.path.join(tmp, f'rank{rank}.npz'), calls=agg.calls,
             **{k: v.values for k, v in out[None][1].items()})
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('reduce_dims', [
    ['init_time', 'latitude', 'longitude'], ['latitude', 'longitude']])
def test_pipeline_two_ranks_gloo(reduce_dims):
  mono = _monolithic(reduce_dims)
  with tempfile.TemporaryDirectory() as tmp:
    mp.spawn(_entry, args=(2, _free_port(), reduce_dims, tmp), nprocs=2,
             join=True)
    res = [np.load(os.path.join(tmp, f'rank{r}.npz')) for r in range(2)]
    written = io_netcdf.open_dataset(os.path.join(tmp, 'm.nc'))
  assert sum(int(r['calls']) for r in res) == 12       # 6 x 2 chunks, shared
  assert all(int(r['calls']) == 6 for r in res)
  for r in res:
    for k in mono:
      np.testing.assert_allclose(r[k], mono[k].values, rtol=1e-12)
  for k in mono:
    np.testing.assert_allclose(written[k].values, mono[k].values, rtol=1e-12)


# ---------------------------------------------------------------------------
# NetCDF-3 round trip (io_netcdf.py)
# ---------------------------------------------------------------------------


def test_netcdf_round_trip_of_every_coordinate_kind(tmp_path):
  lead = LEAD[:3]
  init = INIT[:2]
  a = xl.DataArray(
      np.random.default_rng(0).random((2, 3, 4)),
      ('init_time', 'lead_time', 'level'),
      coords={'init_time': init, 'lead_time': lead,
              'level': np.array([500, 700, 850, 1000]),
              'valid_time': xl.DataArray(init[:, None] + lead[None, :],
                                         ('init_time', 'lead_time'))},
      name='rmse.geopotential')
  scalar = xl.DataArray(np.float64(3.5), ())
  regions = xl.DataArray(np.array([1.0, np.nan]), ('region',),
                         coords={'region': np.array(['global', 'tropics'])})
  flags = xl.DataArray(np.array([True, False, True]), ('k',))
  f32 = xl.DataArray(np.arange(4, dtype=np.float32), ('level',),
                     coords={'level': np.array([500, 700, 850, 1000])})
  ds = xl.Dataset({'rmse.geopotential': a,
                   'SquaredError#z#sum_weights': scalar, 'by_region': regions,
                   'flags': flags, 'f32': f32})
  path = str(tmp_path / 'sub' / 'x.nc')
  io_netcdf.to_netcdf(ds, path)
  back = io_netcdf.open_dataset(path)
  assert set(back) == set(ds)
  for k in ds:
    assert back[k].dims == ds[k].dims
    np.testing.assert_array_equal(back[k].values, ds[k].values)
    assert back[k].values.dtype == ds[k].values.dtype
  z = back['rmse.geopotential']
  assert z.coords['init_time'].values.dtype == np.dtype('datetime64[ns]')
  assert z.coords['lead_time'].values.dtype == np.dtype('timedelta64[ns]')
  np.testing.assert_array_equal(z.coords['valid_time'].values,
                                a.coords['valid_time'].values)
  assert back['by_region'].coords['region'].values.tolist() == [
      'global', 'tropics']
  assert not os.path.exists(path + '.tmp') and len(os.listdir(
      os.path.dirname(path))) == 1            # atomic write left no temp file
  # a dimension used with two sizes cannot be written
  bad = xl.Dataset({'a': xl.DataArray(np.zeros(2), ('x',)),
                    'b': xl.DataArray(np.zeros(3), ('x',))})
  with pytest.raises(ValueError, match='has sizes'):
    io_netcdf.to_netcdf(bad, str(tmp_path / 'bad.nc'))
  assert not os.path.exists(str(tmp_path / 'bad.nc'))


def test_window_view_only_for_lattices():
  payload = np.arange(20 * 3, dtype=np.float32).reshape(20, 3)
  lattice = 2 + 3 * np.arange(4)[:, None] + np.arange(5)[None, :]
  view = array_loaders._window_view(payload, 0, lattice)
  assert view.shape == (4, 5, 3) and np.shares_memory(view, payload)
  np.testing.assert_array_equal(view, payload[lattice])
  assert not view.flags.writeable
  single = array_loaders._window_view(payload, 0, np.array([[7]]))
  np.testing.assert_array_equal(single, payload[[[7]]])
  irregular = lattice.copy()
  irregular[2, 3] += 1
  assert array_loaders._window_view(payload, 0, irregular) is None
  assert array_loaders._window_view(payload, 0, lattice[::-1]) is None
  # along a middle axis
  cube = np.arange(2 * 12 * 3, dtype=np.float32).reshape(2, 12, 3)
  pos = 1 + 2 * np.arange(3)[:, None] + np.arange(4)[None, :]
  np.testing.assert_array_equal(array_loaders._window_view(cube, 1, pos),
                                cube[:, pos])


# ---------------------------------------------------------------------------
# TimeChunks against the reference's own class (build container only)
# ---------------------------------------------------------------------------

_REF_TIME_CHUNKS = '/root/reference/weatherbenchX/time_chunks.py'


@pytest.mark.skipif(not os.path.exists(_REF_TIME_CHUNKS),
                    reason='the reference tree is only present in the build '
                           'container')
@pytest.mark.parametrize('init_chunk,lead_chunk', [
    (None, None), (1, None), (3, 2), (4, 1), (7, 5), (100, 100)])
@pytest.mark.parametrize('lead_kind', ['exact', 'slice'])
def test_time_chunks_equal_the_reference_class(init_chunk, lead_chunk,
                                               lead_kind):
  """time_chunks.py:37-202 is pure NumPy, so the reference class itself is
  imported (from its file, not as a package) and iterated side by side."""
  import importlib.util
  spec = importlib.util.spec_from_file_location('_ref_time_chunks',
                                                _REF_TIME_CHUNKS)
  ref = importlib.util.module_from_spec(spec)
  sys.dont_write_bytecode = True
  spec.loader.exec_module(ref)
  init_times = np.arange('2020-01-01T00', '2020-01-04T12',
                         np.timedelta64(12, 'h'), dtype='datetime64[ns]')
  if lead_kind == 'exact':
    lead_times = np.arange(0, 30, 6, dtype='timedelta64[h]').astype(
        'timedelta64[ns]')
  else:
    if lead_chunk is not None:
      pytest.skip('a lead_time slice cannot be chunked')
    lead_times = slice(np.timedelta64(0, 'h'), np.timedelta64(24, 'h'))
  kwargs = dict(init_time_chunk_size=init_chunk,
                lead_time_chunk_size=lead_chunk)
  theirs = ref.TimeChunks(init_times, lead_times, **kwargs)
  ours = time_chunks.TimeChunks(init_times, lead_times, **kwargs)
  assert len(ours) == len(theirs)
  for (i_a, l_a), (i_b, l_b) in zip(ours, theirs):
    np.testing.assert_array_equal(i_a, i_b)
    if isinstance(l_b, slice):
      assert l_a == l_b
    else:
      np.testing.assert_array_equal(l_a, l_b)
  for idx in range(len(theirs)):
    a, b = ours[idx], theirs[idx]
    np.testing.assert_array_equal(a[0], b[0])
  if lead_kind == 'exact':
    # (for a lead_time slice the reference's iter_with_chunk_offsets raises a
    # TypeError of its own: None * int, time_chunks.py:200)
    for (off_a, _), (off_b, _) in zip(ours.iter_with_chunk_offsets(),
                                      theirs.iter_with_chunk_offsets()):
      assert (off_a.init_time, off_a.lead_time) == (off_b.init_time,
                                                    off_b.lead_time)
