"""On-box evaluation driver: time chunks -> statistics -> aggregation -> files.

Replaces, for runs on one multi-GPU node, the Apache Beam pipeline of the
reference (/root/reference/weatherbenchX/beam_pipeline.py:490-560
``define_pipeline``) with the same arguments and the same result files:

  reference stage (beam_pipeline.py)                   here
  --------------------------------------------------   --------------------------
  Create(times.iter_with_chunk_offsets())  :515        chunk indices, block-
                                                       partitioned over the ranks
  LoadPredictionsAndTargets           :62-118          loader thread, ``prefetch``
                                                       chunks ahead of the GPU
  ComputeStatisticsAggregateAndPrepareForCombine       Aggregator.aggregate_
                                      :140-250         statistics (fused launches)
  CombinePerKey(CombiningSum)         :535             per-(aggregator, statistic,
  ConcatPerStatisticPerVariable       :253-322         variable) outer-join sum:
                                                       blocks of a kept init_time /
                                                       lead_time axis land at their
                                                       coordinates, reduced ones add
  (shuffle between workers)                            ONE collective at the end:
                                                       packed float64 all-reduce,
                                                       or gather + outer-join sum
                                                       when ranks hold different
                                                       coordinates
  ReconstructAggregationState         :325-365         AggregationState per
                                                       aggregator
  ComputeMetrics / WriteMetrics /     :368-448         metric_values + NetCDF-3
  WriteAggregationState                                files (atomic rename)

One process per GPU; under ``torchrun`` every rank runs ``run_pipeline`` with
the same arguments.  With ``checkpoint_path`` each rank periodically saves its
partial sums and the chunk indices they cover; a restarted run skips them.
"""

from __future__ import annotations

import os
import pickle
import queue
import tempfile
import threading
import time
from typing import Callable, Mapping, Optional

import numpy as np

from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import distributed
from weatherbenchx_b200 import io_netcdf
from weatherbenchx_b200 import time_chunks as time_chunks_lib
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base as metrics_base


def _resolve_out_path(out_path, agg_name):
  """beam_pipeline.py:391-402."""
  if isinstance(out_path, str):
    if agg_name is None:
      return out_path
    base, ext = os.path.splitext(out_path)
    return f'{base}_{agg_name}{ext}'
  return out_path[agg_name]


# Optional event trace of a run (WBX_PIPELINE_TRACE=<path>): one JSON line per
# phase of every chunk -- (thread, chunk, phase, start, end) in seconds -- to
# see how loading, planning and the library calls of the lanes overlap.
_TRACE: list = []
_TRACE_LOCK = threading.Lock()


def _trace(chunk, phase: str, start: float) -> None:
  if os.environ.get('WBX_PIPELINE_TRACE'):
    with _TRACE_LOCK:
      _TRACE.append((threading.current_thread().name, int(chunk), phase,
                     start, time.perf_counter()))


def _dump_trace() -> None:
  path = os.environ.get('WBX_PIPELINE_TRACE')
  if not path:
    return
  import json  # pylint: disable=g-import-not-at-top
  with _TRACE_LOCK:
    rows, _TRACE[:] = list(_TRACE), []
  with open(path, 'a') as f:
    for thread, chunk, phase, start, end in rows:
      f.write(json.dumps({'thread': thread, 'chunk': chunk, 'phase': phase,
                          'start': start, 'end': end}) + '\n')


def _current_cuda_device():
  """Index of the calling thread's CUDA device, or None without a GPU."""
  try:
    import torch  # pylint: disable=g-import-not-at-top
    if torch.cuda.is_available():
      return torch.cuda.current_device()
  except ImportError:
    pass
  return None


def _bind_cuda_device(device) -> None:
  """The CUDA current device is per host thread and starts at 0: a thread that
  works for rank r has to be put on the device of the thread that made it."""
  if device is not None:
    import torch  # pylint: disable=g-import-not-at-top
    torch.cuda.set_device(device)


class _Prefetcher:
  """Loads chunks on a background thread, ``depth`` ahead of the consumer."""

  def __init__(self, times, indices, predictions_loader, targets_loader,
               depth: int, setup_fn=None):
    self._times = times
    self._indices = list(indices)
    self._loaders = (predictions_loader, targets_loader)
    self._setup_fn = setup_fn
    self._queue: queue.Queue = queue.Queue(maxsize=max(depth, 1))
    self._thread = None
    self._device = _current_cuda_device()
    self.load_seconds = 0.0
    if depth > 0:
      self._thread = threading.Thread(target=self._work, daemon=True)
      self._thread.start()

  def _load(self, index):
    start = time.perf_counter()
    init_times, lead_times = self._times[index]
    predictions_loader, targets_loader = self._loaders
    # targets first: they may serve as the reference of the predictions
    # loader (beam_pipeline.py:93-105)
    targets = targets_loader.load_chunk(init_times, lead_times)
    predictions = predictions_loader.load_chunk(init_times, lead_times, targets)
    # uploads a loader issued without waiting (TargetsFromArrays device cache):
    # the evaluation waits for them, not this thread
    events = []
    for loader in (targets_loader, predictions_loader):
      take = getattr(loader, 'take_ready_events', None)
      if take is not None:
        events += take()
    self.load_seconds += time.perf_counter() - start
    _trace(index, 'load', start)
    return index, predictions, targets, events

  def _work(self):
    try:
      _bind_cuda_device(self._device)
      if self._setup_fn is not None:
        self._setup_fn()
      for index in self._indices:
        self._queue.put(self._load(index))
      self._queue.put(None)
    except BaseException as e:  # pylint: disable=broad-except
      self._queue.put(e)

  def __iter__(self):
    if self._thread is None:
      if self._setup_fn is not None:
        self._setup_fn()
      for index in self._indices:
        yield self._load(index)
      return
    while True:
      item = self._queue.get()
      if item is None:
        return
      if isinstance(item, BaseException):
        raise item
      yield item


def _evaluate_in_lanes(items, evaluate, lanes: int):
  """Yields evaluate(item) for every item, in item order; with lanes > 1 the
  calls run on ``lanes`` worker threads (engine lane i + 1 each)."""
  if lanes <= 1:
    for item in items:
      yield evaluate(item)
    return
  from weatherbenchx_b200 import _cabi  # pylint: disable=g-import-not-at-top
  inbox: queue.Queue = queue.Queue(maxsize=lanes)
  outbox: queue.Queue = queue.Queue()
  stop = threading.Event()
  device = _current_cuda_device()

  def put(q, value):
    while not stop.is_set():
      try:
        q.put(value, timeout=0.05)
        return True
      except queue.Full:
        continue
    return False

  def worker(lane):
    _bind_cuda_device(device)
    _cabi.set_thread_lane(lane)
    while not stop.is_set():
      try:
        job = inbox.get(timeout=0.05)
      except queue.Empty:
        continue
      if job is None:  # no more work: leave without waiting for `stop`
        return
      seq, item = job
      try:
        outbox.put((seq, evaluate(item), None))
      except BaseException as e:  # pylint: disable=broad-except
        outbox.put((seq, None, e))

  def feeder():
    seq = 0
    _bind_cuda_device(device)
    try:
      for item in items:
        if not put(inbox, (seq, item)):
          return
        seq += 1
      outbox.put(('end', seq, None))
    except BaseException as e:  # pylint: disable=broad-except
      outbox.put(('end', seq, e))
    for _ in range(lanes):
      put(inbox, None)

  threads = [threading.Thread(target=worker, args=(i + 1,), daemon=True)
             for i in range(lanes)]
  threads.append(threading.Thread(target=feeder, daemon=True))
  for t in threads:
    t.start()
  ready: dict = {}
  next_seq, total = 0, None
  try:
    while total is None or next_seq < total:
      seq, result, error = outbox.get()
      if error is not None:
        raise error
      if seq == 'end':
        total = result
        continue
      ready[seq] = result
      while next_seq in ready:
        yield ready.pop(next_seq)
        next_seq += 1
  finally:
    stop.set()
    for t in threads:
      t.join()


def _flatten(state: aggregation.AggregationState) -> dict:
  """{(type, statistic, variable): DataArray} of one AggregationState."""
  out = {}
  if state.sum_weighted_statistics is None:
    return out
  for kind, tree in (('sum_weighted_statistics', state.sum_weighted_statistics),
                     ('sum_weights', state.sum_weights)):
    for stat_name, per_var in tree.items():
      for var, da in per_var.items():
        out[(kind, stat_name, var)] = da
  return out


def reconstruct_aggregation_state(pairs) -> aggregation.AggregationState:
  """AggregationState from ((type, statistic, variable), DataArray) pairs
  (beam_pipeline.py:325-352)."""
  trees = {'sum_weighted_statistics': {}, 'sum_weights': {}}
  for (kind, stat_name, var), da in pairs:
    trees[kind].setdefault(stat_name, {})[var] = da
  if not trees['sum_weighted_statistics']:
    return aggregation.AggregationState.zero()
  return aggregation.AggregationState(trees['sum_weighted_statistics'],
                                      trees['sum_weights'])


class _Accumulator:
  """Outer-join running sum per (aggregator, type, statistic, variable).

  Blocks are buffered and folded ``flush_every`` at a time in one pass
  (aggregation.combining_sum), so a kept init_time axis assembled from
  thousands of chunks is not re-copied per chunk.
  """

  def __init__(self, flush_every: int = 64):
    self._total: dict = {}
    self._pending: dict = {}
    self._flush_every = flush_every

  def add(self, agg_name, state: aggregation.AggregationState):
    for key, da in _flatten(state).items():
      blocks = self._pending.setdefault((agg_name,) + key, [])
      blocks.append(da)
      if len(blocks) >= self._flush_every:
        self._fold((agg_name,) + key)

  def _fold(self, key):
    blocks = self._pending.pop(key, [])
    if key in self._total:
      blocks = [self._total[key]] + blocks
    if blocks:
      self._total[key] = aggregation.combining_sum(blocks)

  def totals(self) -> dict:
    for key in list(self._pending):
      self._fold(key)
    return self._total

  def states(self, agg_names) -> dict:
    totals = self.totals()
    return {name: reconstruct_aggregation_state(
        (key[1:], da) for key, da in totals.items() if key[0] == name)
            for name in agg_names}

  # -- checkpoint ------------------------------------------------------------

  def dump(self) -> dict:
    return {key: _plain(da) for key, da in self.totals().items()}

  def load(self, plain: Mapping):
    for key, item in plain.items():
      self._total[key] = _unplain(item)


def _plain(da: xl.DataArray) -> tuple:
  return (da.to_numpy(), tuple(da.dims),
          {k: (v.dims, v.to_numpy()) for k, v in da.coords.items()}, da.name)


def _unplain(item: tuple) -> xl.DataArray:
  data, dims, coords, name = item
  return xl.DataArray(data, dims, coords={
      k: xl.DataArray(v, d) for k, (d, v) in coords.items()}, name=name)


def _fingerprint(times, metrics, aggregators) -> str:
  """Identifies an evaluation for checkpoint resume: the unique names of the
  statistics behind every metric, the aggregator settings and the init / lead
  times with their chunking.  A checkpoint with another fingerprint holds
  partial sums of a different computation and must not be merged."""
  import hashlib  # pylint: disable=g-import-not-at-top
  h = hashlib.sha256()
  for name in sorted(metrics, key=str):
    metric = metrics[name]
    stats = sorted(s.unique_name for s in metric.statistics.values())
    h.update(repr((str(name), type(metric).__name__, stats)).encode())
  for name in sorted(aggregators, key=str):
    agg = aggregators[name]
    h.update(repr((
        str(name), type(agg).__name__,
        sorted(map(str, getattr(agg, 'reduce_dims', ()))),
        [(type(b).__name__, getattr(b, 'bin_dim_name', None))
         for b in getattr(agg, 'bin_by', None) or []],
        [type(w).__name__ for w in getattr(agg, 'weigh_by', None) or []],
        bool(getattr(agg, 'masked', False)),
        bool(getattr(agg, 'skipna', False)))).encode())
  for attr in ('init_times', 'lead_times'):
    values = getattr(times, attr, None)
    if values is not None and not isinstance(values, slice):
      h.update(attr.lstrip('_').encode())
      h.update(np.ascontiguousarray(np.asarray(values)).tobytes())
    elif isinstance(values, slice):
      h.update(repr(values).encode())
  h.update(repr((len(times),
                 getattr(times, 'init_time_chunk_size', None) or
                 getattr(times, '_init_time_chunk_size', None),
                 getattr(times, 'lead_time_chunk_size', None) or
                 getattr(times, '_lead_time_chunk_size', None))).encode())
  return h.hexdigest()


def _atomic_pickle(path: str, payload) -> None:
  directory = os.path.dirname(os.path.abspath(path)) or '.'
  os.makedirs(directory, exist_ok=True)
  fd, tmp = tempfile.mkstemp(dir=directory, suffix='.tmp')
  with os.fdopen(fd, 'wb') as f:
    pickle.dump(payload, f, protocol=pickle.HIGHEST_PROTOCOL)
  os.replace(tmp, path)


def run_pipeline(
    times: time_chunks_lib.TimeChunks,
    predictions_loader,
    targets_loader,
    metrics: Mapping[str, metrics_base.Metric],
    aggregator: aggregation.Aggregator | Mapping[str, aggregation.Aggregator],
    out_path: str | Mapping[str, str] | None = None,
    aggregation_state_out_path: str | Mapping[str, str] | None = None,
    setup_fn: Optional[Callable[[], None]] = None,
    *,
    checkpoint_path: str | None = None,
    checkpoint_every: int = 0,
    prefetch: int = 2,
    lanes: int = 1,
    group=None,
    progress: Optional[Callable[[int, int], None]] = None,
    require_output: bool = True,
    shard: bool = True,
) -> dict:
  """Evaluates ``metrics`` over every chunk of ``times``.

  The first nine arguments are those of the reference's ``define_pipeline``
  (without the Beam ``root``).  Returns, on every rank,
  ``{aggregator_name: (AggregationState, metric values Dataset)}`` with
  ``None`` as the name of a single unnamed Aggregator; rank 0 writes the files.

  Args:
    times: TimeChunks to evaluate.
    predictions_loader / targets_loader: objects with
      ``load_chunk(init_times, lead_times, reference)``.
    metrics: name -> Metric.
    aggregator: one Aggregator or a mapping name -> Aggregator.
    out_path: NetCDF path of the metric values (or mapping per aggregator; a
      single path gets ``_<aggregator name>`` appended per named aggregator).
    aggregation_state_out_path: same, for the final AggregationState
      (``AggregationState.to_dataset`` form).
    setup_fn: called once in the loading thread before the first chunk.
    checkpoint_path: prefix of per-rank checkpoint files.
    checkpoint_every: save after this many chunks (0 = only never).
    prefetch: chunks loaded ahead of the GPU (0 = load synchronously).
    lanes: chunks evaluated concurrently, each on its own thread and engine
      context.  While one lane waits inside the library for its host->device
      copies and kernels (the GIL is released there), another plans the next
      chunk, which hides the per-chunk Python work behind the PCIe transfer.
      Results are folded in chunk order, so the output does not depend on it.
    group: torch.distributed process group (default: the world).
    progress: optional callback (chunks done on this rank, chunks of this rank).
    require_output: the reference insists on at least one output path
      (beam_pipeline.py:541-545); pass False to only get the return value.
    shard: False makes this process evaluate EVERY chunk on its own, with no
      collective, even inside an initialised process group (the monolithic
      run a sharded result is compared with).
  """
  if isinstance(aggregator, Mapping):
    aggregators = dict(aggregator)
    for paths, what in ((out_path, 'out_path'),
                        (aggregation_state_out_path,
                         'aggregation_state_out_path')):
      if isinstance(paths, Mapping) and paths.keys() != aggregators.keys():
        raise ValueError(f"Keys of {what} don't match aggregator names.")
  else:
    aggregators = {None: aggregator}
  if require_output and out_path is None and aggregation_state_out_path is None:
    raise ValueError(
        'At least one of (metrics) out_path or aggregation_state_out_path must '
        'be specified.')

  rank, world_size = distributed.world() if shard else (0, 1)
  mine = distributed.shard_units(list(range(len(times))), rank, world_size)
  acc = _Accumulator()
  done: list = []
  ckpt_file = None
  fingerprint = _fingerprint(times, metrics, aggregators)
  if checkpoint_path is not None:
    ckpt_file = f'{checkpoint_path}.rank{rank}of{world_size}.pkl'
    if os.path.exists(ckpt_file):
      with open(ckpt_file, 'rb') as f:
        saved = pickle.load(f)
      if saved.get('fingerprint') != fingerprint:
        raise ValueError(
            f'{ckpt_file} was written by a different evaluation (other '
            'metrics, aggregators, init / lead times or chunking); partial '
            'sums of different runs must not be merged -- remove it or use '
            'another checkpoint_path')
      if saved.get('n_chunks') == len(times) and set(
          saved['done']) <= set(mine):
        acc.load(saved['acc'])
        done = list(saved['done'])
  todo = [i for i in mine if i not in set(done)]

  async_loaders = [l for l in (targets_loader, predictions_loader)
                   if getattr(l, 'async_uploads', None) is False]
  for l in async_loaders:
    l.async_uploads = True
  loader = _Prefetcher(times, todo, predictions_loader, targets_loader,
                       max(prefetch, lanes if lanes > 1 else 0), setup_fn)

  def evaluate(item):
    index, predictions, targets, events = item
    start = time.perf_counter()
    for event in events:
      event.synchronize()
    _trace(index, 'wait', start)
    start = time.perf_counter()
    statistics = metrics_base.compute_unique_statistics_for_all_metrics(
        metrics, predictions, targets)
    _trace(index, 'statistics', start)
    start = time.perf_counter()
    states = [(name, agg.aggregate_statistics(statistics))
              for name, agg in aggregators.items()]
    _trace(index, 'aggregate', start)
    return index, states

  try:
    since_ckpt = 0
    for index, states in _evaluate_in_lanes(loader, evaluate, lanes):
      for name, state in states:
        acc.add(name, state)
      done.append(index)
      since_ckpt += 1
      if progress is not None:
        progress(len(done), len(mine))
      if ckpt_file and checkpoint_every and since_ckpt >= checkpoint_every:
        _atomic_pickle(ckpt_file, {'n_chunks': len(times), 'done': done,
                                   'acc': acc.dump(),
                                   'fingerprint': fingerprint})
        since_ckpt = 0
    if ckpt_file and since_ckpt:
      _atomic_pickle(ckpt_file, {'n_chunks': len(times), 'done': done,
                                 'acc': acc.dump(), 'fingerprint': fingerprint})
  finally:
    for l in async_loaders:
      l.async_uploads = False
  _dump_trace()
  results = {}
  local = acc.states(aggregators)
  for name in aggregators:
    state = (distributed.all_reduce_state(local[name], group=group)
             if shard else local[name])
    values = (state.metric_values(metrics)
              if state.sum_weighted_statistics is not None else xl.Dataset())
    results[name] = (state, values)
    if rank != 0 or not (shard or distributed.world()[0] == 0):
      continue
    if out_path is not None:
      io_netcdf.to_netcdf(values, _resolve_out_path(out_path, name))
    if aggregation_state_out_path is not None:
      io_netcdf.to_netcdf(
          state.to_dataset(),
          _resolve_out_path(aggregation_state_out_path, name))
  return results


def load_aggregation_state(path: str) -> aggregation.AggregationState:
  """Reads a file written through ``aggregation_state_out_path``."""
  return aggregation.AggregationState.from_dataset(io_netcdf.open_dataset(path))
