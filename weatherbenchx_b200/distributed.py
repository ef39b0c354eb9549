"""Multi-GPU evaluation: shard (variable, init_time) units over ranks, all-reduce
the AggregationStates.

The reference distributes with Apache Beam: per-chunk states are keyed and
summed by ``CombinePerKey(CombiningSum)`` (beam_pipeline.py:509-510,
aggregation.py:27-60), and when ``init_time`` is a kept dim the per-chunk
results are concatenated (beam_pipeline.py:253-319).  Here there is one process
per GPU (torchrun); every rank aggregates its own units into a local
AggregationState and ONE collective combines them:

  * identical state structure on every rank  -> the sums are packed into one
    float64 buffer and ``all_reduce(SUM)``-ed (NCCL over NVLink on GPUs, gloo
    on CPU-only hosts); a few KB..MB, latency-bound;
  * ranks hold different keys or different coordinates along a kept dim
    (variable sharding, kept ``init_time``) -> the zero-filled outer-join sum
    of the reference (``combining_sum``) applied to the gathered states, which
    for disjoint coordinates is exactly the Beam concatenation.

NaN semantics survive either way (NaN + x = NaN), as in the reference.
"""

from __future__ import annotations

import hashlib
import pickle
from typing import Any, Callable, Iterable, Mapping, Sequence

import numpy as np

from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import xarray_lite as xl


def _dist():
  import torch.distributed as dist  # pylint: disable=g-import-not-at-top
  return dist


def world() -> tuple:
  """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
  dist = _dist()
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def shard_units(units: Sequence[Any], rank: int | None = None,
                world_size: int | None = None) -> list:
  """Contiguous block partition of evaluation units (e.g. (variable,
  init_time chunk) pairs) -- never splits latitude / longitude."""
  if rank is None or world_size is None:
    rank, world_size = world()
  n = len(units)
  lo = (rank * n) // world_size
  hi = ((rank + 1) * n) // world_size
  return list(units[lo:hi])


# ---------------------------------------------------------------------------
# packing
# ---------------------------------------------------------------------------


def _leaves(tree, prefix=()):
  if isinstance(tree, Mapping):
    for k in tree:
      yield from _leaves(tree[k], prefix + (k,))
  elif tree is not None:
    yield prefix, tree


def state_layout(state: aggregation.AggregationState) -> list:
  """[(path, dims, shape, coords)] of every leaf, in deterministic order."""
  if state.sum_weighted_statistics is None:
    return []
  out = []
  for path, da in sorted(_leaves(state.sum_weighted_statistics),
                         key=lambda kv: tuple(map(str, kv[0]))):
    coords = {k: (v.dims, v.to_numpy()) for k, v in da.coords.items()}
    out.append((path, tuple(da.dims), tuple(da.shape), coords))
  return out


def layout_digest(layout: list) -> str:
  h = hashlib.sha256()
  for path, dims, shape, coords in layout:
    h.update(repr((path, dims, shape)).encode())
    for k in sorted(coords, key=str):
      h.update(str(k).encode())
      h.update(np.ascontiguousarray(coords[k][1]).tobytes())
  return h.hexdigest()


def _get(tree, path):
  for k in path:
    tree = tree[k]
  return tree


def pack_state(state: aggregation.AggregationState, layout: list) -> np.ndarray:
  """[sum_weighted_statistics || sum_weights] of all leaves as one f64 vector."""
  parts = []
  for tree in (state.sum_weighted_statistics, state.sum_weights):
    for path, _, _, _ in layout:
      parts.append(np.asarray(_get(tree, path).to_numpy(),
                              dtype=np.float64).reshape(-1))
  return np.concatenate(parts) if parts else np.zeros(0)


def unpack_state(flat: np.ndarray, layout: list) -> aggregation.AggregationState:
  trees: list = [{}, {}]
  off = 0
  for tree in trees:
    for path, dims, shape, coords in layout:
      n = int(np.prod(shape, dtype=np.int64)) if shape else 1
      da = xl.DataArray(
          flat[off:off + n].reshape(shape).copy(), dims,
          coords={k: xl.DataArray(v, d) for k, (d, v) in coords.items()},
          name=path[-1])
      off += n
      node = tree
      for k in path[:-1]:
        node = node.setdefault(k, {})
      node[path[-1]] = da
  return aggregation.AggregationState(trees[0], trees[1])


# ---------------------------------------------------------------------------
# collectives
# ---------------------------------------------------------------------------


def _same_layout_everywhere(digest: str, n_values: int, group, backend,
                           device) -> bool:
  """True if every rank of the group holds the same state layout: ONE small
  fixed-size all-reduce(MAX) over [h, -h, n, -n] words of the layout digest
  (max(x) == -max(-x) for every word iff all ranks agree) -- no pickling, no
  object gather."""
  import torch  # pylint: disable=g-import-not-at-top
  dist = _dist()
  words = [int(digest[i:i + 12], 16) for i in range(0, 48, 12)] + [n_values]
  probe = torch.tensor(words + [-w for w in words], dtype=torch.int64)
  if backend == 'nccl':
    probe = probe.to(device)
  dist.all_reduce(probe, op=dist.ReduceOp.MAX, group=group)
  probe = probe.cpu().tolist()
  k = len(words)
  return all(probe[i] == -probe[k + i] for i in range(k))


def device_for_local_rank(local_rank: int, local_world: int,
                          visible: int | None = None) -> int:
  """CUDA device index for rank `local_rank` of `local_world` ranks on a node.

  With fewer ranks than visible GPUs the ranks are spread evenly over the
  device indices (rank r -> r * (visible // local_world)).  Evaluation is fed
  from host memory, so the placement that matters is the PCIe one, not NVLink:
  on the 8 x B200 boxes of this pool the GPUs 0-3 and 4-7 each share one host
  uplink (measured, profiles/h2d_ceiling_r2_box8_n8.json: four neighbouring
  GPUs copy 115 GB/s together, GPUs 0,2,4,6 copy 213 GB/s), and spreading the
  ranks puts as few of them as possible behind the same uplink."""
  if visible is None:
    import torch  # pylint: disable=g-import-not-at-top
    visible = torch.cuda.device_count()
  if local_world <= 0 or visible < 2 * local_world:
    return local_rank % max(visible, 1)
  return local_rank * (visible // local_world)


def all_reduce_state(state: aggregation.AggregationState, group=None,
                     device=None) -> aggregation.AggregationState:
  """Sum of the AggregationStates of all ranks (every rank gets the result).

  Two collectives when every rank holds the same structure (the usual case:
  reduced init_time): a 10-word agreement probe and the packed float64
  all-reduce of [sum_weighted_statistics || sum_weights]."""
  import torch  # pylint: disable=g-import-not-at-top
  dist = _dist()
  rank, size = world()
  if size == 1:
    return state
  del rank
  layout = state_layout(state)
  digest = layout_digest(layout)
  backend = dist.get_backend(group)
  dev = None
  if backend == 'nccl':
    dev = torch.device('cuda', torch.cuda.current_device()
                       if device is None else device)
  n_values = 2 * sum(int(np.prod(shape, dtype=np.int64)) if shape else 1
                     for _, _, shape, _ in layout)
  if _same_layout_everywhere(digest, n_values, group, backend, dev) and layout:
    # fast path: one packed float64 all-reduce.
    flat = pack_state(state, layout)
    if backend == 'nccl':
      buf = torch.from_numpy(flat).to(dev)
    else:
      buf = torch.from_numpy(flat.copy())
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack_state(buf.cpu().numpy(), layout)
  # general path: structures differ (variable sharding, kept time dims).
  gathered = [None] * size
  dist.all_gather_object(gathered, pickle.dumps(_to_plain(state)), group=group)
  states = [_from_plain(pickle.loads(g)) for g in gathered]
  return combine_states(states)


def combine_states(states: Iterable[aggregation.AggregationState]
                   ) -> aggregation.AggregationState:
  """Outer-join sum that also unions keys (statistics / variables) -- the
  CombinePerKey of the reference over states with different key sets."""
  states = [s for s in states if s.sum_weighted_statistics is not None]
  if not states:
    return aggregation.AggregationState.zero()

  def merge(trees):
    if all(isinstance(t, Mapping) for t in trees):
      keys: list = []
      for t in trees:
        for k in t:
          if k not in keys:
            keys.append(k)
      return {k: merge([t[k] for t in trees if k in t]) for k in keys}
    return aggregation.combining_sum(list(trees))

  return aggregation.AggregationState(
      merge([s.sum_weighted_statistics for s in states]),
      merge([s.sum_weights for s in states]))


def _to_plain(state):
  def conv(tree):
    if isinstance(tree, Mapping):
      return {k: conv(v) for k, v in tree.items()}
    if tree is None:
      return None
    return ('__da__', tree.to_numpy(), tuple(tree.dims),
            {k: (v.dims, v.to_numpy()) for k, v in tree.coords.items()},
            tree.name)
  return (conv(state.sum_weighted_statistics), conv(state.sum_weights))


def _from_plain(plain):
  def conv(tree):
    if isinstance(tree, dict):
      return {k: conv(v) for k, v in tree.items()}
    if tree is None:
      return None
    _, data, dims, coords, name = tree
    return xl.DataArray(data, dims, coords={
        k: xl.DataArray(v, d) for k, (d, v) in coords.items()}, name=name)
  return aggregation.AggregationState(conv(plain[0]), conv(plain[1]))


# ---------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------


def evaluate_sharded(metrics: Mapping[str, Any],
                     aggregator: aggregation.Aggregator,
                     units: Sequence[Any],
                     load_unit: Callable[[Any], tuple],
                     group=None) -> xl.Dataset:
  """Evaluates ``metrics`` over all ``units`` with the ranks of the process
  group: each rank loads and aggregates its contiguous share of the units
  (``load_unit(unit) -> (predictions, targets)``, typically one variable and
  one init_time chunk), states are summed locally and all-reduced once.
  """
  from weatherbenchx_b200.metrics import base as metrics_base  # pylint: disable=g-import-not-at-top
  local = aggregation.AggregationState.zero()
  for unit in shard_units(units):
    predictions, targets = load_unit(unit)
    stats = metrics_base.compute_unique_statistics_for_all_metrics(
        metrics, predictions, targets)
    local = combine_states([local, aggregator.aggregate_statistics(stats)])
  total = all_reduce_state(local, group=group)
  return total.metric_values(metrics)
