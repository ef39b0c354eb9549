"""AggregationState algebra: this package next to the reference's own class.

    python tests/golden/compare_state_algebra.py        (build container only)

TEST INFRASTRUCTURE.  Imports the UNMODIFIED /root/reference/weatherbenchX/
aggregation.py through the stand-in modules of ``reference_runtime.py`` (in
this process only -- that is why it is a script run in a subprocess by
tests/test_host_logic.py and not a test module) and drives both
``AggregationState`` classes with the same randomly generated nested states:
``sum`` / ``+`` with zero states and with differing coordinates (the
zero-filled outer join of ``combining_sum``, aggregation.py:27-60), ``zero``
handling (:84-110), ``mean_statistics`` (:112-121), ``sum_along_dims``
(:150-175), ``dot`` (:177-181), ``map`` / ``map_multi`` (:183-202).  The
container arithmetic underneath is the stand-in's on both sides; what this pins
is the control flow of the reference class (which states are skipped, how the
trees are walked, what a zero state does).  Prints ``state algebra ok``.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_runtime  # noqa: E402  pylint: disable=g-import-not-at-top

xr = reference_runtime.install()
from weatherbenchX import aggregation as ref_agg  # noqa: E402
from weatherbenchx_b200 import aggregation as our_agg  # noqa: E402
from weatherbenchx_b200 import xarray_lite as xl  # noqa: E402


def random_state(rng, lead_labels, with_region):
  """{stat: {var: DataArray}} pairs with the given lead_time labels."""
  sws, sw = {}, {}
  for stat in ('SquaredError', 'Error'):
    sws[stat], sw[stat] = {}, {}
    for var, extra in (('t2m', ()), ('z', ('level',))):
      dims = ('init_time', 'lead_time') + extra + (
          ('region',) if with_region else ())
      coords = {'init_time': np.arange(3), 'lead_time': lead_labels,
                'level': np.array([500, 850]),
                'region': np.array(['global', 'tropics'])}
      shape = tuple(len(coords[d]) for d in dims)
      use = {d: coords[d] for d in dims}
      sws[stat][var] = xl.DataArray(rng.normal(size=shape), dims, coords=use,
                                    name=var)
      sw[stat][var] = xl.DataArray(rng.random(shape) + 0.5, dims, coords=use,
                                   name=var)
  return sws, sw


def both(sws, sw):
  return (ref_agg.AggregationState(sws, sw), our_agg.AggregationState(sws, sw))


def same(a, b, what):
  """Compares two states (or nested trees of DataArrays) entry by entry."""
  if hasattr(a, 'sum_weighted_statistics'):
    if a.sum_weighted_statistics is None or b.sum_weighted_statistics is None:
      assert a.sum_weighted_statistics is None, what
      assert b.sum_weighted_statistics is None, what
      return
    same(a.sum_weighted_statistics, b.sum_weighted_statistics, what + '.sws')
    same(a.sum_weights, b.sum_weights, what + '.sw')
    return
  if isinstance(a, dict):
    assert set(a) == set(b), (what, set(a), set(b))
    for k in a:
      same(a[k], b[k], f'{what}/{k}')
    return
  assert a.dims == b.dims, (what, a.dims, b.dims)
  np.testing.assert_array_equal(a.values, b.values, err_msg=what)
  for d in a.dims:
    if d in a.coords or d in b.coords:
      np.testing.assert_array_equal(a.coords[d].values, b.coords[d].values,
                                    err_msg=f'{what} coord {d}')


def main():
  rng = np.random.default_rng(5)
  leads_a = np.array([0, 6, 12])
  leads_b = np.array([12, 18])            # overlaps a in one label
  for with_region in (False, True):
    ref_a, our_a = both(*random_state(rng, leads_a, with_region))
    ref_b, our_b = both(*random_state(rng, leads_b, with_region))
    ref_c, our_c = both(*random_state(rng, leads_a, with_region))
    # sums: same coordinates, outer join, zero states in any position
    same(ref_a + ref_c, our_a + our_c, 'a+c')
    same(ref_a + ref_b, our_a + our_b, 'a+b (outer join)')
    same(ref_agg.AggregationState.zero() + ref_a,
         our_agg.AggregationState.zero() + our_a, 'zero+a')
    same(ref_a + ref_agg.AggregationState.zero(),
         our_a + our_agg.AggregationState.zero(), 'a+zero')
    same(ref_agg.AggregationState.sum([ref_agg.AggregationState.zero()] * 2),
         our_agg.AggregationState.sum([our_agg.AggregationState.zero()] * 2),
         'zero+zero')
    same(ref_agg.AggregationState.sum([ref_a, ref_b, ref_c]),
         our_agg.AggregationState.sum([our_a, our_b, our_c]), 'sum of three')
    # normalisation, further reduction, dot with resampling weights, map
    same(ref_a.mean_statistics(), our_a.mean_statistics(), 'mean_statistics')
    same((ref_a + ref_b).mean_statistics(), (our_a + our_b).mean_statistics(),
         'mean_statistics after outer join')
    same(ref_a.sum_along_dims(['init_time']),
         our_a.sum_along_dims(['init_time']), 'sum_along_dims')
    same(ref_agg.AggregationState.zero().sum_along_dims(['init_time']),
         our_agg.AggregationState.zero().sum_along_dims(['init_time']),
         'sum_along_dims of zero')
    weights = xl.DataArray(rng.integers(0, 3, (4, 3)).astype(float),
                           ('replicate', 'init_time'),
                           coords={'init_time': np.arange(3)})
    same(ref_a.dot(weights, dim='init_time'),
         our_a.dot(weights, dim='init_time'), 'dot')
    same(ref_a.map(lambda x: x * 2.0), our_a.map(lambda x: x * 2.0), 'map')
    same(ref_agg.AggregationState.map_multi(lambda x, y: x - y, ref_a, ref_c),
         our_agg.AggregationState.map_multi(lambda x, y: x - y, our_a, our_c),
         'map_multi')
    for cls, state in ((ref_agg.AggregationState, ref_a),
                       (our_agg.AggregationState, our_a)):
      try:
        cls.map_multi(lambda x, y: x, state, cls.zero())
      except ValueError as e:
        assert 'zero AggregationState' in str(e)
      else:
        raise AssertionError('mapping a zero state must raise')
  print('state algebra ok')


if __name__ == '__main__':
  main()
