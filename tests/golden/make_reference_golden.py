"""Generates tests/golden/reference_golden.npz by RUNNING THE REFERENCE.

    python tests/golden/make_reference_golden.py

Executes the unmodified code of /root/reference/weatherbenchX (statistics,
metrics incl. categorical, Aggregator, AggregationState, weighting, binning,
wrappers) on small
seeded inputs, through the stand-in modules of ``reference_runtime.py``
(xarray / jax / absl are not installable in this container; read that file's
docstring for exactly what is the reference's and what is the stand-in's).
The cases are defined once in ``reference_cases.py`` against a namespace of
modules; here that namespace is the reference's.  The inputs are stored next to
the outputs, so the consumers (tests/test_reference_golden.py: the NumPy oracle
on CPU, the CUDA path under ``-m gpu``) need neither /root/reference nor the
stand-ins.

Layout of the .npz: ``in/<name>`` input arrays, ``<case>/sws/<stat>/<var>``
and ``<case>/sw/<stat>/<var>`` the two halves of the AggregationState
(aggregation.py:63-83), ``<case>/value/<metric>.<var>`` the metric values
(aggregation.py:122-148) and ``<...>@dims`` the dim names of each output.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_cases  # noqa: E402  pylint: disable=g-import-not-at-top
import reference_runtime  # noqa: E402  pylint: disable=g-import-not-at-top

xr = reference_runtime.install()
from weatherbenchX import aggregation  # noqa: E402
from weatherbenchX import binning  # noqa: E402
from weatherbenchX import weighting  # noqa: E402
from weatherbenchX.metrics import base as metrics_base  # noqa: E402
from weatherbenchX.metrics import categorical  # noqa: E402
from weatherbenchX.metrics import deterministic  # noqa: E402
from weatherbenchX.metrics import probabilistic  # noqa: E402
from weatherbenchX.metrics import wrappers  # noqa: E402

OUT = os.path.join(HERE, 'reference_golden.npz')
STORE: dict = {}

NS = reference_cases.namespace(
    xr=xr, aggregation=aggregation, binning=binning, weighting=weighting,
    base=metrics_base, deterministic=deterministic,
    probabilistic=probabilistic, wrappers=wrappers, categorical=categorical)


def put(key, array):
  STORE[key] = np.asarray(array)


def put_labelled(key, da):
  STORE[key] = np.asarray(da.values)
  STORE[key + '@dims'] = np.array(list(da.dims), dtype='U32')
  for d in da.dims:
    if d in da.coords and da.coords[d].values.dtype.kind in 'USO':
      STORE[f'{key}@labels/{d}'] = np.array(
          [str(v) for v in da.coords[d].values], dtype='U64')


def record(case, metrics, aggregator, predictions, targets):
  """Runs the reference end to end and stores state + values."""
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  state = aggregator.aggregate_statistics(statistics)
  for stat_name, per_var in state.sum_weighted_statistics.items():
    for var, da in per_var.items():
      put_labelled(f'{case}/sws/{stat_name}/{var}', da)
      put_labelled(f'{case}/sw/{stat_name}/{var}',
                   state.sum_weights[stat_name][var])
  values = state.metric_values(metrics)
  for name, da in values.items():
    put_labelled(f'{case}/value/{name}', da)
  return state


def weighting_cases():
  """GridAreaWeighting on several latitude grids (weighting.py:45-130)."""
  grids = {
      'poles_ascending': np.linspace(-90, 90, 19),
      'poles_descending': np.linspace(90, -90, 33),
      'no_poles': np.linspace(-87.1875, 87.1875, 32),
      'quarter_degree': np.linspace(-90, 90, 721),
      'float32_coord': np.linspace(-90, 90, 37).astype(np.float32),
  }
  for name, lat in grids.items():
    stat = xr.DataArray(np.zeros((len(lat), 4), np.float32),
                        ('latitude', 'longitude'),
                        coords={'latitude': lat,
                                'longitude': np.arange(4) * 90.0})
    w = weighting.GridAreaWeighting().weights(stat)
    put(f'weights/{name}/latitude', lat)
    put(f'weights/{name}/weights', w.values)
  scalar = weighting.GridAreaWeighting().weights(
      xr.DataArray(np.zeros(3, np.float32), ('init_time',)))
  put('weights/no_latitude_dim', np.asarray(scalar))


def main():
  inputs = reference_cases.make_inputs()
  for name, values in inputs.items():
    put(f'in/{name}', values)
  names = []
  for name, _, metrics, aggregator, predictions, targets in (
      reference_cases.build_cases(NS, inputs)):
    record(name, metrics, aggregator, predictions, targets)
    names.append(name)
  metrics, total = reference_cases.chunked_case(NS, inputs)
  for name, da in total.metric_values(metrics).items():
    put_labelled(f'det/chunked/value/{name}', da)
  weighting_cases()
  # bin masks of the coordinate-value binnings (binning.py:301-704)
  mask_names = []
  for name, instance, stat in reference_cases.mask_cases(NS):
    mask = instance.create_bin_mask(stat)
    bdim = instance.bin_dim_name
    order = [bdim] + [d for d in mask.dims if d != bdim]
    put(f'masks/{name}/mask', mask.transpose(*order).values)
    put(f'masks/{name}/dims', np.array(order, dtype='U32'))
    put(f'masks/{name}/labels',
        np.array([str(v) for v in mask.coords[bdim].values], dtype='U32'))
    put(f'masks/{name}/label_kind', np.array(
        mask.coords[bdim].values.dtype.kind))
    mask_names.append(name)
  put('mask_cases', np.array(mask_names, dtype='U64'))
  put('cases', np.array(names, dtype='U64'))
  np.savez_compressed(OUT, **STORE)
  size = os.path.getsize(OUT)
  print(f'wrote {OUT}: {len(names)} cases, {len(STORE)} arrays, '
        f'{size / 1024:.0f} KiB')


if __name__ == '__main__':
  main()
