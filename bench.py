#!/usr/bin/env python
"""bench.py -- throughput of the WeatherBench-X statistic + aggregation hot path.

Headline workload (BASELINE.json metric: grid-points/sec for lat-weighted RMSE
on 0.25 degree fields): one step = one pass of lat-weighted RMSE
(SquaredError statistic x GridAreaWeighting, reduced over init_time, latitude,
longitude) over a batch of 5 variables x 20 init_times x 721 x 1440 float32
prediction/target fields = 103.8 M grid points, 830 MB read (8 algorithmic
bytes per point).  Inputs are synthetic (Gaussian, ERA5-like magnitudes).

  value  : device-resident inputs, K timed steps, CUDA events, max over ranks.
  e2e    : the same step through the public class API
           (aggregation.compute_metric_values_for_single_chunk) with HOST
           (pinned) numpy inputs; host->device streaming and the device->host
           read of the result are inside the timed region.
  --impl reference : the reference's CPU path (NumPy restatement that mirrors
           it op for op: oracle/wbx_oracle.py::reference_path_rmse) on all host
           cores, same workload.

N > 1 (torchrun): weak scaling -- every rank owns its own batches of the same
shape (shards of (variable, init_time)); every step adds its chunk into the
rank's device-resident AggregationState, and after the K steps ONE all-reduce of
the packed float64 state over NCCL combines the ranks (the reference's
CombinePerKey over all chunks) -- inside the timed region.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

N_VARS, N_INIT, NLAT, NLON = 5, 20, 721, 1440
VAR_NAMES = ['2m_temperature', '10m_u_component_of_wind',
             '10m_v_component_of_wind', 'mean_sea_level_pressure',
             'total_precipitation_6hr']
POINTS_PER_STEP = N_VARS * N_INIT * NLAT * NLON
ALG_BYTES_PER_POINT = 8  # read prediction + target once
METRIC = 'grid-points/sec, lat-weighted RMSE on 0.25deg (721x1440) fields'
WORKLOAD = (f'lat-weighted RMSE (SquaredError x GridAreaWeighting, reduce '
            f'init_time/latitude/longitude), {N_VARS} vars x {N_INIT} init x '
            f'{NLAT}x{NLON} f32')


def make_config(world: int) -> dict:
  """The workload description; both arms (--impl b200 / reference) emit this
  very dict so that the driver can match them."""
  return {
      'workload': WORKLOAD,
      'points_per_step_per_gpu': POINTS_PER_STEP,
      'bytes_per_step_per_gpu': POINTS_PER_STEP * ALG_BYTES_PER_POINT,
      'l2': 'inputs (830 MB per step) exceed the 126 MB L2; no flush needed',
      'parallelism': (f'dp{world}: every rank aggregates its own (variable, '
                      'init_time) chunks; ONE all-reduce of the packed float64 '
                      'AggregationState combines the ranks')
                     if world > 1 else 'single GPU',
  }


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
  """nvidia-smi clock / throttle sampling during the timed region."""

  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,'
           'clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, device_index: int):
    self.device_index = device_index
    self.proc = None
    self.lines = []
    self.thread = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--query-gpu={self.QUERY}',
           '--format=csv,noheader,nounits', '-lms', '100',
           '-i', str(self.device_index)],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
      self.proc = None
      return

    def pump():
      for line in self.proc.stdout:
        self.lines.append((time.time(), line.strip()))

    self.thread = threading.Thread(target=pump, daemon=True)
    self.thread.start()

  def stop(self, windows):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    for ts, line in self.lines:
      if not any(a <= ts <= b for a, b in windows):
        continue
      parts = [x.strip() for x in line.split(',')]
      if len(parts) < 9:
        continue
      try:
        sm.append(float(parts[1]))
        smax.append(float(parts[2]))
      except ValueError:
        continue
      for name, val in zip(names, parts[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    return {
        'sm_mhz': float(np.median(sm)) if sm else None,
        'sm_max_mhz': float(max(smax)) if smax else None,
        'reasons': sorted(reasons), 'samples': len(sm),
    }


# ---------------------------------------------------------------------------
# reference arm (CPU)
# ---------------------------------------------------------------------------

_W = {}


def _ref_worker_init(seed, n_init):
  import wbx_oracle as oracle
  rng = np.random.default_rng(seed)
  _W['p'] = rng.standard_normal((n_init, NLAT, NLON), dtype=np.float32)
  _W['t'] = rng.standard_normal((n_init, NLAT, NLON), dtype=np.float32)
  _W['w'] = oracle.grid_area_weights(np.linspace(-90, 90, NLAT))
  _W['oracle'] = oracle


def _ref_worker_step(n):
  return _W['oracle'].reference_path_rmse(_W['p'][:n], _W['t'][:n], _W['w'])


def run_reference(args):
  """The reference's CPU path on all host cores (kind: port, see module doc)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import multiprocessing as mp
  cores = os.cpu_count() or 1
  total_units = N_VARS * N_INIT          # (variable, init_time) fields per step
  workers = min(cores, total_units)
  # every worker owns a contiguous block of (variable, init_time) fields; the
  # parent sums the per-worker states (the CombinePerKey of the reference).
  per = [total_units // workers + (1 if i < total_units % workers else 0)
         for i in range(workers)]
  ctx = mp.get_context('fork')
  pools = []
  for i, n in enumerate(per):
    pool = ctx.Pool(1, initializer=_ref_worker_init, initargs=(1000 + i, n))
    pools.append(pool)
  counts = list(per)

  def step():
    results = [p.apply_async(_ref_worker_step, (n,))
               for p, n in zip(pools, counts)]
    sws = sum(r.get()[0] for r in results)
    sw = sum(r.get()[1] for r in results)
    return float(np.sqrt(sws / sw))
  step()
  t0 = time.perf_counter()
  step()
  t_full = time.perf_counter() - t0
  # Bound the whole run to ~150 s: if K full passes would take longer, every
  # step processes a fixed fraction of each worker's fields instead.
  budget = 150.0
  frac = min(1.0, budget / max(t_full * (args.steps + args.warmup), 1e-9))
  counts = [max(1, int(n * frac)) for n in per]
  points = sum(counts) * NLAT * NLON
  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step()
  dt = time.perf_counter() - t0
  for p in pools:
    p.terminate()
  value = points * args.steps / dt
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value,
      'unit': 'grid-points/s', 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': make_config(args.gpus),
      'cpu_baseline': {
          'value': value, 'unit': 'grid-points/s', 'cores': workers,
          'kind': 'port',
          'sample': (f'{points} of {POINTS_PER_STEP} grid points per step '
                     f'({sum(counts)} of {total_units} fields); '
                     'NumPy restatement mirroring the '
                     'reference op for op (unfused temporaries, ones_like, '
                     'two einsums); xarray label overhead not included '
                     '(xarray not installable)')},
      'e2e': {'value': value, 'unit': 'grid-points/s',
              'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  emit_line(line)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------


def cpu_baseline_sample():
  """Single-process NumPy reference path on a bounded sample (rank 0, N=1)."""
  import wbx_oracle as oracle
  n_init = 4
  rng = np.random.default_rng(0)
  p = rng.standard_normal((n_init, NLAT, NLON), dtype=np.float32)
  t = rng.standard_normal((n_init, NLAT, NLON), dtype=np.float32)
  w = oracle.grid_area_weights(np.linspace(-90, 90, NLAT))
  oracle.reference_path_rmse(p, t, w)
  reps, t0 = 0, time.perf_counter()
  while time.perf_counter() - t0 < 8.0:
    oracle.reference_path_rmse(p, t, w)
    reps += 1
  dt = time.perf_counter() - t0
  return {
      'value': reps * n_init * NLAT * NLON / dt, 'unit': 'grid-points/s',
      'cores': 1, 'kind': 'port',
      'sample': (f'{reps} passes over {n_init} x {NLAT} x {NLON} f32 fields '
                 f'({dt:.1f} s), single NumPy process, reference-mirroring '
                 'path (oracle.reference_path_rmse)'),
  }


def run_suite(ctx, dev, peak, oos_legs=False):
  """Secondary workloads of BASELINE.json (single GPU, device resident):
  config[1] RMSE+ACC on 1.4 deg pressure-level fields and config[2] CRPS with
  a 50-member ensemble on 0.25 deg fields.  Reported next to the headline; each
  entry carries its own roofline fraction (kernel time from CUDA events)."""
  import torch
  from weatherbenchx_b200 import aggregation, weighting
  from weatherbenchx_b200 import xarray_lite as xl
  from weatherbenchx_b200.metrics import deterministic, probabilistic
  out = {}

  def consume(result):
    # read every value of a metric Dataset (forces the decode of a replayed
    # chunk); device arrays (spectra) stay where they are
    if isinstance(result, dict):
      for name in result:
        result[name].values   # pylint: disable=expression-not-assigned

  def timed(fn, steps):
    """(ms per step, kernel ms per step, launches per step).  The step time
    is taken WITHOUT per-kernel event bracketing and with the API used the way
    a chunk loop uses it: the values of call i - 1 are read after call i has
    been issued, all of them inside the timed region.  The kernel time comes
    from a second pass with every main kernel bracketed by CUDA events."""
    consume(fn())
    consume(fn())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    previous = None
    for _ in range(steps):
      current = fn()
      consume(previous)
      previous = current
    consume(previous)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.profile(True)
    ctx.kernel_time(reset=True)
    for _ in range(steps):
      consume(fn())
    torch.cuda.synchronize()
    kms, kn = ctx.kernel_time(reset=True)
    ctx.profile(False)
    return ms, kms / steps, kn // steps

  # ---- config[1]: RMSE + ACC, 6 vars x [40 init, 10 lead, 13 levels, 128, 256]
  n_var, n_init, n_lead, n_lev, ny, nx = 6, 40, 10, 13, 128, 256
  lat = np.linspace(-90, 90, ny)
  coords = {
      'init_time': np.datetime64('2020-01-01T00', 'ns') +
                   np.arange(n_init) * np.timedelta64(12, 'h'),
      'lead_time': (np.arange(n_lead) * np.timedelta64(6, 'h')
                    ).astype('timedelta64[ns]'),
      'level': np.arange(n_lev), 'latitude': lat,
      'longitude': np.linspace(0, 360, nx, endpoint=False)}
  dims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  cdims = ('dayofyear', 'hour', 'level', 'latitude', 'longitude')
  ccoords = {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
             'level': coords['level'], 'latitude': lat,
             'longitude': coords['longitude']}
  gen = torch.Generator(device=dev)
  gen.manual_seed(2000)
  preds, tgts, clim = {}, {}, {}
  for v in range(n_var):
    name = f'var{v}'
    t = torch.empty((n_init, n_lead, n_lev, ny, nx), device=dev)
    t.normal_(0.0, 1.0, generator=gen)
    p = t + 0.3 * torch.empty_like(t).normal_(0.0, 1.0, generator=gen)
    c = torch.empty((366, 4, n_lev, ny, nx), device=dev)
    c.normal_(0.0, 0.5, generator=gen)
    preds[name] = xl.DataArray(p, dims, coords=coords, name=name)
    tgts[name] = xl.DataArray(t, dims, coords=coords, name=name)
    clim[name] = xl.DataArray(c, cdims, coords=ccoords, name=name)
  metrics = {'rmse': deterministic.RMSE(), 'acc': deterministic.ACC(clim)}
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, aggregator, preds, tgts)
  ms, kms, kn = timed(step, 5)
  # the numbers being timed are the right numbers: RMSE and ACC of one variable
  # recomputed with torch in float64 (climatology gathered at the valid times)
  got = step()
  w64 = torch.as_tensor(weighting.GridAreaWeighting().weights(
      tgts['var0']).values, device=dev)[None, None, None, :, None]
  valid = coords['init_time'][:, None] + coords['lead_time'][None, :]
  doy = ((valid.astype('datetime64[D]') - valid.astype('datetime64[Y]').astype(
      'datetime64[D]')).astype(np.int64))
  hour = ((valid - valid.astype('datetime64[D]')) // np.timedelta64(6, 'h')
          ).astype(np.int64)
  p64, t64 = preds['var0'].data.double(), tgts['var0'].data.double()
  c64 = clim['var0'].data[torch.as_tensor(doy, device=dev),
                          torch.as_tensor(hour, device=dev)].double()
  wsum = (w64 * torch.ones_like(p64)).sum(dim=(0, 3, 4))
  mean = lambda x: (x * w64).sum(dim=(0, 3, 4)) / wsum  # noqa: E731
  ref_rmse = torch.sqrt(mean((p64 - t64) ** 2)).cpu().numpy()
  ref_acc = (mean((p64 - c64) * (t64 - c64)) / torch.sqrt(
      mean((p64 - c64) ** 2) * mean((t64 - c64) ** 2))).cpu().numpy()
  np.testing.assert_allclose(
      got['rmse.var0'].transpose('lead_time', 'level').values, ref_rmse,
      rtol=1e-5)
  np.testing.assert_allclose(
      got['acc.var0'].transpose('lead_time', 'level').values, ref_acc,
      rtol=1e-5)
  del p64, t64, c64, got, wsum
  pts = n_var * n_init * n_lead * n_lev * ny * nx
  out['rmse_acc_c2'] = {
      'workload': 'RMSE+ACC fused (4 statistics, one pass), 6 vars x '
                  '[40 init,10 lead,13 level,128,256] f32 + climatology '
                  '[366,4,13,128,256], class API, device inputs',
      'checked': 'rmse.var0 and acc.var0 == torch float64 recomputation '
                 '(rtol 1e-5)',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * 12 / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * 12 / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': 12,
                   'note': 'SURVEY 8(d) counts the climatology row once per '
                           '(init, lead) use (12 B per point); rows of the '
                           'same (dayofyear, hour) are shared by several init '
                           'times and served from L2, so the DRAM traffic is '
                           'below the algorithmic bytes and the fraction can '
                           'exceed 1'}}
  del preds, tgts, clim, metrics, step
  torch.cuda.empty_cache()

  # ---- binned RMSE: 8 regions x {all, land} = 16 bins, 0.25 deg
  from weatherbenchx_b200 import binning
  lat025 = np.linspace(-90, 90, NLAT)
  lon025 = np.linspace(0, 360, NLON, endpoint=False)
  rng = np.random.default_rng(7)
  land = xl.DataArray(
      # continent-sized land blocks (8.75 x 15 degrees), like real coastlines
      np.kron(rng.random((21, 24)) > 0.7, np.ones((35, 60), bool))[:NLAT],
      ('latitude', 'longitude'), coords={'latitude': lat025, 'longitude': lon025})
  # the public benchmark's regions (run_benchmark_evaluation.py:110-132): with
  # the land-sea mask 17 regions x {all, land} = 34 bins (:369)
  regions = {
      'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
      'northern-hemisphere': ((20, 90), (0, 360)),
      'southern-hemisphere': ((-90, -20), (0, 360)),
      'europe': ((35, 75), (-12.5, 42.5)),
      'north-america': ((25, 60), (360 - 120, 360 - 75)),
      'north-atlantic': ((25, 65), (360 - 70, 360 - 10)),
      'north-pacific': ((25, 60), (145, 360 - 130)),
      'east-asia': ((25, 60), (102.5, 150)),
      'ausnz': ((-45, -12.5), (120, 175)),
      'arctic': ((60, 90), (0, 360)), 'antarctic': ((-90, -60), (0, 360)),
      'northern-africa': ((5, 32.5), (-12.5, 37.5)),
      'southern-africa': ((-30, 5), (12.5, 37.5)),
      'south-america': ((-40, 5), (-75, -45)),
      'west-asia': ((15, 60), (42.5, 102.5)),
      'south-east-asia': ((-12.5, 25), (95, 125))}
  bcoords = {'init_time': np.arange(20), 'latitude': lat025,
             'longitude': lon025}
  bdims = ('init_time', 'latitude', 'longitude')
  preds, tgts = {}, {}
  for v in range(5):
    name = f'var{v}'
    t = torch.empty((20, NLAT, NLON), device=dev)
    t.normal_(280.0, 10.0, generator=gen)
    p = t + torch.empty_like(t).normal_(0.0, 2.0, generator=gen)
    preds[name] = xl.DataArray(p, bdims, coords=bcoords, name=name)
    tgts[name] = xl.DataArray(t, bdims, coords=bcoords, name=name)
  metrics = {'rmse': deterministic.RMSE()}
  bin_agg = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()],
      bin_by=[binning.Regions(regions, land_sea_mask=land)])
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, bin_agg, preds, tgts)
  ms, kms, kn = timed(step, 10)
  pts = 5 * 20 * NLAT * NLON
  out['rmse_bins34'] = {
      'workload': 'lat-weighted RMSE in the 34 bins of the public benchmark '
                  '(17 regions x {all, land}), 5 vars x 20 init x 721x1440 '
                  'f32, fused binned kernel, class API, device inputs',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * 8 / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * 8 / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': 8,
                   'note': '8 B fields; the class map is compiled into a '
                           'static reduction schedule on the host and never '
                           'read by the GPU (csrc/det_bins3.cuh)'}}
  # ---- the public benchmark's chunk shape: (init=1, lead=12) chunks keep
  # lead_time, so every field is its own output cell (one flush per field),
  # the suite asks for Error, AbsoluteError and SquaredError at once and
  # masked=True counts a mask (run_benchmark_evaluation.py:96-101,301-361,374)
  lead_dims = ('lead_time', 'latitude', 'longitude')
  lcoords = {'lead_time': np.arange(20), 'latitude': lat025,
             'longitude': lon025}
  mask20 = torch.rand((20, NLAT, NLON), device=dev, generator=gen) > 0.02
  lpreds = {n: xl.DataArray(a.data, lead_dims, coords=lcoords, name=n)
            for n, a in preds.items()}
  ltgts = {n: xl.DataArray(a.data, lead_dims, coords=lcoords, name=n
                           ).assign_coords(mask=xl.DataArray(mask20, lead_dims))
           for n, a in tgts.items()}
  lmetrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE(),
              'bias': deterministic.Bias()}
  lead_agg = aggregation.Aggregator(
      reduce_dims=['latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()],
      bin_by=[binning.Regions(regions, land_sea_mask=land)], masked=True)
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      lmetrics, lead_agg, lpreds, ltgts)
  ms, kms, kn = timed(step, 10)
  out['suite3_bins34_per_lead_masked'] = {
      'workload': 'RMSE + MAE + Bias (3 statistics, one pass) in the 34 bins, '
                  'masked=True, lead_time kept: 100 fields = 100 output cells '
                  'x 91 classes, 721x1440 f32, class API, device inputs',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * 9 / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * 9 / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': 9,
                   'note': '8 B fields + 1 B mask'}}
  del lpreds, ltgts, mask20, lmetrics
  # ---- the same fields stored longitude-major (latitude is the fastest axis,
  # as in the 1440x721 WeatherBench archives): the latitude weight then varies
  # along the rows of the slab (w_x), 721 is odd so float4 groups straddle rows
  lm_dims = ('init_time', 'longitude', 'latitude')
  lm_preds = {n: xl.DataArray(a.data.transpose(1, 2).contiguous(), lm_dims,
                              coords=bcoords, name=n) for n, a in preds.items()}
  lm_tgts = {n: xl.DataArray(a.data.transpose(1, 2).contiguous(), lm_dims,
                             coords=bcoords, name=n) for n, a in tgts.items()}
  del preds, tgts
  plain_agg = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, plain_agg, lm_preds, lm_tgts)
  ms, kms, kn = timed(step, 10)
  out['rmse_lon_major'] = {
      'workload': 'lat-weighted RMSE, 5 vars x 20 init x 1440x721 f32 stored '
                  '[init, longitude, latitude] (latitude fastest): per-column '
                  'weights, class API, device inputs',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * 8 / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * 8 / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': 8}}
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, bin_agg, lm_preds, lm_tgts)
  ms, kms, kn = timed(step, 10)
  out['rmse_bins34_lon_major'] = {
      'workload': 'rmse_bins34 on the longitude-major arrays of '
                  'rmse_lon_major (binned kernel with element weights)',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * 8 / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * 8 / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': 8}}
  del lm_preds, lm_tgts, metrics, step
  torch.cuda.empty_cache()

  # ---- config[2]: CRPS, M = 50, 5 vars x 20 init x 721 x 1440
  n_var, n_init, m = 5, 20, 50
  lat = np.linspace(-90, 90, NLAT)
  ecoords = {'init_time': np.arange(n_init), 'number': np.arange(m),
             'latitude': lat,
             'longitude': np.linspace(0, 360, NLON, endpoint=False)}
  preds, tgts = {}, {}
  for v in range(n_var):
    name = f'var{v}'
    y = torch.empty((n_init, NLAT, NLON), device=dev)
    y.normal_(0.0, 1.0, generator=gen)
    x = torch.empty((n_init, m, NLAT, NLON), device=dev)
    x.normal_(0.0, 1.0, generator=gen)
    x += y[:, None]
    preds[name] = xl.DataArray(
        x, ('init_time', 'number', 'latitude', 'longitude'), coords=ecoords,
        name=name)
    tgts[name] = xl.DataArray(
        y, ('init_time', 'latitude', 'longitude'),
        coords={k: ecoords[k] for k in ('init_time', 'latitude', 'longitude')},
        name=name)
  from weatherbenchx_b200 import engine
  metrics = {'crps': probabilistic.CRPSEnsemble()}
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, aggregator, preds, tgts)
  engine.CRPS_KERNEL = 'pair'     # this leg measures the O(M^2) pair kernel
  ms, kms, kn = timed(step, 3)
  engine.CRPS_KERNEL = 'auto'
  pts = n_var * n_init * NLAT * NLON
  bpp = 4 * (m + 1)
  flops = 2.0 * (m * (m - 1) / 2) * 2 + 2 * m   # sub+abs-add per pair, skill
  out['crps_c3'] = {
      'workload': 'CRPSEnsemble fair, pairwise O(M^2) kernel forced '
                  '(engine.CRPS_KERNEL="pair"; the default routes M <= 64 to '
                  'the sorting network, see crps_c3_sort), M=50, 5 vars x 20 '
                  'init x 721x1440 f32, ensemble layout [init, member, lat, '
                  'lon], class API, device inputs',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * bpp / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * bpp / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': bpp,
                   'fp32_tflops': pts * flops / (kms * 1e-3) / 1e12,
                   'note': 'at the FP32-issue / HBM ridge: ~2.5 kFLOP per '
                           '204 B point (SURVEY.md 8d)'}}
  # the same call under the default kernel policy: what a user of the
  # reference's default CRPSEnsemble() (use_sort=False) gets -- both
  # estimators are one statistic (probabilistic.py:190-192), so member-major
  # ensembles of up to 64 members run the sorting network either way
  ms, kms, kn = timed(step, 3)
  out['crps_c3_default_policy'] = {
      'workload': 'CRPSEnsemble() exactly as crps_c3 but with the default '
                  'engine.CRPS_KERNEL="auto" (sorting network for M <= 64)',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * bpp / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * bpp / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': bpp}}
  metrics = {'crps': probabilistic.CRPSEnsemble(use_sort=True)}
  step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, aggregator, preds, tgts)
  ms, kms, kn = timed(step, 3)
  # the number being timed is the right number: fair CRPS of one variable
  # recomputed with torch in float64 (sorted-member form), init by init
  got = float(step()['crps.var0'].values)
  w_lat = torch.as_tensor(weighting.GridAreaWeighting().weights(
      tgts['var0']).values, device=dev)[:, None]
  coef = (2.0 * torch.arange(m, device=dev, dtype=torch.float64) - (m - 1)
          )[:, None, None]
  num = 0.0
  for i in range(n_init):
    xs = preds['var0'].data[i].double()
    skill = (xs - tgts['var0'].data[i].double()[None]).abs().mean(dim=0)
    xs = torch.sort(xs, dim=0).values
    spread = 2.0 * (coef * xs).sum(dim=0) / (m * (m - 1))
    num += float(((skill - 0.5 * spread) * w_lat).sum())
    del xs, skill, spread
  ref = num / (n_init * float(w_lat.sum()) * NLON)
  assert abs(got - ref) <= 1e-5 * abs(ref), (got, ref)
  out['crps_c3_sort'] = {
      'workload': 'CRPSEnsemble fair, use_sort=True (sort/PWM estimator in a '
                  'register sorting network), same data as crps_c3',
      'checked': 'crps.var0 == torch float64 recomputation (rtol 1e-5)',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * bpp / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * bpp / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': bpp}}
  for key, metrics, what in (
      ('ens_moments_c3',
       {'ssr': probabilistic.UnbiasedSpreadSkillRatio(),
        'rmse': probabilistic.UnbiasedEnsembleMeanRMSE()},
       'UnbiasedSpreadSkillRatio + UnbiasedEnsembleMeanRMSE (ensemble '
       'variance and unbiased ensemble-mean MSE; moments-only launch)'),
      ('crps_ssr_c3',
       {'crps': probabilistic.CRPSEnsemble(use_sort=True),
        'ssr': probabilistic.UnbiasedSpreadSkillRatio()},
       'CRPSEnsemble(use_sort=True) + UnbiasedSpreadSkillRatio, all four '
       'ensemble statistics from one read of the ensemble')):
    step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
        metrics, aggregator, preds, tgts)
    ms, kms, kn = timed(step, 3)
    out[key] = {
        'workload': what + ', same data as crps_c3',
        'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
        'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
        'roofline': {'bound': 'hbm',
                     'achieved': pts * bpp / (kms * 1e-3) / 1e9,
                     'peak': peak, 'unit': 'GB/s',
                     'frac': pts * bpp / (kms * 1e-3) / 1e9 / peak,
                     'algorithmic_bytes_per_point': bpp}}
  del preds, tgts, metrics, step
  torch.cuda.empty_cache()

  # ---- config[3]: zonal energy spectrum, 13 levels x 6 vars x 721 x 1440
  from weatherbenchx_b200.metrics import spectral
  n_fields = 13 * 6
  f = torch.empty((n_fields, NLAT, NLON), device=dev)
  f.normal_(0.0, 1.0, generator=gen)
  field = xl.DataArray(
      f, ('field', 'latitude', 'longitude'),
      coords={'latitude': lat,
              'longitude': np.linspace(0, 360, NLON, endpoint=False)},
      name='u')
  step = lambda: spectral.zonal_energy_spectrum(field)  # noqa: E731
  ms, kms, kn = timed(step, 10)
  pts = n_fields * NLAT * NLON
  bpp = 4.0 + 4.0 * (NLON // 2 + 1) / NLON
  out['spectrum_c4'] = {
      'workload': 'ZonalEnergySpectrum (rfft N=1440 per latitude row, '
                  'per-row spectra written), 13 levels x 6 vars x 721x1440 '
                  'f32; parity unpinned (numpy.fft oracle only)',
      'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
      'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
      'roofline': {'bound': 'hbm', 'achieved': pts * bpp / (kms * 1e-3) / 1e9,
                   'peak': peak, 'unit': 'GB/s',
                   'frac': pts * bpp / (kms * 1e-3) / 1e9 / peak,
                   'algorithmic_bytes_per_point': bpp}}
  del f, field, step
  torch.cuda.empty_cache()

  if oos_legs:
    run_oos_legs(ctx, dev, peak, gen, lat, timed, out)
  return out


def run_oos_legs(ctx, dev, peak, gen, lat, timed, out):
  """Legs of components SURVEY.md marks out of scope (categorical statistics,
  SEEPS); kept runnable (`--oos-legs`) but not part of the default line."""
  import torch
  from weatherbenchx_b200 import aggregation, weighting
  from weatherbenchx_b200 import xarray_lite as xl
  # ---- thresholded contingency table (CSI / ETS / ...; categorical.py,
  # wrappers.ContinuousToBinary): 3 thresholds, 0.25 deg.  Added after the last
  # GPU session of round 1, so a failure here must not take the line down.
  try:
    from weatherbenchx_b200.metrics import categorical, wrappers
    n_var, n_init, thresholds = 2, 20, [0.1, 1.0, 5.0]
    dims = ('init_time', 'latitude', 'longitude')
    coords = {'init_time': np.arange(n_init), 'latitude': lat,
              'longitude': np.linspace(0, 360, NLON, endpoint=False)}
    preds, tgts = {}, {}
    for v in range(n_var):
      t = torch.empty((n_init, NLAT, NLON), device=dev)
      t.exponential_(0.5, generator=gen)
      p = (t + torch.empty_like(t).normal_(0.0, 1.0, generator=gen)).clamp_(0)
      preds[f'rain{v}'] = xl.DataArray(p, dims, coords=coords, name=f'rain{v}')
      tgts[f'rain{v}'] = xl.DataArray(t, dims, coords=coords, name=f'rain{v}')
    both = [wrappers.ContinuousToBinary('both', thresholds, 'threshold')]
    metrics = {'csi': wrappers.WrappedMetric(categorical.CSI(), both),
               'ets': wrappers.WrappedMetric(categorical.ETS(), both)}
    aggregator = aggregation.Aggregator(
        reduce_dims=['init_time', 'latitude', 'longitude'],
        weigh_by=[weighting.GridAreaWeighting()])
    step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
        metrics, aggregator, preds, tgts)
    ms, kms, kn = timed(step, 5)
    pts = n_var * n_init * NLAT * NLON
    bpp = 8.0 * len(thresholds)
    out['contingency_3thr'] = {
        'workload': 'CSI + ETS (whole 2x2 table, thresholds applied on load), '
                    '3 thresholds, 2 vars x 20 init x 721x1440 f32, class API, '
                    'device inputs',
        'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
        'kernel_ms_per_step': kms, 'launches_per_step': int(kn),
        'roofline': {'bound': 'hbm',
                     'achieved': pts * bpp / (kms * 1e-3) / 1e9, 'peak': peak,
                     'unit': 'GB/s',
                     'frac': pts * bpp / (kms * 1e-3) / 1e9 / peak,
                     'algorithmic_bytes_per_point': bpp,
                     'note': '8 B per point and threshold: every threshold is '
                             'a pass over the two fields'}}
    del preds, tgts, metrics, step
  except Exception as e:  # pylint: disable=broad-except
    out['contingency_3thr'] = {'error': f'{type(e).__name__}: {e}'}
    try:
      ctx.profile(False)
    except Exception:  # pylint: disable=broad-except
      pass
  torch.cuda.empty_cache()

  # ---- SEEPS (categorical.SEEPS: elementwise kernel + fused masked reduction),
  # 0.25 deg.  Also added late in round 1: guarded like the leg above.
  try:
    from weatherbenchx_b200.metrics import categorical
    n_init = 20
    dims = ('init_time', 'lead_time', 'latitude', 'longitude')
    lon = np.linspace(0, 360, NLON, endpoint=False)
    coords = {
        'init_time': np.datetime64('2020-01-01T00', 'ns') +
                     np.arange(n_init) * np.timedelta64(6, 'h'),
        'lead_time': np.zeros(1, 'timedelta64[ns]'),
        'latitude': lat, 'longitude': lon}
    t = torch.empty((n_init, 1, NLAT, NLON), device=dev)
    t.exponential_(500.0, generator=gen)            # metres, mean 2 mm
    p = (t + torch.empty_like(t).normal_(0.0, 2e-3, generator=gen)).clamp_(0)
    preds = {'rain': xl.DataArray(p, dims, coords=coords, name='rain')}
    tgts = {'rain': xl.DataArray(t, dims, coords=coords, name='rain')}
    cdims = ('dayofyear', 'hour', 'latitude', 'longitude')
    ccoords = {'dayofyear': np.arange(1, 9), 'hour': np.arange(0, 24, 6),
               'latitude': lat, 'longitude': lon}
    wet = torch.empty((8, 4, NLAT, NLON), device=dev)
    wet.uniform_(1e-3, 8e-3, generator=gen)
    rng = np.random.default_rng(11)
    dry = np.broadcast_to(
        rng.uniform(0.0, 1.0, (NLAT, NLON)).astype(np.float32),
        (8, 4, NLAT, NLON))
    clim = xl.Dataset({
        'rain_seeps_threshold': xl.DataArray(wet, cdims, coords=ccoords),
        'rain_seeps_dry_fraction': xl.DataArray(dry, cdims, coords=ccoords)})
    metrics = {'seeps': categorical.SEEPS(['rain'], clim)}
    aggregator = aggregation.Aggregator(
        reduce_dims=['init_time', 'latitude', 'longitude'],
        weigh_by=[weighting.GridAreaWeighting()], masked=True)
    step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
        metrics, aggregator, preds, tgts)
    ms, kms, kn = timed(step, 5)
    pts = n_init * NLAT * NLON
    bpp = 21.0
    out['seeps'] = {
        'workload': 'SEEPS, 1 var x 20 init x 721x1440 f32, climatology on '
                    'the device, masked lat-weighted mean, class API, device '
                    'inputs',
        'value': pts / (ms * 1e-3), 'unit': 'grid-points/s', 'ms_per_step': ms,
        'reduction_kernel_ms_per_step': kms, 'launches_per_step': int(kn),
        'roofline': {'bound': 'hbm',
                     'achieved': pts * bpp / (ms * 1e-3) / 1e9, 'peak': peak,
                     'unit': 'GB/s', 'frac': pts * bpp / (ms * 1e-3) / 1e9 / peak,
                     'algorithmic_bytes_per_point': bpp,
                     'note': 'WHOLE STEP (threshold gather, elementwise SEEPS '
                             'kernel 12 B read + 4 B written, masked reduction '
                             '4 B + 1 B, host planning), CUDA events around '
                             'the class-API call; the elementwise kernel is '
                             'not timed on its own yet'}}
    del preds, tgts, clim, metrics, step, wet, p, t
  except Exception as e:  # pylint: disable=broad-except
    out['seeps'] = {'error': f'{type(e).__name__}: {e}'}
    try:
      ctx.profile(False)
    except Exception:  # pylint: disable=broad-except
      pass
  torch.cuda.empty_cache()
  return out


class _RecyclingForecasts:
  """Synthetic forecast archive for the C5 leg: init time i serves the host
  (pinned) block i % n_blocks, so that an archive of any length needs a few
  hundred MB of host memory.  Same ``load_chunk`` contract as
  data_loaders.array_loaders.PredictionsFromArrays (exact init / lead
  selection); the chunk is a zero-copy view of the pinned block."""

  def __init__(self, blocks, init_times, lead_times, grid, extra_dims=()):
    self._blocks = blocks               # {var: ndarray [n_blocks, lead, ...]}
    self._init = np.asarray(init_times)
    self._lead = np.asarray(lead_times)
    self._grid = grid
    self._extra = tuple(extra_dims)     # dims between lead_time and the grid

  def load_chunk(self, init_times, lead_times=None, reference=None):
    from weatherbenchx_b200 import xarray_lite as xl
    del reference
    out = {}
    pos = np.searchsorted(self._init, np.asarray(init_times))
    lead_pos = np.searchsorted(self._lead, np.asarray(lead_times))
    lo, hi = int(lead_pos[0]), int(lead_pos[-1]) + 1
    assert len(pos) == 1 and np.array_equal(lead_pos, np.arange(lo, hi))
    for var, blocks in self._blocks.items():
      b = int(pos[0]) % blocks.shape[0]
      payload = blocks[b:b + 1, lo:hi]
      dims = ('init_time', 'lead_time') + self._extra + (
          'latitude', 'longitude')
      coords = dict(self._grid, init_time=np.asarray(init_times),
                    lead_time=self._lead[lo:hi])
      for d, n in zip(self._extra, payload.shape[2:2 + len(self._extra)]):
        coords[d] = np.arange(n)
      out[var] = xl.DataArray(payload, dims, coords=coords, name=var)
    return out


def run_c5(args, dev, rank, world, peak):
  """config[4] (C5) through the product path at this many GPUs: the
  deterministic suite (RMSE, MSE, MAE, Bias, ACC) and the ensemble suite
  (CRPS, spread/skill) at 0.25 degree, streamed from pinned HOST memory by
  pipeline.run_pipeline in (init=1, lead=12) chunks
  (run_benchmark_evaluation.py:96-101,301-361), chunks block-partitioned over
  the ranks, ONE distributed.all_reduce_state per aggregator.  Weak scaling:
  every rank owns `c5_inits` init times.  Analysis rows are kept on the GPU
  (TargetsFromArrays device_cache): consecutive chunks share 10 of their 12
  valid times, so only the forecasts and 2 new analysis rows per chunk cross
  PCIe.  Checked inside the leg: sharded result == the same evaluation done
  by rank 0 alone (shard=False), and a torch float64 recomputation of one
  variable."""
  import torch
  import torch.distributed as dist
  from weatherbenchx_b200 import aggregation, pipeline, time_chunks, weighting
  from weatherbenchx_b200 import xarray_lite as xl
  from weatherbenchx_b200.data_loaders import array_loaders
  from weatherbenchx_b200.metrics import deterministic, probabilistic
  n_lead, n_blocks = 12, 4
  # chunks per rank: the ensemble chunks are 25x larger than the deterministic
  # ones, so the deterministic suite gets twice as many
  per_rank_of = {'deterministic': 2 * args.c5_inits, 'ensemble': args.c5_inits}
  n_init = max(per_rank_of.values()) * world
  six = np.timedelta64(6, 'h')
  t0 = np.datetime64('2020-01-01T00', 'ns')
  init = t0 + np.arange(n_init) * 2 * six
  lead = (np.arange(n_lead) * six).astype('timedelta64[ns]')
  n_valid = 2 * (n_init - 1) + n_lead
  valid = t0 + np.arange(n_valid) * six
  lat = np.linspace(-90, 90, NLAT)
  lon = np.linspace(0, 360, NLON, endpoint=False)
  grid = {'latitude': lat, 'longitude': lon}
  gen = torch.Generator(device=dev)
  gen.manual_seed(5000)          # the SAME archive on every rank
  det_vars = ('2m_temperature', 'geopotential_500')
  ens_var, members = 'ens_2m_temperature', args.c5_members
  keep = []

  def pinned(shape, fill):
    host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    for i in range(shape[0]):
      host[i].copy_(fill(shape[1:]))
    keep.append(host)
    return host.numpy()

  def normal(scale):
    return lambda shape: torch.empty(shape, device=dev).normal_(
        0.0, scale, generator=gen)

  analyses, clim, blocks = {}, {}, {}
  for name in det_vars:
    analyses[name] = xl.DataArray(
        pinned((n_valid, NLAT, NLON), normal(1.0)),
        ('valid_time', 'latitude', 'longitude'),
        coords=dict(grid, valid_time=valid), name=name)
  for name in det_vars:
    blocks[name] = pinned((n_blocks, n_lead, NLAT, NLON), normal(1.1))
    clim[name] = xl.DataArray(
        torch.empty((366, 4, NLAT, NLON), device=dev).normal_(
            0.0, 0.5, generator=gen),
        ('dayofyear', 'hour', 'latitude', 'longitude'),
        coords=dict(grid, dayofyear=np.arange(1, 367),
                    hour=np.arange(0, 24, 6)), name=name)
  ens_blocks = {ens_var: pinned((1, n_lead, members, NLAT, NLON), normal(1.1))}
  # the ensemble is verified against the analysis of the first variable
  analyses[ens_var] = analyses[det_vars[0]].rename(ens_var)
  torch.cuda.synchronize()
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  suites = {
      'deterministic': (
          {'rmse': deterministic.RMSE(), 'mse': deterministic.MSE(),
           'mae': deterministic.MAE(), 'bias': deterministic.Bias(),
           'acc': deterministic.ACC(clim)},
          lambda it: _RecyclingForecasts(blocks, it, lead, grid),
          det_vars, 4 * n_lead * len(det_vars)),
      'ensemble': (
          {'crps': probabilistic.CRPSEnsemble(
              ensemble_dim='number', use_sort=True),
           'ssr': probabilistic.UnbiasedSpreadSkillRatio(
               ensemble_dim='number')},
          lambda it: _RecyclingForecasts(ens_blocks, it, lead, grid,
                                         extra_dims=('number',)),
          (ens_var,), 4 * n_lead * members),
  }

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  out = {}
  for suite, (metrics, forecasts, variables, fc_bytes_per_pt) in suites.items():
    targets = {v: analyses[v] for v in variables}
    per_rank = per_rank_of[suite]
    suite_init = init[:per_rank * world]
    times = time_chunks.TimeChunks(suite_init, lead, init_time_chunk_size=1,
                                   lead_time_chunk_size=n_lead)

    def run(shard=True, cache=(suite == 'deterministic')):
      # (the ensemble launch streams 51 fields per grid point: its one target
      # field per point is not worth a second memory space)
      loader = array_loaders.TargetsFromArrays(targets, device_cache=cache)
      result = pipeline.run_pipeline(
          times, forecasts(suite_init), loader, metrics, aggregator,
          require_output=False, lanes=args.c5_lanes, shard=shard)
      return result[None][1], loader.uploaded_bytes

    run()                        # warm-up: plans, pinned slots, NCCL
    barrier()
    t_start = time.perf_counter()
    values, target_bytes = run()
    torch.cuda.synchronize()
    seconds = time.perf_counter() - t_start
    if world > 1:
      tmax = torch.tensor([seconds], dtype=torch.float64, device=dev)
      dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
      seconds = float(tmax.item())
    pts_per_rank = per_rank * n_lead * len(variables) * NLAT * NLON
    h2d_per_rank = (pts_per_rank // n_lead // len(variables) *
                    fc_bytes_per_pt + target_bytes)
    entry = {
        'value': world * pts_per_rank / seconds, 'unit': 'grid-points/s',
        'seconds': seconds, 'chunks_per_rank': per_rank,
        'h2d_GBps_per_rank': h2d_per_rank / seconds / 1e9,
        'h2d_bytes_per_point': h2d_per_rank / pts_per_rank,
        'variables': list(variables), 'metrics': sorted(metrics)}
    # ---- sharded == monolithic: rank 0 evaluates every chunk on its own
    if rank == 0:
      mono, _ = run(shard=False)
      worst = 0.0
      for k in mono:
        a, b = np.asarray(values[k].values), np.asarray(mono[k].values)
        worst = max(worst, float(np.max(np.abs(a - b) / np.abs(b))))
      assert worst <= 1e-9, (suite, worst)
      entry['sharded_vs_monolithic_max_rel_diff'] = worst
      # ---- and one variable against a torch float64 recomputation
      v0 = variables[0]
      w_lat = torch.as_tensor(weighting.GridAreaWeighting().weights(
          analyses[v0]).values, device=dev)[None, :, None]
      an = torch.as_tensor(analyses[v0].values, device=dev).double()
      num = torch.zeros(n_lead, dtype=torch.float64, device=dev)
      if suite == 'ensemble':
        # every init time serves the same ensemble block: its spread term is
        # computed once (sorted-member form of the fair estimator)
        ens = torch.as_tensor(ens_blocks[v0][0], device=dev)
        coef = (2.0 * torch.arange(members, device=dev, dtype=torch.float64)
                - (members - 1))[:, None, None]
        spread_term = torch.zeros(n_lead, dtype=torch.float64, device=dev)
        for j in range(n_lead):
          srt = torch.sort(ens[j].double(), dim=0).values
          spread = 2.0 * (coef * srt).sum(dim=0) / (members * (members - 1))
          spread_term[j] = (spread * w_lat[0]).sum()
          del srt, spread
      for i in range(len(suite_init)):
        tg = an[2 * i:2 * i + n_lead]
        if suite == 'deterministic':
          fc = torch.as_tensor(blocks[v0][i % n_blocks], device=dev).double()
          num += (((fc - tg) ** 2) * w_lat).sum(dim=(1, 2))
          del fc
        else:
          for j in range(n_lead):
            skill = (ens[j].double() - tg[j][None]).abs().mean(dim=0)
            num[j] += (skill * w_lat[0]).sum() - 0.5 * spread_term[j]
            del skill
      den = len(suite_init) * float(w_lat.sum()) * NLON
      ref = (num / den).cpu().numpy()
      key = 'rmse' if suite == 'deterministic' else 'crps'
      got = np.asarray(values[f'{key}.{v0}'].values)
      if suite == 'deterministic':
        ref = np.sqrt(ref)
      np.testing.assert_allclose(got, ref, rtol=1e-5)
      entry['checked'] = (f'{key}.{v0} == torch float64 recomputation '
                          '(rtol 1e-5); N-rank result == rank-0 monolithic run')
      del an
    barrier()
    out[suite] = entry
    torch.cuda.empty_cache()
  del keep
  total_pts = sum(e['value'] * e['seconds'] for e in out.values())
  total_s = sum(e['seconds'] for e in out.values())
  return {
      'workload': (f'config[4]: deterministic suite (RMSE, MSE, MAE, Bias, ACC; '
                   f'{len(det_vars)} vars) + ensemble suite (CRPS use_sort, '
                   f'unbiased spread/skill; {members} members), 0.25 deg, '
                   f'{per_rank_of["deterministic"]} / {per_rank_of["ensemble"]} '
                   f'init x {n_lead} lead per rank from pinned HOST '
                   'memory through pipeline.run_pipeline in (init=1, lead=12) '
                   f'chunks, {args.c5_lanes} evaluation lanes, analysis rows cached on the '
                   'GPU, climatology resident on the GPU, chunks sharded over '
                   'the ranks, one all_reduce_state per suite; wall clock, max '
                   'over ranks'),
      'value': total_pts / total_s, 'unit': 'grid-points/s',
      'n_gpus': world, 'suites': out}


def run_b200(args):
  os.environ.setdefault('NCCL_DEBUG', 'WARN')  # keep stdout to the JSON line
  import torch
  import torch.distributed as dist
  from weatherbenchx_b200 import _cabi, aggregation, weighting
  from weatherbenchx_b200 import xarray_lite as xl
  from weatherbenchx_b200.metrics import deterministic

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise RuntimeError('bench.py needs a B200; there is no CPU fallback')
  from weatherbenchx_b200 import distributed as wbx_distributed
  # rank -> GPU: spread over the PCIe uplinks when GPUs are left over
  device_index = wbx_distributed.device_for_local_rank(
      local_rank, int(os.environ.get('LOCAL_WORLD_SIZE', str(world))))
  local_rank = device_index
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  ctx = _cabi.get_context(local_rank)
  ctx.use_torch_stream()

  lat = np.linspace(-90, 90, NLAT)
  lon = np.linspace(0, 360, NLON, endpoint=False)
  gen = torch.Generator(device=dev)
  gen.manual_seed(1000 + rank)
  # device-resident batch: [var, init, lat, lon]
  tgt = torch.empty((N_VARS, N_INIT, NLAT, NLON), device=dev)
  prd = torch.empty_like(tgt)
  for v in range(N_VARS):
    tgt[v].normal_(280.0, 10.0, generator=gen)
    prd[v].copy_(tgt[v]).add_(torch.empty_like(tgt[v]).normal_(
        0.0, 2.0, generator=gen))
  slab_bytes = NLAT * NLON * 4
  jobs = [(v, i) for v in range(N_VARS) for i in range(N_INIT)]
  gaw = weighting.GridAreaWeighting().weights(
      xl.DataArray(np.zeros(NLAT), ('latitude',), coords={'latitude': lat}))
  plan = _cabi.DetPlan(
      ctx, space=_cabi.SPACE_DEVICE, flags=0, ny=NLAT, nx=NLON,
      pred=np.array([prd.data_ptr() + (v * N_INIT + i) * slab_bytes
                     for v, i in jobs], np.uint64),
      target=np.array([tgt.data_ptr() + (v * N_INIT + i) * slab_bytes
                       for v, i in jobs], np.uint64),
      cell=np.array([v for v, _ in jobs], np.int32), n_cells=N_VARS,
      w_y=gaw.values, stat_mask=1 << _cabi.STAT_SLOT['SquaredError'])
  # Device-resident AggregationState of this rank: [sum_ws | sum_w], packed so
  # that the cross-rank combine is ONE all-reduce over one buffer.
  state = torch.zeros(N_VARS * 10, dtype=torch.float64, device=dev)
  out_ws = state[:N_VARS * 6].view(N_VARS, 6)
  out_w = state[N_VARS * 6:].view(N_VARS, 4)

  def step():
    # one chunk: fused statistic + aggregation, summed into the rank's state on
    # the device (AggregationState.__add__, aggregation.py:84-110)
    plan.run_to_device(out_ws.data_ptr(), out_w.data_ptr(), accumulate=True)

  def combine():
    # the reference's CombinePerKey over all chunks (beam_pipeline.py:509-510):
    # a single all-reduce of the packed sufficient statistics.
    if world > 1:
      dist.all_reduce(state)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(max(args.warmup, 3)):
    step()
  combine()
  barrier()
  # sanity: the number being timed is the right number (rank-local state).
  plan.run_to_device(out_ws.data_ptr(), out_w.data_ptr())
  torch.cuda.synchronize()
  d = (prd[0].double() - tgt[0].double())
  ref = float(((d * d) * torch.as_tensor(gaw.values, device=dev)[None, :, None]
               ).sum())
  got = float(out_ws[0, 2])
  assert abs(got - ref) <= 1e-5 * abs(ref), (got, ref)

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
    time.sleep(0.3)
  barrier()
  launches0 = ctx.kernel_launches()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  win0 = time.time()
  ev0.record()
  state.zero_()
  for _ in range(args.steps):
    step()
  combine()
  ev1.record()
  barrier()
  win1 = time.time()
  # K identical chunks on every rank: the combined SquaredError sum must be
  # K * (sum over ranks of one chunk); checked on rank-local data only here.
  elapsed_ms = ev0.elapsed_time(ev1)
  launches = ctx.kernel_launches() - launches0
  # Per-kernel time for the roofline: a second pass over the same steps with
  # every reduction kernel bracketed by CUDA events on its own stream (the
  # bracketing adds gaps, so it is kept out of the timed region above).
  prof_steps = min(args.steps, 200)
  ctx.profile(True)
  ctx.kernel_time(reset=True)
  for _ in range(prof_steps):
    step()
  barrier()
  kernel_ms, kernel_n = ctx.kernel_time(reset=True)
  ctx.profile(False)
  if world > 1:
    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax.item())
  value = world * POINTS_PER_STEP * args.steps / (elapsed_ms * 1e-3)

  # ---- value_api: the same device-resident workload through the public class
  # API (aggregation.compute_metric_values_for_single_chunk).  The call is
  # asynchronous after its first (planning) evaluation: the loop reads the
  # values of call i - 1 after it has issued call i, the way a chunk loop
  # consumes an asynchronous API; every result is read inside the timed region.
  from weatherbenchx_b200 import distributed, fastpath
  from weatherbenchx_b200.metrics import base as metrics_base
  coords = {'init_time': np.arange(N_INIT), 'latitude': lat, 'longitude': lon}
  dims = ('init_time', 'latitude', 'longitude')
  metrics = {'rmse': deterministic.RMSE()}
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  dev_preds = {n: xl.DataArray(prd[v], dims, coords=coords, name=n)
               for v, n in enumerate(VAR_NAMES)}
  dev_tgts = {n: xl.DataArray(tgt[v], dims, coords=coords, name=n)
              for v, n in enumerate(VAR_NAMES)}

  def api_loop(n):
    previous = None
    for _ in range(n):
      current = aggregation.compute_metric_values_for_single_chunk(
          metrics, aggregator, dev_preds, dev_tgts)
      if previous is not None:
        for name in VAR_NAMES:
          previous[f'rmse.{name}'].values   # pylint: disable=expression-not-assigned
      previous = current
    return {k: float(previous[f'rmse.{k}'].values) for k in VAR_NAMES}

  api_values = api_loop(max(args.warmup, 3))
  # the state above holds the all-reduced sums of every rank: the value the
  # API result is compared with is this rank's own chunk
  plan.run_to_device(out_ws.data_ptr(), out_w.data_ptr())
  torch.cuda.synchronize()
  rmse0 = float(np.sqrt(out_ws[0, 2].item() / out_w[0, 0].item()))
  assert abs(api_values[VAR_NAMES[0]] - rmse0) <= 1e-9 * rmse0, (
      api_values, rmse0)
  barrier()
  a0 = time.time()
  ev0.record()
  api_loop(args.steps)
  ev1.record()
  torch.cuda.synchronize()
  a1 = time.time()
  api_ms = ev0.elapsed_time(ev1)
  if world > 1:
    tmax = torch.tensor([api_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    api_ms = float(tmax.item())
  value_api = {
      'value': world * POINTS_PER_STEP * args.steps / (api_ms * 1e-3),
      'unit': 'grid-points/s', 'ms_per_step': api_ms / args.steps,
      'api': 'aggregation.compute_metric_values_for_single_chunk, device '
             'inputs; replayed plan, values of call i-1 read after call i is '
             'issued (all values read inside the timed region)',
      'replays': fastpath.STATS['replayed']}

  # ---- e2e: public class API, host (pinned) inputs --------------------------
  e2e_steps = max(1, min(args.steps, args.e2e_steps))
  host_p = torch.empty((N_VARS, N_INIT, NLAT, NLON), dtype=torch.float32,
                       pin_memory=True)
  host_t = torch.empty_like(host_p).pin_memory()
  host_p.copy_(prd)
  host_t.copy_(tgt)
  torch.cuda.synchronize()
  preds = {n: xl.DataArray(host_p[v].numpy(), dims, coords=coords, name=n)
           for v, n in enumerate(VAR_NAMES)}
  tgts = {n: xl.DataArray(host_t[v].numpy(), dims, coords=coords, name=n)
          for v, n in enumerate(VAR_NAMES)}

  def e2e_step():
    # one chunk per rank from HOST memory through the class API; the ranks'
    # AggregationStates are combined by the product's collective
    # (distributed.all_reduce_state: the CombinePerKey of the reference,
    # beam_pipeline.py:509-510) and the metric values are formed from the sum.
    statistics = metrics_base.compute_unique_statistics_for_all_metrics(
        metrics, preds, tgts)
    state = aggregator.aggregate_statistics(statistics)
    state = distributed.all_reduce_state(state)
    return state.metric_values(metrics)

  for _ in range(2):
    values = e2e_step()
  if world == 1:
    assert abs(values[f'rmse.{VAR_NAMES[0]}'].item() - rmse0) <= 1e-6 * rmse0
  barrier()
  e0 = time.time()
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    e2e_step()
  torch.cuda.synchronize()
  e2e_s = time.perf_counter() - t0
  e1 = time.time()
  if world > 1:
    tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    e2e_s = float(tmax.item())
  e2e_value = world * POINTS_PER_STEP * e2e_steps / e2e_s

  # ---- the box's host->device ceiling for this many ranks copying at once:
  # plain pinned cudaMemcpyAsync of the same host buffers (no engine code),
  # CUDA events, max over ranks.  e2e cannot exceed it: every step moves
  # h2d_bytes_per_step over PCIe.
  copy_stream = torch.cuda.Stream(dev)
  reps = 4
  best = None
  for _ in range(3):
    barrier()
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
      c0.record()
      for _ in range(reps):
        prd.copy_(host_p, non_blocking=True)
        tgt.copy_(host_t, non_blocking=True)
      c1.record()
    copy_stream.synchronize()
    ms = c0.elapsed_time(c1)
    if world > 1:
      tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
      dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
      ms = float(tmax.item())
    best = ms if best is None else min(best, ms)
  step_bytes = POINTS_PER_STEP * ALG_BYTES_PER_POINT
  h2d_ceiling = world * reps * step_bytes / (best * 1e-3) / 1e9
  h2d_achieved = world * step_bytes * e2e_steps / e2e_s / 1e9

  # ---- config[4] through the product path (every rank takes part)
  c5 = None
  if not args.no_c5:
    del dev_preds, dev_tgts, preds, tgts
    host_p = host_t = None
    fastpath.clear()
    peak_c5, _ = measured_peaks()
    try:
      c5 = run_c5(args, dev, rank, world, peak_c5)
    except Exception as e:  # pylint: disable=broad-except
      c5 = {'error': f'{type(e).__name__}: {e}'}
      if world > 1:
        raise   # a rank that left the leg would dead-lock the others
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  clocks = sampler.stop([(win0, win1), (a0, a1), (e0, e1)])
  peak, peak_src = measured_peaks()
  per_launch_ms = kernel_ms / max(kernel_n, 1)
  achieved = POINTS_PER_STEP * ALG_BYTES_PER_POINT / (per_launch_ms * 1e-3) / 1e9
  traffic, traffic_src = headline_traffic()
  line = {
      'metric': METRIC, 'value': value, 'unit': 'grid-points/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
      'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic',
      'config': make_config(world),
      'notes': {
          'value': 'device-resident chunks: every step adds its chunk into the '
                   "rank's device-resident AggregationState (accumulate=1); ONE "
                   'f64 all-reduce of the packed state after the K chunks, '
                   'inside the timed region',
          'e2e_steps': e2e_steps,
          'devices': [wbx_distributed.device_for_local_rank(r, world)
                      for r in range(world)],
          'placement': 'ranks spread over the visible GPUs (the GPUs 0-3 and '
                       '4-7 of the box share one host uplink each: '
                       'profiles/h2d_ceiling_r2_box8_n8.json)'},
      'e2e': {
          'value': e2e_value, 'unit': 'grid-points/s',
          'ms_per_step': 1e3 * e2e_s / e2e_steps,
          'h2d_bytes_per_step': POINTS_PER_STEP * ALG_BYTES_PER_POINT,
          'd2h_bytes_per_step': N_VARS * 10 * 8,
          'h2d_achieved_gbs': h2d_achieved,
          'h2d_ceiling_gbs': h2d_ceiling,
          'frac_of_ceiling': h2d_achieved / h2d_ceiling,
          'ceiling_how': (f'{world} rank(s) copying the same pinned host '
                          'buffers at once with plain cudaMemcpyAsync (torch '
                          'copy_, no engine code), CUDA events, max over ranks, '
                          'best of 3; aggregate GB/s over all ranks'),
          'api': 'compute_unique_statistics_for_all_metrics + Aggregator.'
                 'aggregate_statistics on pinned host numpy inputs (host-space '
                 'C-ABI plan, H2D inside) + distributed.all_reduce_state + '
                 'metric_values'},
      'value_api': value_api,
      'gpu_launches': int(launches),
      'roofline': {
          'bound': 'hbm', 'kernel': 'det_reduce_tma_kernel<0,0,0,0>',
          'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
          'frac': achieved / peak, 'traffic': traffic,
          'traffic_source': traffic_src, 'peak_source': peak_src,
          'kernel_ms_per_launch': per_launch_ms, 'launches_timed': int(kernel_n),
          'algorithmic_bytes_per_launch': POINTS_PER_STEP * ALG_BYTES_PER_POINT},
      'clocks': clocks,
  }
  if c5 is not None:
    line['c5'] = c5
  if world == 1 and not args.no_cpu_baseline:
    line['cpu_baseline'] = cpu_baseline_sample()
  if world == 1 and not args.no_suite:
    del tgt, prd, plan
    from weatherbenchx_b200 import engine
    engine.clear_plan_cache()
    torch.cuda.empty_cache()
    try:
      line['suite'] = run_suite(ctx, dev, peak, oos_legs=args.oos_legs)
      # the per-configuration roofline fractions, where the driver keeps them
      line['roofline']['secondary'] = {
          name: {'frac': leg['roofline']['frac'],
                 'kernel_ms': leg.get('kernel_ms_per_step'),
                 'step_ms': leg.get('ms_per_step')}
          for name, leg in line['suite'].items()
          if isinstance(leg, dict) and 'frac' in leg.get('roofline', {})}
    except Exception as e:  # pylint: disable=broad-except
      # the secondary workloads must never take the headline line down
      line['suite_error'] = f'{type(e).__name__}: {e}'
  emit_line(line)
  if world > 1:
    dist.destroy_process_group()


def headline_traffic():
  """DRAM bytes per launch of the headline kernel (dram__bytes_read.sum +
  dram__bytes_write.sum of one `ncu --set full` capture of this workload),
  recorded in profiles/traffic.json next to the capture it was read from."""
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles',
                      'traffic.json')
  try:
    with open(path) as f:
      rec = json.load(f)['det_reduce_tma_kernel<0,0,0,0>']
    return rec['dram_bytes_read'] + rec['dram_bytes_write'], rec['source']
  except (OSError, KeyError, ValueError):
    return None, None


_RESULT_FD = None


def claim_stdout():
  """Keeps stdout for the ONE JSON line: whatever libraries print on fd 1
  during the run (NCCL's version banner, for one) is sent to stderr."""
  global _RESULT_FD
  if _RESULT_FD is None:
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit_line(line: dict):
  sys.stdout.flush()
  payload = (json.dumps(line) + '\n').encode()
  if _RESULT_FD is None:
    os.write(1, payload)
  else:
    os.write(_RESULT_FD, payload)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=1000)
  ap.add_argument('--warmup', type=int, default=10)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--e2e-steps', type=int, default=20)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-suite', action='store_true',
                  help='skip the secondary workloads (RMSE+ACC, CRPS)')
  ap.add_argument('--no-c5', action='store_true',
                  help='skip the config[4] pipeline leg')
  ap.add_argument('--c5-inits', type=int, default=12,
                  help='init times per rank of the config[4] leg')
  ap.add_argument('--c5-members', type=int, default=50)
  ap.add_argument('--c5-lanes', type=int, default=2,
                  help='evaluation lanes (threads) of the config[4] leg')
  ap.add_argument('--oos-legs', action='store_true',
                  help='also time the out-of-scope legs (categorical, SEEPS)')
  args = ap.parse_args()
  claim_stdout()
  if args.impl == 'reference':
    if args.steps == 1000:
      args.steps = 20
    if args.warmup == 10:
      args.warmup = 3
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
