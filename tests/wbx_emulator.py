"""NumPy interpreter of the C ABI's plan contract (TEST INFRASTRUCTURE).

``installed(monkeypatch)`` replaces ``_cabi.get_context``, ``_cabi.DetPlan``,
``_cabi.CrpsPlan`` and ``engine.seeps_field`` -- the names through which the
host side reaches the CUDA library on host-resident data -- by stand-ins that execute the job tables the planner built
(``include/wbx_b200.h``: wbx_det_desc / wbx_crps_desc) with NumPy on HOST
addresses.  Everything above the C ABI runs unmodified: statistic classes, the
Aggregator's grouping of statistics into launches, the planner (slab layout,
job / cell tables, weights, class maps), merged launches, the unpacking of the
flat result buffers into labelled AggregationStates.

Purpose: host-side logic can be checked on the GPU-less build box against the
reference's golden vectors.  It is never used by the product, by ``-m gpu``
tests, ``smoke()`` or ``bench.py``; the kernels themselves are only ever
validated on a B200.
"""

from __future__ import annotations

import ctypes
import warnings

import numpy as np

from weatherbenchx_b200 import _cabi


def _read(addr, n, dtype):
  nbytes = n * np.dtype(dtype).itemsize
  buf = (ctypes.c_char * nbytes).from_address(int(addr))
  return np.frombuffer(buf, dtype=dtype, count=n)


def _xf_values(xform, p, t, thr_p, thr_t):
  """Slots of a categorical launch (wbx_b200.h, WBX_XF_*), float32 compares."""
  nan = np.float32(np.nan)
  if xform & 3 == _cabi.XF_CONTINGENCY:
    bp = (p != 0) if xform & _cabi.XF_PRED_NONZERO else (p > np.float32(thr_p))
    bt = (t != 0) if xform & _cabi.XF_TARGET_NONZERO else (t > np.float32(thr_t))
    ok = ~(np.isnan(p) | np.isnan(t))
    table = [bp & bt, bp & ~bt, ~bp & bt, ~bp & ~bt]
    return [np.where(ok, v.astype(np.float32), nan) for v in table]
  d = np.abs(p - t)
  ok = ~np.isnan(d) & ~np.isnan(np.float32(thr_p))
  ex = np.where(ok, (d > np.float32(thr_p)).astype(np.float32), nan)
  zero = np.zeros_like(ex)
  return [ex, zero, zero, zero]


class _Context:
  device = 0
  handle = None

  def use_torch_stream(self):
    pass

  def synchronize(self):
    pass


class DetPlan:
  """wbx_det_plan_create + wbx_det_plan_run(space=HOST) on host addresses."""

  def __init__(self, ctx, **desc):
    if desc['space'] != _cabi.SPACE_HOST:
      raise RuntimeError('the emulator only reads host memory')
    self.desc = desc
    self.n_cells = int(desc['n_cells'])
    self.n_classes = int(desc.get('n_classes') or 0)
    self.keepalive = None

  def run_to_host(self):
    d = self.desc
    ny, nx = d['ny'], d['nx']
    slab = ny * nx
    ncls = max(self.n_classes, 1)
    rows = self.n_cells * ncls
    ws = np.zeros((rows, _cabi.NUM_DET_STATS))
    w = np.zeros((rows, _cabi.NUM_DET_WCLASSES))
    wy = d['w_y'] if d.get('w_y') is not None else np.ones(ny)
    wx = d['w_x'] if d.get('w_x') is not None else np.ones(nx)
    wgt = (np.asarray(wy)[:, None] * np.asarray(wx)[None, :]).reshape(-1)
    skipna = bool(d['flags'] & _cabi.FLAG_SKIPNA)
    stat_mask = d.get('stat_mask') or 63
    xform = int(d.get('xform') or 0)
    thr_p, thr_t = d.get('thr_pred'), d.get('thr_target')
    if xform:
      if d.get('clim') is not None or self.n_classes:
        raise RuntimeError('xform with clim / class_map: WBX_ERR_UNSUPPORTED')
      stat_mask &= 15 if xform & 3 == _cabi.XF_CONTINGENCY else 1
    classes = (np.asarray(d['class_map']).reshape(-1).astype(np.int64)
               if self.n_classes else np.zeros(slab, np.int64))
    for j in range(len(d['pred'])):
      p = _read(d['pred'][j], slab, np.float32)
      t = _read(d['target'][j], slab, np.float32)
      c = (_read(d['clim'][j], slab, np.float32)
           if d.get('clim') is not None else np.zeros(slab, np.float32))
      m = (_read(d['mask'][j], slab, np.uint8) != 0
           if d.get('mask') is not None else np.ones(slab, bool))
      wo = d['w_outer'][j] if d.get('w_outer') is not None else 1.0
      with np.errstate(invalid='ignore'):
        vals = [p - t, np.abs(p - t), (p - t) ** 2, (p - c) ** 2,
                (t - c) ** 2, (p - c) * (t - c)]
        if xform:
          vals = _xf_values(xform, p, t,
                            None if thr_p is None else thr_p[j],
                            None if thr_t is None else thr_t[j])
      base = int(d['cell'][j]) * ncls
      done_w = set()
      for s, v in enumerate(vals):
        if not stat_mask >> s & 1:
          continue
        valid = m & ~np.isnan(v) if skipna else m
        v64 = np.where(valid, v, 0).astype(np.float64) * wgt
        ws[base:base + ncls, s] += wo * np.bincount(
            classes, weights=v64, minlength=ncls)
        k = 0 if xform else _cabi.STAT_WCLASS[s]
        if k not in done_w:
          done_w.add(k)
          w[base:base + ncls, k] += wo * np.bincount(
              classes, weights=valid * wgt, minlength=ncls)
    if xform:  # one NaN pattern: every sum_weights class holds the same sum
      w[:, 1:] = w[:, :1]
    return ws, w

  def close(self):
    pass


class CrpsPlan:
  """wbx_crps_plan_create + wbx_crps_plan_run(space=HOST) on host addresses."""

  def __init__(self, ctx, **desc):
    if desc['space'] != _cabi.SPACE_HOST:
      raise RuntimeError('the emulator only reads host memory')
    if not (desc['flags'] & _cabi.CRPS_SKIPNA_ENSEMBLE) and (
        desc['n_members'] < 2 and (desc.get('stat_mask') or 15) & 2):
      raise ValueError('Cannot estimate CRPS spread with n_ensemble < 2.')
    self.desc = desc
    self.n_cells = int(desc['n_cells'])
    self.keepalive = None

  def run_to_host(self):
    d = self.desc
    ny, nx, n_mem = d['ny'], d['nx'], d['n_members']
    slab = ny * nx
    ms, ps = int(d['member_stride']), int(d['point_stride'])
    ws = np.zeros((self.n_cells, 4))
    w = np.zeros((self.n_cells, 4))
    wy = d['w_y'] if d.get('w_y') is not None else np.ones(ny)
    wx = d['w_x'] if d.get('w_x') is not None else np.ones(nx)
    wgt = (np.asarray(wy)[:, None] * np.asarray(wx)[None, :]).reshape(-1)
    skipna_stat = bool(d['flags'] & _cabi.FLAG_SKIPNA)
    ens_skipna = bool(d['flags'] & _cabi.CRPS_SKIPNA_ENSEMBLE)
    fair = bool(d['flags'] & _cabi.CRPS_FAIR)
    stat_mask = d.get('stat_mask') or 15
    extent = (n_mem - 1) * ms + (slab - 1) * ps + 1
    for j in range(len(d['ens'])):
      flat = _read(d['ens'][j], extent, np.float32)
      x = np.lib.stride_tricks.as_strided(
          flat, (n_mem, slab), (ms * 4, ps * 4)).astype(np.float32)
      y = _read(d['target'][j], slab, np.float32)
      m = (_read(d['mask'][j], slab, np.uint8) != 0
           if d.get('mask') is not None else np.ones(slab, bool))
      wo = d['w_outer'][j] if d.get('w_outer') is not None else 1.0
      with np.errstate(all='ignore'), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        n = (~np.isnan(x)).sum(0) if ens_skipna else np.full(slab, n_mem)
        mean_fn = np.nanmean if ens_skipna else np.mean
        sum_fn = np.nansum if ens_skipna else np.sum
        var_fn = np.nanvar if ens_skipna else np.var
        skill = mean_fn(np.abs(x - y[None]), axis=0)
        pair = sum_fn(np.abs(x[:, None, :] - x[None, :, :]), axis=(0, 1))
        spread = pair / (n * (n - (1 if fair else 0)))
        var = var_fn(x, axis=0, ddof=1)
        umse = (mean_fn(x, axis=0) - y) ** 2 - var / n
      cell = int(d['cell'][j])
      for s, v in enumerate((skill, spread, var, umse)):
        if not stat_mask >> s & 1:
          continue
        valid = m & ~np.isnan(v) if skipna_stat else m
        ws[cell, s] += wo * (np.where(valid, v, 0).astype(np.float64)
                             * wgt).sum()
        w[cell, s] += wo * (valid * wgt).sum()
    return ws, w

  def close(self):
    pass


def seeps_field(predictions, targets, wet_threshold, p1, dry_threshold,
                device=None):
  """Host stand-in for ``engine.seeps_field`` (wbx_seeps_elementwise + the
  device gather of the wet threshold): same arguments, same labelled result,
  values from the documented per-point contract of include/wbx_b200.h."""
  from weatherbenchx_b200 import xarray_lite as xl
  dims = predictions.dims + tuple(
      d for d in targets.dims if d not in predictions.dims)
  sizes = dict(targets.sizes, **predictions.sizes)

  def expand(da):
    order = [d for d in dims if d in da.dims]
    arr = da.transpose(*order).to_numpy()
    shape = [sizes[d] if d in order else 1 for d in dims]
    return np.broadcast_to(arr.reshape(shape), [sizes[d] for d in dims])

  ac = wet_threshold
  clim = ac.climatology
  front = list(ac.clim_time_dims)
  rest = [d for d in clim.dims if d not in front]
  arr = clim.transpose(*(front + rest)).to_numpy()
  gathered = arr[tuple(ac.positions[d] for d in front)]
  wet = expand(xl.DataArray(gathered, tuple(ac.time_dims) + tuple(rest)))
  p = expand(predictions).astype(np.float32)
  t = expand(targets).astype(np.float32)
  q = expand(p1).astype(np.float32)
  thr = np.float32(dry_threshold)
  with np.errstate(all='ignore'):
    def cats(x):
      return [x <= thr, (x > thr) & (x < wet), x >= wet]
    one_minus, two_plus = np.float32(1) - q, np.float32(2) + q
    inv_p1, three_over = np.float32(1) / q, np.float32(3) / two_plus
    half = np.float32(0.5)
    zero = np.zeros_like(q)
    score = [[zero, half * (1 / one_minus), half * (4 / one_minus)],
             [half * inv_p1, zero, half * (3 / one_minus)],
             [half * (inv_p1 + three_over), half * three_over, zero]]
    acc = np.zeros(p.shape, np.float64)
    for f, fc in enumerate(cats(p)):
      for k, tc in enumerate(cats(t)):
        acc += (fc & tc).astype(np.float64) * score[f][k].astype(np.float64)
  out = acc.astype(np.float32)
  out[np.isnan(p) | np.isnan(t) | np.isnan(q)] = np.nan
  coords = xl._merge_coords(predictions, targets, dims)  # pylint: disable=protected-access
  coords.pop('mask', None)
  return xl.DataArray(out, dims, coords=coords, name=predictions.name)


def installed(monkeypatch):
  """Routes the host side's entry points to the interpreter."""
  from weatherbenchx_b200 import engine
  ctx = _Context()
  monkeypatch.setattr(_cabi, 'get_context', lambda device=None: ctx)
  monkeypatch.setattr(_cabi, 'DetPlan', DetPlan)
  monkeypatch.setattr(_cabi, 'CrpsPlan', CrpsPlan)
  monkeypatch.setattr(engine, 'seeps_field', seeps_field)
  return ctx
