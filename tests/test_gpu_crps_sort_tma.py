"""The TMA-staged sorting kernel (crps_sort_tma_kernel, M = 50 / 51) through
the C ABI: same sums as the register-fed sorting kernel (WBX_FLAG_FORCE_LDG)
and as a NumPy float64 evaluation of the fair CRPS statistics, including
ragged last tiles, masks, the moment statistics and host-space streaming."""

import numpy as np
import pytest
import torch

from weatherbenchx_b200 import _cabi

pytestmark = pytest.mark.gpu


def _reference(x, y, w_y, mask):
  """float64 skill / spread(fair) / variance / unbiased MSE sums per job."""
  x = x.astype(np.float64)
  y = y.astype(np.float64)
  m = x.shape[1]
  skill = np.abs(x - y[:, None]).mean(axis=1)
  xs = np.sort(x, axis=1)
  coef = (2.0 * np.arange(m) - (m - 1)).reshape(1, m, 1, 1)
  spread = 2.0 * (coef * xs).sum(axis=1) / (m * (m - 1))
  var = x.var(axis=1, ddof=1)
  umse = (x.mean(axis=1) - y) ** 2 - var / m
  w = np.broadcast_to(w_y[None, :, None], y.shape) * (
      1.0 if mask is None else mask)
  stats = np.stack([skill, spread, var, umse], axis=-1)
  ws = (stats * w[..., None]).reshape(len(y), -1, 4).sum(axis=1)
  sw = np.repeat(w.reshape(len(y), -1).sum(axis=1)[:, None], 4, axis=1)
  return ws, sw


@pytest.mark.parametrize('space', ['device', 'host'])
@pytest.mark.parametrize('members', [50, 51])
@pytest.mark.parametrize('masked,stat_mask', [(False, 0b0011), (True, 0b1111),
                                              (False, 0b1111)])
def test_sort_tma_matches_ldg_and_numpy(space, members, masked, stat_mask):
  ny, nx, n_jobs = 37, 52, 5            # 1924 points: ragged last tile
  rng = np.random.default_rng(members + 10 * masked + stat_mask)
  y = rng.normal(size=(n_jobs, ny, nx)).astype(np.float32)
  x = (y[:, None] + rng.normal(size=(n_jobs, members, ny, nx))
       ).astype(np.float32)
  mask = (rng.random(y.shape) > 0.3) if masked else None
  w_y = rng.uniform(0.2, 1.0, ny)
  ctx = _cabi.get_context(0)
  arrays = {'ens': x, 'target': y}
  if masked:
    arrays['mask'] = mask.view(np.uint8)
  keep, tables = [], {}
  for name, arr in arrays.items():
    if space == 'device':
      dev = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
      keep.append(dev)
      base, step = dev.data_ptr(), dev[0].numel() * dev.element_size()
    else:
      host = np.ascontiguousarray(arr)
      keep.append(host)
      base, step = host.ctypes.data, host[0].nbytes
    tables[name] = (np.uint64(base) +
                    np.arange(n_jobs, dtype=np.uint64) * np.uint64(step))
  results = {}
  for flag in (0, _cabi.FLAG_FORCE_LDG):
    plan = _cabi.CrpsPlan(
        ctx, space=_cabi.SPACE_DEVICE if space == 'device' else _cabi.SPACE_HOST,
        flags=(flag | _cabi.CRPS_FAIR | _cabi.CRPS_USE_SORT |
               (_cabi.FLAG_MASKED if masked else 0)),
        ny=ny, nx=nx, n_members=members, member_stride=ny * nx, point_stride=1,
        ens=tables['ens'], target=tables['target'], mask=tables.get('mask'),
        cell=np.arange(n_jobs, dtype=np.int32), n_cells=n_jobs, w_y=w_y,
        stat_mask=stat_mask)
    results[flag] = plan.run_to_host()
    plan.close()
  ws_ref, sw_ref = _reference(x, y, w_y, mask)
  cols = [k for k in range(4) if stat_mask & (1 << k)]
  for flag, (ws, sw) in results.items():
    np.testing.assert_allclose(ws[:, cols], ws_ref[:, cols], rtol=2e-5,
                               atol=1e-5 * np.abs(ws_ref[:, cols]).max(),
                               err_msg=f'flag {flag}')
    np.testing.assert_allclose(sw[:, cols], sw_ref[:, cols], rtol=1e-10)
  a, b = results[0], results[_cabi.FLAG_FORCE_LDG]
  np.testing.assert_allclose(a[0][:, cols], b[0][:, cols], rtol=1e-12)
  np.testing.assert_array_equal(a[1][:, cols], b[1][:, cols])


def test_sort_tma_propagates_nan_members():
  ny, nx, members = 16, 32, 50
  rng = np.random.default_rng(1)
  y = rng.normal(size=(2, ny, nx)).astype(np.float32)
  x = rng.normal(size=(2, members, ny, nx)).astype(np.float32)
  x[1, 7, 3, 5] = np.nan
  X, Y = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
  plan = _cabi.CrpsPlan(
      _cabi.get_context(0), space=_cabi.SPACE_DEVICE,
      flags=_cabi.CRPS_FAIR | _cabi.CRPS_USE_SORT, ny=ny, nx=nx,
      n_members=members, member_stride=ny * nx, point_stride=1,
      ens=np.uint64(X.data_ptr()) + np.arange(2, dtype=np.uint64) * np.uint64(
          members * ny * nx * 4),
      target=np.uint64(Y.data_ptr()) + np.arange(2, dtype=np.uint64) * np.uint64(
          ny * nx * 4),
      cell=np.arange(2, dtype=np.int32), n_cells=2, stat_mask=0b0011)
  ws, _ = plan.run_to_host()
  assert np.isfinite(ws[0, :2]).all() and np.isnan(ws[1, :2]).all()
