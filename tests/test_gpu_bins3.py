"""The binned kernel with the host-compiled reduction schedule
(csrc/det_bins3.cuh) through the C ABI: per-(cell, class) sums against NumPy
float64 across grid geometries (slab parts x job groups, sequential rounds), job -> cell layouts,
masks, climatology, per-element weights and class maps with any number of
classes inside a quad."""

import numpy as np
import pytest
import torch

from weatherbenchx_b200 import _cabi

pytestmark = pytest.mark.gpu


def _class_map(rng, ny, nx, kind):
  if kind == 'blocks':      # land-like blocks x latitude bands: <= 2 per 8
    bands = (np.arange(ny) * 3 // ny)[:, None]
    land = np.kron(rng.random(((ny + 7) // 8, (nx + 15) // 16)) > 0.6,
                   np.ones((8, 16), bool))[:ny, :nx]
    cmap = bands * 2 + land
  elif kind == 'columns':   # boundaries at arbitrary columns
    edges = np.sort(rng.choice(np.arange(1, nx), size=5, replace=False))
    cmap = np.broadcast_to(np.searchsorted(edges, np.arange(nx), side='right'),
                           (ny, nx))
    # keep at most two classes per aligned block of 8
    flat = cmap.reshape(-1).copy()
    for b in range(0, flat.size, 8):
      blk = flat[b:b + 8]
      u = np.unique(blk)
      if len(u) > 2:
        blk[~np.isin(blk, u[:2])] = u[1]
    cmap = flat.reshape(ny, nx)
  elif kind == 'edges3':    # a region edge one column after a land edge:
    # some aligned blocks of 8 hold THREE classes (the serial path)
    land = np.kron(rng.random(((ny + 7) // 8, (nx + 11) // 12)) > 0.5,
                   np.ones((8, 12), bool))[:ny, :nx]
    region = (np.arange(nx) > 37)[None, :] & (np.arange(ny) > ny // 3)[:, None]
    cmap = land * 2 + region
  elif kind == 'many':      # 120 classes x 7 accumulators: no table of sums
    cmap = (np.arange(ny)[:, None] * 10 // ny) * 12 + (
        np.arange(nx)[None, :] * 12 // nx)
  elif kind == 'noise':     # up to four classes per quad
    cmap = rng.integers(0, 5, (ny, nx))
  else:
    cmap = np.zeros((ny, nx), np.int64)
  _, inv = np.unique(cmap, return_inverse=True)
  return inv.reshape(-1).astype(np.uint8), int(inv.max()) + 1


def _reference(p, t, c, mask, cell, n_cells, cmap, n_classes, w_o, w_y, w_x,
               stat_mask):
  n_jobs, ny, nx = p.shape
  p, t = p.astype(np.float32), t.astype(np.float32)
  d = (p - t)
  stats = [d, np.abs(d), d * d]
  if c is not None:
    a, b = p - c, t - c
    stats += [a * a, b * b, a * b]
  w = np.ones((n_jobs, ny, nx))
  if w_o is not None:
    w = w * w_o[:, None, None]
  if w_y is not None:
    w = w * w_y[None, :, None]
  if w_x is not None:
    w = w * w_x[None, None, :]
  valid = np.ones_like(w) if mask is None else mask.astype(np.float64)
  ws = np.zeros((n_cells * n_classes, 6))
  sw = np.zeros((n_cells * n_classes, 4))
  key = (cell[:, None] * n_classes + cmap[None, :]).reshape(-1)
  for k, s in enumerate(stats):
    if stat_mask & (1 << k):
      v = np.where(valid > 0, s.astype(np.float64), 0.0) * w
      np.add.at(ws[:, k], key, v.reshape(-1))
  wsum = np.zeros(n_cells * n_classes)
  np.add.at(wsum, key, (valid * w).reshape(-1))
  sw[:] = wsum[:, None]
  return ws, sw


CASES = [
    # ny, nx, n_jobs, cells, map, clim, mask, wx, stat_mask
    (721, 1440, 12, 'per_job', 'blocks', False, False, False, 0b100),
    (721, 1440, 10, 'two', 'blocks', True, False, False, 0b111101),
    (128, 256, 40, 'per_job', 'blocks', False, True, False, 0b111),
    (128, 256, 7, 'one', 'columns', True, True, False, 0b111111),
    (64, 128, 200, 'many', 'blocks', False, False, False, 0b101),
    (64, 128, 3, 'one', 'noise', False, False, False, 0b111),
    (240, 484, 9, 'two', 'blocks', False, False, True, 0b100),
    (144, 146, 16, 'per_job', 'columns', False, True, True, 0b111),
    (32, 64, 5, 'one', 'single', False, False, False, 0b100),
    (721, 1440, 6, 'two', 'edges3', False, False, False, 0b101),
    (128, 256, 9, 'per_job', 'edges3', True, True, False, 0b111111),
    (96, 146, 5, 'one', 'edges3', False, True, True, 0b110),
    (721, 1440, 5, 'per_job', 'noise', False, False, False, 0b111),
    (1440, 721, 6, 'two', 'edges3', False, False, True, 0b100),
    (1440, 721, 4, 'per_job', 'blocks', True, True, True, 0b111111),
    (721, 1440, 3, 'per_job', 'many', True, True, False, 0b111111),
]


@pytest.mark.parametrize('space', ['device', 'host'])
@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join(map(str, c)))
def test_bins3_matches_numpy(case, space):
  ny, nx, n_jobs, cells, kind, clim, masked, wx, stat_mask = case
  assert (ny * nx) % 16 == 0, 'binned plans need slab % 16 == 0'
  rng = np.random.default_rng(100 + CASES.index(case))
  cmap, n_classes = _class_map(rng, ny, nx, kind)
  p = rng.normal(0, 1, (n_jobs, ny, nx)).astype(np.float32)
  t = (p + rng.normal(0, 1, p.shape)).astype(np.float32)
  c = rng.normal(0, 1, p.shape).astype(np.float32) if clim else None
  mask = (rng.random(p.shape) > 0.25) if masked else None
  if cells == 'per_job':
    cell = np.arange(n_jobs)
  elif cells == 'one':
    cell = np.zeros(n_jobs, np.int64)
  elif cells == 'two':
    cell = (np.arange(n_jobs) >= n_jobs // 3).astype(np.int64)
  else:
    cell = np.arange(n_jobs) // 7
  n_cells = int(cell.max()) + 1
  w_o = rng.uniform(0.5, 1.5, n_jobs) if cells != 'per_job' else None
  w_y = rng.uniform(0.2, 1.0, ny)
  w_x = rng.uniform(0.2, 1.0, nx) if wx else None
  ctx = _cabi.get_context(0)
  arrays = {'pred': p, 'target': t}
  if clim:
    arrays['clim'] = c
  if masked:
    arrays['mask'] = mask.view(np.uint8)
  keep, tables = [], {}
  for name, arr in arrays.items():
    if space == 'device':
      dev = torch.from_numpy(arr).cuda()
      keep.append(dev)
      base, step = dev.data_ptr(), dev[0].numel() * dev.element_size()
    else:
      host = np.ascontiguousarray(arr)
      keep.append(host)
      base, step = host.ctypes.data, host[0].nbytes
    tables[name] = np.uint64(base) + np.arange(n_jobs, dtype=np.uint64) * np.uint64(step)
  plan = _cabi.DetPlan(
      ctx, space=_cabi.SPACE_DEVICE if space == 'device' else _cabi.SPACE_HOST,
      flags=_cabi.FLAG_MASKED if masked else 0, ny=ny, nx=nx,
      pred=tables['pred'], target=tables['target'], clim=tables.get('clim'),
      mask=tables.get('mask'), cell=cell.astype(np.int32), n_cells=n_cells,
      w_outer=w_o, w_y=w_y, w_x=w_x, stat_mask=stat_mask, class_map=cmap,
      n_classes=n_classes)
  assert plan.kernel() == _cabi.KERNEL_BINS_V3
  ws, sw = plan.run_to_host()
  again = plan.run_to_host()
  assert again[0].tobytes() == ws.tobytes()   # bit-stable
  assert again[1].tobytes() == sw.tobytes()
  plan.close()
  ws_ref, sw_ref = _reference(p, t, c, mask, cell, n_cells, cmap, n_classes,
                              w_o, w_y, w_x, stat_mask & (0x3f if clim else 7))
  scale = np.abs(ws_ref).max(axis=0, keepdims=True) + 1e-30
  np.testing.assert_allclose(ws / scale, ws_ref / scale, rtol=0, atol=2e-6)
  # element weights (w_x / odd rows) are rounded to float32 inside the kernel
  per_element = wx or nx % 4 != 0
  np.testing.assert_allclose(sw, sw_ref, rtol=1e-7 if per_element else 1e-10)


def test_bins3_nan_stays_in_its_class():
  """A NaN poisons the class it belongs to and no other (the host's
  class-to-bin product then spreads it like the reference's einsum)."""
  ny, nx, n_jobs = 64, 128, 4
  rng = np.random.default_rng(3)
  cmap = (np.arange(ny * nx) % nx >= 60).astype(np.uint8)   # boundary inside a block
  p = rng.normal(size=(n_jobs, ny, nx)).astype(np.float32)
  t = rng.normal(size=(n_jobs, ny, nx)).astype(np.float32)
  p[2, 10, 59] = np.nan            # class 0, three elements before the boundary
  P, T = torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda()
  step = ny * nx * 4
  ctx = _cabi.get_context(0)
  plan = _cabi.DetPlan(
      ctx, space=_cabi.SPACE_DEVICE, flags=0, ny=ny, nx=nx,
      pred=np.uint64(P.data_ptr()) + np.arange(n_jobs, dtype=np.uint64) * np.uint64(step),
      target=np.uint64(T.data_ptr()) + np.arange(n_jobs, dtype=np.uint64) * np.uint64(step),
      cell=np.arange(n_jobs, dtype=np.int32), n_cells=n_jobs, stat_mask=0b100,
      class_map=cmap, n_classes=2)
  ws, sw = plan.run_to_host()
  ws = ws.reshape(n_jobs, 2, 6)[:, :, 2]
  assert np.isnan(ws[2, 0]) and np.isfinite(ws[2, 1])
  assert np.isfinite(ws[[0, 1, 3]]).all()
