"""Binned kernel (csrc/det_bins3.cuh) on the public 34-bin class map, through
the C ABI: kernel time and fraction of the HBM roofline over layouts
(latitude-major / longitude-major), jobs per output cell (20: reduced
init_time; 1: the (init=1, lead=12) chunks of the public benchmark),
statistic sets, mask and climatology operands.

  python profiles/exp_bins.py [reps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from weatherbenchx_b200 import _cabi, binning, weighting
from weatherbenchx_b200 import xarray_lite as xl

NLAT, NLON = 721, 1440
REGIONS = {
    'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
    'northern-hemisphere': ((20, 90), (0, 360)),
    'southern-hemisphere': ((-90, -20), (0, 360)),
    'europe': ((35, 75), (-12.5, 42.5)),
    'north-america': ((25, 60), (240, 285)),
    'north-atlantic': ((25, 65), (290, 350)),
    'north-pacific': ((25, 60), (145, 230)),
    'east-asia': ((25, 60), (102.5, 150)), 'ausnz': ((-45, -12.5), (120, 175)),
    'arctic': ((60, 90), (0, 360)), 'antarctic': ((-90, -60), (0, 360)),
    'northern-africa': ((5, 32.5), (-12.5, 37.5)),
    'southern-africa': ((-30, 5), (12.5, 37.5)),
    'south-america': ((-40, 5), (-75, -45)),
    'west-asia': ((15, 60), (42.5, 102.5)),
    'south-east-asia': ((-12.5, 25), (95, 125))}


def public_class_map():
  lat = np.linspace(-90, 90, NLAT)
  lon = np.linspace(0, 360, NLON, endpoint=False)
  rng = np.random.default_rng(7)
  land = xl.DataArray(
      np.kron(rng.random((21, 24)) > 0.7, np.ones((35, 60), bool))[:NLAT],
      ('latitude', 'longitude'), coords={'latitude': lat, 'longitude': lon})
  stat = xl.DataArray(np.zeros((NLAT, NLON), np.float32),
                      ('latitude', 'longitude'),
                      coords={'latitude': lat, 'longitude': lon})
  masks = binning.Regions(REGIONS, land_sea_mask=land).create_bin_mask(stat)
  m = np.asarray(masks.values).reshape(-1, NLAT, NLON)
  key = np.zeros((NLAT, NLON), np.uint64)
  for b in range(m.shape[0]):
    key |= m[b].astype(np.uint64) << np.uint64(b)
  _, inv = np.unique(key.ravel(), return_inverse=True)
  w = weighting.GridAreaWeighting().weights(
      xl.DataArray(np.zeros(NLAT), ('latitude',), coords={'latitude': lat}))
  return inv.reshape(NLAT, NLON).astype(np.uint8), np.asarray(w.values)


def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
  only = os.environ.get('EXP_ONLY')
  peak = 6534.5
  try:
    with open(os.path.join(os.path.dirname(__file__), '..',
                           'MEASURED_PEAKS.json')) as f:
      peak = float(json.load(f).get('hbm_gbs', peak))
  except (OSError, ValueError):
    pass
  cmap, w_lat = public_class_map()
  few = os.environ.get('EXP_CLASSES')
  if few:   # few classes: bands of the public map's class numbers
    cmap = (cmap.astype(np.int64) * int(few) // (int(cmap.max()) + 1)
            ).astype(np.uint8)
  n_classes = int(cmap.max()) + 1
  ctx = _cabi.get_context(0)
  ctx.use_torch_stream()
  n_jobs = 100
  dev = torch.device('cuda', 0)
  gen = torch.Generator(device=dev)
  gen.manual_seed(1)
  t = torch.empty((n_jobs, NLAT * NLON), device=dev).normal_(280, 10, generator=gen)
  p = t + torch.empty_like(t).normal_(0, 2, generator=gen)
  c = torch.empty((n_jobs, NLAT * NLON), device=dev).normal_(280, 5, generator=gen)
  mask = (torch.rand((n_jobs, NLAT * NLON), device=dev, generator=gen) > 0.1
          ).to(torch.uint8)
  step = NLAT * NLON * 4
  addr = lambda x, s: (np.uint64(x.data_ptr()) +  # noqa: E731
                       np.arange(n_jobs, dtype=np.uint64) * np.uint64(s))
  cases = []
  for layout in ('lat_major', 'lon_major'):
    for per_cell in (20, 1):
      for name, stat_mask, clim, masked in (
          ('se', 0b100, False, False), ('e_ae_se', 0b111, False, False),
          ('se_masked', 0b100, False, True),
          ('acc6', 0b111111, True, False),
          ('acc6_masked', 0b111111, True, True)):
        cases.append((layout, per_cell, name, stat_mask, clim, masked))
  for layout, per_cell, name, stat_mask, clim, masked in cases:
    tag = f'{layout}/{per_cell}/{name}'
    if only and only not in tag:
      continue
    if layout == 'lat_major':
      ny, nx, cm, w_y, w_x = NLAT, NLON, cmap.reshape(-1), w_lat, None
    else:
      ny, nx, cm, w_y, w_x = NLON, NLAT, np.ascontiguousarray(cmap.T).reshape(-1), None, w_lat
    cell = (np.arange(n_jobs) // per_cell).astype(np.int32)
    plan = _cabi.DetPlan(
        ctx, space=_cabi.SPACE_DEVICE,
        flags=_cabi.FLAG_MASKED if masked else 0,
        ny=ny, nx=nx, pred=addr(p, step), target=addr(t, step),
        clim=addr(c, step) if clim else None,
        mask=addr(mask, NLAT * NLON) if masked else None, cell=cell,
        n_cells=int(cell.max()) + 1, w_y=w_y, w_x=w_x, stat_mask=stat_mask,
        class_map=cm, n_classes=n_classes)
    n_cells = int(cell.max()) + 1
    out_ws = torch.zeros((n_cells * n_classes, 6), dtype=torch.float64, device=dev)
    out_w = torch.zeros((n_cells * n_classes, 4), dtype=torch.float64, device=dev)
    for _ in range(3):
      plan.run_to_device(out_ws.data_ptr(), out_w.data_ptr())
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.kernel_time(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      plan.run_to_device(out_ws.data_ptr(), out_w.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    kernel_ms, kernel_n = ctx.kernel_time(reset=True)
    ctx.profile(False)
    bytes_pt = 8 + (4 if clim else 0) + (1 if masked else 0)
    pts = n_jobs * NLAT * NLON
    kms = kernel_ms / max(kernel_n, 1)
    # first SquaredError sum of the biggest class against torch float64
    d = (p[0].double() - t[0].double())
    k0 = int(np.bincount(cm).argmax())
    sel = torch.as_tensor(cm == k0, device=dev)
    wfull = torch.as_tensor(
        np.repeat(w_lat, NLON) if layout == 'lat_major' else np.tile(w_lat, NLON),
        device=dev)
    ok = None
    if per_cell == 1:
      m0 = mask[0].double() if masked else 1.0
      ref = float((d * d * wfull * m0)[sel].sum())
      got = float(out_ws[k0, 2])
      ok = abs(got - ref) <= 1e-6 * abs(ref)
    print(json.dumps({
        'case': tag, 'kernel': plan.kernel(), 'kernel_ms': kms,
        'step_ms': e0.elapsed_time(e1) / reps,
        'gpts_per_s': pts / (kms * 1e-3) / 1e9,
        'hbm_frac': pts * bytes_pt / (kms * 1e-3) / 1e9 / peak,
        'bytes_per_point': bytes_pt, 'checked': ok}), flush=True)
    plan.close()


if __name__ == '__main__':
  main()
