"""A/B of the fixed-shape zonal spectrum kernels on config[3] (13 levels x 6
variables x 721 x 1440 f32) and on the 0.5 degree grid, CUDA events around the
kernel on its stream, result checked against numpy.fft (float64) on a sample
of rows -- PARITY UNPINNED (no spectrum exists in the reference):

    python profiles/exp_spectrum.py [steps]

WBX_SPECTRUM_KERNEL selects the kernel: "generic" = the runtime-shaped kernel,
"fixed2" = three passes (5, 12 | 6, 12) with the first pass from registers and
the paired split, unset = the library's choice (N = 1440: the two-pass 24 x 30
kernel; N = 720: fixed2).  GPU calls 28 / 30 of round 2 also ran "fixed", the
round-1 fixed-shape kernel ((9, 10, 8) radices, row staged through shared
memory: 0.178 ms), which has been removed since."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weatherbenchx_b200 import _cabi  # noqa: E402
from weatherbenchx_b200 import xarray_lite as xl  # noqa: E402
from weatherbenchx_b200.metrics import spectral  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.cuda.set_device(0)
ctx = _cabi.get_context(0)
ctx.use_torch_stream()
peak = 6534.5
try:
  peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:  # pylint: disable=broad-except
  pass
gen = torch.Generator(device='cuda')
gen.manual_seed(11)
for nlat, nlon in ((721, 1440), (361, 720)):
  n_fields = 13 * 6
  f = torch.empty((n_fields, nlat, nlon), device='cuda')
  f.normal_(0.0, 1.0, generator=gen)
  f += torch.linspace(0, 50, nlon, device='cuda').sin() * 3   # a few strong bins
  lat = np.linspace(-90, 90, nlat)
  field = xl.DataArray(
      f, ('field', 'latitude', 'longitude'),
      coords={'latitude': lat,
              'longitude': np.linspace(0, 360, nlon, endpoint=False)},
      name='u')
  rows = [(0, 0), (3, 100), (77, nlat - 1), (40, nlat // 2), (12, 7)]
  ref = {}
  for (i, y) in rows:
    x = f[i, y].double().cpu().numpy()
    F = np.fft.rfft(x) / nlon
    ref[i, y] = np.abs(F) ** 2 * np.r_[1.0, 2.0 * np.ones(nlon // 2)]
  for which in ('generic', 'fixed2', None) * 2:
    if which:
      os.environ['WBX_SPECTRUM_KERNEL'] = which
    else:
      os.environ.pop('WBX_SPECTRUM_KERNEL', None)
    step = lambda: spectral.zonal_energy_spectrum(field)  # noqa: E731
    out = step(); step()
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.kernel_time(reset=True)
    for _ in range(steps):
      out = step()
    torch.cuda.synchronize()
    kms, kn = ctx.kernel_time(reset=True)
    ctx.profile(False)
    kms /= steps
    got = out.data if hasattr(out, 'data') else out
    got = got.cpu().numpy() if hasattr(got, 'cpu') else np.asarray(got)
    # the metric scales rows by 2 pi R cos(lat); undo for the comparison
    worst = 0.0
    for (i, y) in rows:
      c = 2 * np.pi * spectral.EARTH_RADIUS_M * np.cos(np.deg2rad(lat[y]))
      if abs(c) < 1e-3:
        continue
      g = got[i, y].astype(np.float64) / c
      worst = max(worst, float(np.abs(g - ref[i, y]).max() / ref[i, y].max()))
    pts = n_fields * nlat * nlon
    bpp = 4.0 + 4.0 * (nlon // 2 + 1) / nlon
    print(json.dumps({
        'nlon': nlon, 'kernel': which or 'default',
        'kernel_ms': round(kms, 4), 'gpts_per_s': round(pts / kms / 1e6, 1),
        'hbm_frac': round(pts * bpp / (kms * 1e-3) / 1e9 / peak, 4),
        'max_abs_err_over_row_max': worst}), flush=True)
os.environ.pop('WBX_SPECTRUM_KERNEL', None)
