"""One binned RMSE evaluation (public 34-bin configuration, 5 vars x 20 init x
721 x 1440) for `ncu -k regex:det_reduce_bins`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from weatherbenchx_b200 import aggregation, binning, weighting, xarray_lite as xl
from weatherbenchx_b200.metrics import deterministic
NLAT, NLON = 721, 1440
lat = np.linspace(-90, 90, NLAT); lon = np.linspace(0, 360, NLON, endpoint=False)
rng = np.random.default_rng(7)
land = xl.DataArray(np.kron(rng.random((21, 24)) > 0.7, np.ones((35, 60), bool))[:NLAT],
                    ('latitude', 'longitude'), coords={'latitude': lat, 'longitude': lon})
regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
  'northern-hemisphere': ((20, 90), (0, 360)), 'southern-hemisphere': ((-90, -20), (0, 360)),
  'europe': ((35, 75), (-12.5, 42.5)), 'north-america': ((25, 60), (240, 285)),
  'north-atlantic': ((25, 65), (290, 350)), 'north-pacific': ((25, 60), (145, 230)),
  'east-asia': ((25, 60), (102.5, 150)), 'ausnz': ((-45, -12.5), (120, 175)),
  'arctic': ((60, 90), (0, 360)), 'antarctic': ((-90, -60), (0, 360)),
  'northern-africa': ((5, 32.5), (-12.5, 37.5)), 'southern-africa': ((-30, 5), (12.5, 37.5)),
  'south-america': ((-40, 5), (-75, -45)), 'west-asia': ((15, 60), (42.5, 102.5)),
  'south-east-asia': ((-12.5, 25), (95, 125))}
coords = {'init_time': np.arange(20), 'latitude': lat, 'longitude': lon}
dims = ('init_time', 'latitude', 'longitude')
P, T = {}, {}
for v in range(5):
  t = torch.empty((20, NLAT, NLON), device='cuda').normal_(280, 10)
  P[f'v{v}'] = xl.DataArray(t + torch.randn_like(t), dims, coords=coords, name=f'v{v}')
  T[f'v{v}'] = xl.DataArray(t, dims, coords=coords, name=f'v{v}')
agg = aggregation.Aggregator(reduce_dims=list(dims), weigh_by=[weighting.GridAreaWeighting()],
                             bin_by=[binning.Regions(regions, land_sea_mask=land)])
for _ in range(4):
  out = aggregation.compute_metric_values_for_single_chunk({'rmse': deterministic.RMSE()}, agg, P, T)
  print(float(out['rmse.v0'].values[0]))
