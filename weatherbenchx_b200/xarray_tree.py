"""map_structure over nested containers of DataArrays.

Mirrors /root/reference/weatherbenchX/xarray_tree.py:44-68 for the containers
this package produces: dict / Dataset (mapping name -> DataArray), list, tuple.
Leaves for which ``func`` returns None are dropped from mappings, as the
reference does for Datasets.
"""

from __future__ import annotations

from typing import Any, Callable, Mapping

from weatherbenchx_b200 import xarray_lite as xl


def map_structure(func: Callable[..., Any], *structures: Any) -> Any:
  if not callable(func):
    raise TypeError(f'func must be callable, got: {func}')
  if not structures:
    raise ValueError('Must provide at least one structure')
  first = structures[0]
  if isinstance(first, Mapping):
    out = {k: map_structure(func, *[s[k] for s in structures])
           for k in first.keys()}
    if isinstance(first, xl.Dataset):
      return xl.Dataset({k: v for k, v in out.items() if v is not None})
    return out
  if isinstance(first, (list, tuple, set)):
    return type(first)(map_structure(func, *s) for s in zip(*structures))
  return func(*structures)
