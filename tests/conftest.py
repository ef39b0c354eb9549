"""pytest configuration: the `gpu` marker and import paths."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, 'oracle'), os.path.dirname(__file__)):
  if path not in sys.path:
    sys.path.insert(0, path)


def pytest_configure(config):
  config.addinivalue_line(
      'markers', 'gpu: needs a B200 (run with `-m gpu` on the GPU box)')


def _has_gpu() -> bool:
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:  # pylint: disable=broad-except
    return False


def pytest_collection_modifyitems(config, items):
  if _has_gpu():
    return
  skip = pytest.mark.skip(reason='no CUDA device visible')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
