"""Builds libwbx_b200.so in-tree with nvcc for sm_100a (cross-compiles on CPU)."""

from __future__ import annotations

import hashlib
import os
import pathlib
import shutil
import subprocess

PKG_DIR = pathlib.Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
LIB_DIR = PKG_DIR / 'lib'
LIB_PATH = LIB_DIR / 'libwbx_b200.so'
STAMP = LIB_DIR / 'libwbx_b200.stamp'
INCLUDE = PKG_DIR.parent / 'include'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '-shared',
]


def sources() -> list[pathlib.Path]:
  return sorted(CSRC.glob('*.cu'))


def _fingerprint() -> str:
  h = hashlib.sha256()
  for path in sorted(list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.inc')) + list(CSRC.glob('*.h')) +
                     list(INCLUDE.glob('*.h'))):
    h.update(path.name.encode())
    h.update(path.read_bytes())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def find_nvcc() -> str | None:
  nvcc = shutil.which('nvcc')
  if nvcc:
    return nvcc
  for cand in ('/usr/local/cuda/bin/nvcc',):
    if os.path.exists(cand):
      return cand
  return None


def build_library(force: bool = False, verbose: bool = False) -> pathlib.Path:
  """Compiles every .cu under csrc/ into lib/libwbx_b200.so (if stale)."""
  fp = _fingerprint()
  if (not force and LIB_PATH.exists() and STAMP.exists()
      and STAMP.read_text().strip() == fp):
    return LIB_PATH
  nvcc = find_nvcc()
  if nvcc is None:
    raise RuntimeError('nvcc not found; cannot build libwbx_b200.so')
  LIB_DIR.mkdir(exist_ok=True)
  cmd = [nvcc, *NVCC_FLAGS, '-I', str(INCLUDE), '-o', str(LIB_PATH),
         *[str(s) for s in sources()]]
  if verbose:
    print(' '.join(cmd))
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError(
        f'nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}')
  STAMP.write_text(fp)
  return LIB_PATH


if __name__ == '__main__':
  print(build_library(force=True, verbose=True))
