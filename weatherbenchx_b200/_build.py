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


def _object_fingerprint(src: pathlib.Path) -> str:
  """Fingerprint of one translation unit: its source, every header and the
  flags (headers are few and shared, so any header change rebuilds all)."""
  h = hashlib.sha256()
  h.update(src.read_bytes())
  for path in sorted(list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.inc')) +
                     list(CSRC.glob('*.h')) + list(INCLUDE.glob('*.h'))):
    h.update(path.name.encode())
    h.update(path.read_bytes())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> pathlib.Path:
  """Compiles every .cu under csrc/ into lib/libwbx_b200.so (if stale): one
  nvcc per translation unit, in parallel, objects cached under lib/obj/; the
  library is linked to a temporary name and renamed into place."""
  from concurrent.futures import ThreadPoolExecutor  # pylint: disable=g-import-not-at-top
  fp = _fingerprint()
  if (not force and LIB_PATH.exists() and STAMP.exists()
      and STAMP.read_text().strip() == fp):
    return LIB_PATH
  nvcc = find_nvcc()
  if nvcc is None:
    raise RuntimeError('nvcc not found; cannot build libwbx_b200.so')
  LIB_DIR.mkdir(exist_ok=True)
  obj_dir = LIB_DIR / 'obj'
  obj_dir.mkdir(exist_ok=True)
  compile_flags = [f for f in NVCC_FLAGS if f != '-shared']

  def compile_one(src: pathlib.Path):
    obj = obj_dir / (src.stem + '.o')
    stamp = obj_dir / (src.stem + '.stamp')
    ofp = _object_fingerprint(src)
    if (not force and obj.exists() and stamp.exists()
        and stamp.read_text().strip() == ofp):
      return obj, None
    cmd = [nvcc, *compile_flags, '-I', str(INCLUDE), '-c', str(src), '-o',
           str(obj)]
    if verbose:
      print(' '.join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
      return obj, (f'nvcc failed ({proc.returncode}) on {src.name}:\n'
                   f'{proc.stdout}\n{proc.stderr}')
    stamp.write_text(ofp)
    return obj, None

  srcs = sources()
  with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
    results = list(ex.map(compile_one, srcs))
  errors = [e for _, e in results if e]
  if errors:
    raise RuntimeError('\n'.join(errors))
  tmp = LIB_DIR / f'.libwbx_b200.{os.getpid()}.so'
  cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared',
         '-Xcompiler', '-fPIC', '-o', str(tmp),
         *[str(obj) for obj, _ in results]]
  if verbose:
    print(' '.join(cmd))
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError(
        f'nvcc link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}')
  os.replace(tmp, LIB_PATH)
  STAMP.write_text(fp)
  return LIB_PATH


if __name__ == '__main__':
  print(build_library(force=True, verbose=True))
