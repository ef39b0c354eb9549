"""ctypes binding of libwbx_b200.so (the C ABI declared in include/wbx_b200.h).

The library is loaded lazily, once per process; contexts are cached per
(process, device).  Nothing here is stored on Metric / Aggregator instances, so
those stay picklable (Beam pickles them, beam_pipeline.py:150-159).

There is deliberately no CPU fallback: if the shared library is missing or no
B200 is visible, the calls raise.
"""

from __future__ import annotations

import ctypes
import os
import threading
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int32, c_int64,
                    c_size_t, c_uint64, c_void_p)

import numpy as np

from weatherbenchx_b200 import _build

WBX_OK = 0
ABI_VERSION = 2
SPACE_DEVICE, SPACE_HOST = 0, 1
FLAG_SKIPNA, FLAG_MASKED, FLAG_FORCE_LDG, FLAG_FORCE_TMA = 1, 2, 16, 32
FLAG_CLIM_DEVICE = 64
FLAG_TARGET_DEVICE, FLAG_MASK_DEVICE = 128, 256
FLAG_BINS_V1 = 512
KERNEL_TMA, KERNEL_LDG4, KERNEL_LDG1, KERNEL_BINS_V1, KERNEL_BINS_V2 = range(5)
KERNEL_BINS_V3 = 5
NUM_DET_STATS = 6
NUM_DET_WCLASSES = 4
STAT_SLOT = {
    'Error': 0,
    'AbsoluteError': 1,
    'SquaredError': 2,
    'SquaredPredictionAnomaly': 3,
    'SquaredTargetAnomaly': 4,
    'AnomalyCovariance': 5,
}
# sum_weights class of each statistic slot (see wbx_b200.h).
STAT_WCLASS = {0: 0, 1: 0, 2: 0, 3: 1, 4: 2, 5: 3}
# Categorical transform of the operands (wbx_det_desc.xform, WBX_XF_*).
XF_CONTINGENCY, XF_ERROR_EXCEEDANCE = 1, 2
XF_PRED_NONZERO, XF_TARGET_NONZERO = 16, 32
NUM_XF_STATS = 4
XF_BINARIZED_PRED = 4
# slot of each categorical statistic in an XF launch; every slot of such a
# launch shares one NaN pattern, i.e. sum_weights class 0.
XF_SLOT = {
    'TruePositives': 0,
    'FalsePositives': 1,
    'FalseNegatives': 2,
    'TrueNegatives': 3,
    'ErrorExceedance': 0,
}
XF_REQUEST = {
    'TruePositives': XF_CONTINGENCY,
    'FalsePositives': XF_CONTINGENCY,
    'FalseNegatives': XF_CONTINGENCY,
    'TrueNegatives': XF_CONTINGENCY,
    'ErrorExceedance': XF_ERROR_EXCEEDANCE,
}

_ERR_NAMES = {
    -1: 'WBX_ERR_INVALID', -2: 'WBX_ERR_CUDA', -3: 'WBX_ERR_NOMEM',
    -4: 'WBX_ERR_UNSUPPORTED', -5: 'WBX_ERR_NO_DEVICE',
}


class WbxError(RuntimeError):
  """A libwbx_b200 call failed."""

  def __init__(self, code: int, message: str):
    super().__init__(f'{_ERR_NAMES.get(code, code)}: {message}')
    self.code = code


class DetDesc(ctypes.Structure):
  """Mirror of wbx_det_desc."""
  _fields_ = [
      ('space', c_int32), ('flags', c_int32),
      ('n_jobs', c_int64), ('ny', c_int64), ('nx', c_int64),
      ('n_cells', c_int64),
      ('pred', POINTER(c_uint64)), ('target', POINTER(c_uint64)),
      ('clim', POINTER(c_uint64)), ('mask', POINTER(c_uint64)),
      ('cell', POINTER(c_int32)),
      ('w_outer', POINTER(c_double)), ('w_y', POINTER(c_double)),
      ('w_x', POINTER(c_double)),
      ('stat_mask', c_int32), ('n_classes', c_int32),
      ('class_map', POINTER(ctypes.c_uint8)),
      ('xform', c_int32), ('reserved', c_int32),
      ('thr_pred', POINTER(ctypes.c_float)),
      ('thr_target', POINTER(ctypes.c_float)),
  ]


MAX_DIMS = 8
MAX_FACTORS = 6
DTYPE_F64, DTYPE_F32, DTYPE_U8 = 0, 1, 2


class GenericDesc(ctypes.Structure):
  """Mirror of wbx_generic_desc."""
  _fields_ = [
      ('ndim', c_int32), ('op', c_int32), ('flags', c_int32),
      ('n_factors', c_int32),
      ('size', c_int64 * MAX_DIMS), ('reduced', c_int32 * MAX_DIMS),
      ('a', c_void_p), ('a_stride', c_int64 * MAX_DIMS),
      ('b', c_void_p), ('b_stride', c_int64 * MAX_DIMS),
      ('c', c_void_p), ('c_stride', c_int64 * MAX_DIMS),
      ('mask', c_void_p), ('mask_stride', c_int64 * MAX_DIMS),
      ('factor', c_void_p * MAX_FACTORS),
      ('factor_dtype', c_int32 * MAX_FACTORS),
      ('factor_stride', (c_int64 * MAX_DIMS) * MAX_FACTORS),
  ]


CRPS_FAIR, CRPS_SKIPNA_ENSEMBLE, CRPS_USE_SORT = 256, 512, 1024
# elementwise-only statistic codes of wbx_det_elementwise
EW_PASS_PRED, EW_PASS_PRED_NAN_TARGET, EW_ACCUMULATE = 16, 17, 256


class CrpsDesc(ctypes.Structure):
  """Mirror of wbx_crps_desc."""
  _fields_ = [
      ('space', c_int32), ('flags', c_int32),
      ('n_jobs', c_int64), ('ny', c_int64), ('nx', c_int64),
      ('n_members', c_int64), ('member_stride', c_int64),
      ('point_stride', c_int64), ('n_cells', c_int64),
      ('ens', POINTER(c_uint64)), ('target', POINTER(c_uint64)),
      ('mask', POINTER(c_uint64)), ('cell', POINTER(c_int32)),
      ('w_outer', POINTER(c_double)), ('w_y', POINTER(c_double)),
      ('w_x', POINTER(c_double)),
      ('stat_mask', c_int32), ('reserved', c_int32),
  ]


class CrpsPointDesc(ctypes.Structure):
  """Mirror of wbx_crps_point_desc."""
  _fields_ = [
      ('ndim', c_int32), ('flags', c_int32),
      ('n_members', c_int64), ('member_stride', c_int64),
      ('size', c_int64 * MAX_DIMS), ('ens_stride', c_int64 * MAX_DIMS),
      ('target_stride', c_int64 * MAX_DIMS),
      ('ens', c_void_p), ('target', c_void_p),
      ('variance', c_void_p), ('unbiased_mse', c_void_p),
  ]


class SpectrumDesc(ctypes.Structure):
  """Mirror of wbx_spectrum_desc."""
  _fields_ = [
      ('n_jobs', c_int64), ('ny', c_int64), ('nx', c_int64),
      ('field', POINTER(c_uint64)), ('row_scale', POINTER(c_double)),
      ('spectrum', c_void_p),
  ]


# name -> (restype, argtypes); every symbol declared in include/wbx_b200.h.
SIGNATURES = {
    'wbx_abi_version': (c_int, []),
    'wbx_last_error': (c_char_p, []),
    'wbx_ctx_create': (c_int, [c_int, POINTER(c_void_p)]),
    'wbx_ctx_destroy': (c_int, [c_void_p]),
    'wbx_ctx_set_stream': (c_int, [c_void_p, c_void_p]),
    'wbx_ctx_synchronize': (c_int, [c_void_p]),
    'wbx_ctx_info': (c_int, [c_void_p, POINTER(c_int), POINTER(c_uint64),
                             POINTER(c_uint64)]),
    'wbx_ctx_profile': (c_int, [c_void_p, c_int32]),
    'wbx_ctx_kernel_time': (c_int, [c_void_p, POINTER(c_double),
                                    POINTER(c_uint64), c_int32]),
    'wbx_ctx_set_staging_bytes': (c_int, [c_void_p, c_uint64]),
    'wbx_host_alloc': (c_int, [c_size_t, POINTER(c_void_p)]),
    'wbx_host_free': (c_int, [c_void_p]),
    'wbx_host_register': (c_int, [c_void_p, c_size_t]),
    'wbx_host_unregister': (c_int, [c_void_p]),
    'wbx_det_plan_create': (c_int, [c_void_p, POINTER(DetDesc),
                                    POINTER(c_void_p)]),
    'wbx_det_plan_destroy': (c_int, [c_void_p, c_void_p]),
    'wbx_det_plan_run': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int32, c_int32]),
    'wbx_det_plan_kernel': (c_int, [c_void_p, c_void_p, POINTER(c_int32)]),
    'wbx_bins_schedule_tables': (c_int, [
        c_void_p, c_int32, c_int64, c_int64, c_int32, c_void_p, c_void_p,
        c_void_p, c_void_p, POINTER(c_int32)]),
    'wbx_det_reduce': (c_int, [c_void_p, POINTER(DetDesc), c_void_p,
                               c_void_p]),
    'wbx_det_elementwise': (c_int, [c_void_p, c_int32, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_void_p]),
    'wbx_xf_elementwise': (c_int, [c_void_p, c_int32, c_int32, ctypes.c_float,
                                   ctypes.c_float, c_void_p, c_void_p, c_int64,
                                   c_void_p]),
    'wbx_struct_layout': (c_int, [c_int32, POINTER(c_uint64), c_int32]),
    'wbx_seeps_elementwise': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int64, ctypes.c_float,
                                      c_int64, c_void_p]),
    'wbx_crps_plan_create': (c_int, [c_void_p, POINTER(CrpsDesc),
                                     POINTER(c_void_p)]),
    'wbx_crps_plan_destroy': (c_int, [c_void_p, c_void_p]),
    'wbx_crps_plan_run': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int32, c_int32]),
    'wbx_crps_plan_run_fields': (c_int, [c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int32,
                                         POINTER(c_void_p)]),
    'wbx_crps_pointwise': (c_int, [c_void_p, POINTER(CrpsPointDesc), c_void_p,
                                   c_void_p]),
    'wbx_ensemble_mean': (c_int, [c_void_p, POINTER(CrpsPointDesc), c_void_p]),
    'wbx_zonal_spectrum': (c_int, [c_void_p, POINTER(SpectrumDesc)]),
    'wbx_reduce_generic': (c_int, [c_void_p, POINTER(GenericDesc), c_void_p,
                                   c_void_p, c_int32]),
}

_lock = threading.Lock()
_lib = None
_lib_pid = None
_contexts: dict = {}


def library_path() -> str:
  return os.environ.get('WBX_B200_LIBRARY', str(_build.LIB_PATH))


def load_library():
  """Loads (once per process) and returns the ctypes CDLL."""
  global _lib, _lib_pid
  with _lock:
    if _lib is not None and _lib_pid == os.getpid():
      return _lib
    path = library_path()
    if not os.path.exists(path):
      raise RuntimeError(
          f'{path} not found. Build it with `python -c "import '
          '__graft_entry__ as g; g.build()"` (needs nvcc). There is no CPU '
          'fallback for the statistic/aggregation kernels.')
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(lib, name)  # AttributeError if the symbol is missing
      fn.restype = restype
      fn.argtypes = argtypes
    abi = lib.wbx_abi_version()
    if abi != ABI_VERSION:
      raise RuntimeError(f'libwbx_b200 ABI {abi} != {ABI_VERSION}')
    _lib, _lib_pid = lib, os.getpid()
    _contexts.clear()
    return lib


def check(code: int):
  if code != WBX_OK:
    msg = load_library().wbx_last_error()
    raise WbxError(code, msg.decode() if msg else '')


class Context:
  """Owns one wbx_ctx (a device, a stream, scratch + staging buffers)."""

  def __init__(self, device: int = 0):
    self.lib = load_library()
    handle = c_void_p()
    check(self.lib.wbx_ctx_create(device, ctypes.byref(handle)))
    self.handle = handle
    self.device = device
    sm = c_int()
    hbm = c_uint64()
    check(self.lib.wbx_ctx_info(self.handle, ctypes.byref(sm),
                                ctypes.byref(hbm), None))
    self.sm_count = sm.value
    self.hbm_bytes = hbm.value

  def close(self):
    if self.handle:
      self.lib.wbx_ctx_destroy(self.handle)
      self.handle = None

  def set_stream(self, cuda_stream: int | None):
    check(self.lib.wbx_ctx_set_stream(self.handle, c_void_p(cuda_stream or 0)))

  def use_torch_stream(self):
    import torch  # pylint: disable=g-import-not-at-top
    # torch's default stream has handle 0, which the C ABI reads as "use the
    # context's own stream"; cudaStreamLegacy (0x1) names it explicitly.
    handle = torch.cuda.current_stream(self.device).cuda_stream
    self.set_stream(handle if handle else 1)

  def synchronize(self):
    check(self.lib.wbx_ctx_synchronize(self.handle))

  def kernel_launches(self) -> int:
    n = c_uint64()
    check(self.lib.wbx_ctx_info(self.handle, None, None, ctypes.byref(n)))
    return n.value

  def profile(self, enable: bool = True):
    check(self.lib.wbx_ctx_profile(self.handle, 1 if enable else 0))

  def kernel_time(self, reset: bool = True):
    """(total milliseconds, launches) of the main kernels since last reset."""
    ms, n = c_double(), c_uint64()
    check(self.lib.wbx_ctx_kernel_time(self.handle, ctypes.byref(ms),
                                       ctypes.byref(n), 1 if reset else 0))
    return ms.value, n.value

  def set_staging_bytes(self, nbytes: int):
    check(self.lib.wbx_ctx_set_staging_bytes(self.handle, nbytes))


_lane = threading.local()


def set_thread_lane(index: int) -> None:
  """Gives the calling thread its own context (streams, staging buffers,
  scratch) per device: lane 0 is the default context every thread shares;
  worker threads that evaluate chunks concurrently take lanes 1, 2, ...  A
  wbx_ctx is not re-entrant, distinct contexts are (include/wbx_b200.h)."""
  _lane.index = int(index)


def get_context(device: int | None = None) -> Context:
  """Per-(process, device, thread lane) cached context."""
  if device is None:
    device = int(os.environ.get('LOCAL_RANK', '0')) if os.environ.get(
        'WBX_B200_DEVICE_FROM_LOCAL_RANK') else 0
    try:
      import torch  # pylint: disable=g-import-not-at-top
      if torch.cuda.is_available():
        device = torch.cuda.current_device()
    except ImportError:
      pass
  load_library()
  key = (os.getpid(), device, getattr(_lane, 'index', 0))
  with _lock:
    ctx = _contexts.get(key)
  if ctx is None:
    ctx = Context(device)
    with _lock:
      _contexts[key] = ctx
  return ctx


def _as_ptr(arr: np.ndarray | None, ctype):
  if arr is None:
    return ctypes.cast(None, POINTER(ctype))
  return arr.ctypes.data_as(POINTER(ctype))


class DetPlan:
  """wbx_det_plan: job tables uploaded once, runnable many times."""

  def __init__(self, ctx: Context, *, space: int, flags: int, ny: int, nx: int,
               pred: np.ndarray, target: np.ndarray, cell: np.ndarray,
               n_cells: int, clim: np.ndarray | None = None,
               mask: np.ndarray | None = None,
               w_outer: np.ndarray | None = None,
               w_y: np.ndarray | None = None, w_x: np.ndarray | None = None,
               stat_mask: int = 0, class_map: np.ndarray | None = None,
               n_classes: int = 0, xform: int = 0,
               thr_pred: np.ndarray | None = None,
               thr_target: np.ndarray | None = None):
    self.ctx = ctx
    self.n_cells = int(n_cells)
    self.n_classes = int(n_classes)
    keep = []

    def prep(a, dtype):
      if a is None:
        return None
      a = np.ascontiguousarray(a, dtype=dtype)
      keep.append(a)
      return a

    pred, target = prep(pred, np.uint64), prep(target, np.uint64)
    clim, mask = prep(clim, np.uint64), prep(mask, np.uint64)
    cell = prep(cell, np.int32)
    w_outer, w_y, w_x = (prep(w_outer, np.float64), prep(w_y, np.float64),
                         prep(w_x, np.float64))
    thr_pred, thr_target = prep(thr_pred, np.float32), prep(thr_target,
                                                            np.float32)
    n_jobs = len(pred)
    for name, arr, n in (('target', target, n_jobs), ('clim', clim, n_jobs),
                         ('mask', mask, n_jobs), ('cell', cell, n_jobs),
                         ('w_outer', w_outer, n_jobs), ('w_y', w_y, ny),
                         ('w_x', w_x, nx), ('thr_pred', thr_pred, n_jobs),
                         ('thr_target', thr_target, n_jobs)):
      if arr is not None and len(arr) != n:
        raise ValueError(f'{name} has {len(arr)} entries, expected {n}')
    desc = DetDesc(
        space=space, flags=flags, n_jobs=n_jobs, ny=ny, nx=nx,
        n_cells=n_cells,
        pred=_as_ptr(pred, c_uint64), target=_as_ptr(target, c_uint64),
        clim=_as_ptr(clim, c_uint64), mask=_as_ptr(mask, c_uint64),
        cell=_as_ptr(cell, c_int32), w_outer=_as_ptr(w_outer, c_double),
        w_y=_as_ptr(w_y, c_double), w_x=_as_ptr(w_x, c_double),
        stat_mask=stat_mask, n_classes=n_classes,
        class_map=_as_ptr(prep(class_map, np.uint8), ctypes.c_uint8),
        xform=xform, thr_pred=_as_ptr(thr_pred, ctypes.c_float),
        thr_target=_as_ptr(thr_target, ctypes.c_float))
    handle = c_void_p()
    check(ctx.lib.wbx_det_plan_create(ctx.handle, ctypes.byref(desc),
                                      ctypes.byref(handle)))
    self.handle = handle
    del keep

  def run_to_host(self):
    """Runs the plan; returns (sum_ws [n_cells, 6], sum_w [n_cells, 4])."""
    rows = self.n_cells * max(self.n_classes, 1)
    ws = np.empty((rows, NUM_DET_STATS), np.float64)
    w = np.empty((rows, NUM_DET_WCLASSES), np.float64)
    check(self.ctx.lib.wbx_det_plan_run(
        self.ctx.handle, self.handle, ws.ctypes.data, w.ctypes.data,
        SPACE_HOST, 0))
    return ws, w

  def run_to_device(self, ws_ptr: int, w_ptr: int, accumulate: bool = False):
    """Asynchronous run into device buffers (f64 [n_cells*6], [n_cells*4])."""
    check(self.ctx.lib.wbx_det_plan_run(
        self.ctx.handle, self.handle, c_void_p(ws_ptr), c_void_p(w_ptr),
        SPACE_DEVICE, 1 if accumulate else 0))

  def kernel(self) -> int:
    """WBX_KERNEL_* code of the kernel that serves this plan."""
    out = c_int32()
    check(self.ctx.lib.wbx_det_plan_kernel(self.ctx.handle, self.handle,
                                           ctypes.byref(out)))
    return out.value

  def close(self):
    if self.handle:
      self.ctx.lib.wbx_det_plan_destroy(self.ctx.handle, self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


def det_elementwise(ctx: Context, stat: int, pred_ptr: int, target_ptr: int,
                    clim_ptr: int | None, n: int, out_ptr: int):
  check(ctx.lib.wbx_det_elementwise(
      ctx.handle, stat, c_void_p(pred_ptr), c_void_p(target_ptr),
      c_void_p(clim_ptr or 0), n, c_void_p(out_ptr)))


def xf_elementwise(ctx: Context, xform: int, slot: int, thr_pred: float,
                   thr_target: float, pred_ptr: int, target_ptr: int | None,
                   n: int, out_ptr: int):
  check(ctx.lib.wbx_xf_elementwise(
      ctx.handle, xform, slot, float(thr_pred), float(thr_target),
      c_void_p(pred_ptr), c_void_p(target_ptr or 0), n, c_void_p(out_ptr)))


def seeps_elementwise(ctx: Context, pred_ptr: int, target_ptr: int,
                      wet_ptr: int, p1_ptr: int, p1_len: int,
                      dry_threshold: float, n: int, out_ptr: int):
  check(ctx.lib.wbx_seeps_elementwise(
      ctx.handle, c_void_p(pred_ptr), c_void_p(target_ptr), c_void_p(wet_ptr),
      c_void_p(p1_ptr), p1_len, float(dry_threshold), n, c_void_p(out_ptr)))


def struct_layout(which: int) -> list:
  """[sizeof, offsetof(member 0), ...] of descriptor struct `which` as the
  loaded library was compiled (pure host code, works without a GPU)."""
  buf = (c_uint64 * 64)()
  n = load_library().wbx_struct_layout(which, buf, 64)
  if n < 0:
    check(n)
  return [int(buf[i]) for i in range(n)]


class CrpsPlan:
  """wbx_crps_plan: CRPSSkill + CRPSSpread aggregated in one launch."""

  def __init__(self, ctx: Context, *, space: int, flags: int, ny: int, nx: int,
               n_members: int, member_stride: int, point_stride: int,
               ens: np.ndarray, target: np.ndarray, cell: np.ndarray,
               n_cells: int, mask: np.ndarray | None = None,
               w_outer: np.ndarray | None = None,
               w_y: np.ndarray | None = None, w_x: np.ndarray | None = None,
               stat_mask: int = 0):
    self.ctx = ctx
    self.n_cells = int(n_cells)
    keep = []

    def prep(a, dtype):
      if a is None:
        return None
      a = np.ascontiguousarray(a, dtype=dtype)
      keep.append(a)
      return a

    ens, target, mask = (prep(ens, np.uint64), prep(target, np.uint64),
                         prep(mask, np.uint64))
    cell = prep(cell, np.int32)
    w_outer, w_y, w_x = (prep(w_outer, np.float64), prep(w_y, np.float64),
                         prep(w_x, np.float64))
    n_jobs = len(ens)
    for name, arr, n in (('target', target, n_jobs), ('mask', mask, n_jobs),
                         ('cell', cell, n_jobs), ('w_outer', w_outer, n_jobs),
                         ('w_y', w_y, ny), ('w_x', w_x, nx)):
      if arr is not None and len(arr) != n:
        raise ValueError(f'{name} has {len(arr)} entries, expected {n}')
    desc = CrpsDesc(
        space=space, flags=flags, n_jobs=n_jobs, ny=ny, nx=nx,
        n_members=n_members, member_stride=member_stride,
        point_stride=point_stride, n_cells=n_cells,
        ens=_as_ptr(ens, c_uint64), target=_as_ptr(target, c_uint64),
        mask=_as_ptr(mask, c_uint64), cell=_as_ptr(cell, c_int32),
        w_outer=_as_ptr(w_outer, c_double), w_y=_as_ptr(w_y, c_double),
        w_x=_as_ptr(w_x, c_double), stat_mask=stat_mask)
    handle = c_void_p()
    code = ctx.lib.wbx_crps_plan_create(ctx.handle, ctypes.byref(desc),
                                        ctypes.byref(handle))
    if code != WBX_OK:
      msg = (ctx.lib.wbx_last_error() or b'').decode()
      if 'n_ensemble < 2' in msg:
        raise ValueError(msg)  # probabilistic.py:210-212
      raise WbxError(code, msg)
    self.handle = handle
    del keep

  def run_to_host(self):
    """(sum_ws [n_cells, 4], sum_w [n_cells, 4]); columns: CRPSSkill,
    CRPSSpread, EnsembleVariance, UnbiasedEnsembleMeanSquaredError."""
    ws = np.empty((self.n_cells, 4), np.float64)
    w = np.empty((self.n_cells, 4), np.float64)
    check(self.ctx.lib.wbx_crps_plan_run(
        self.ctx.handle, self.handle, ws.ctypes.data, w.ctypes.data,
        SPACE_HOST, 0))
    return ws, w

  def run_fields(self, field_ptrs):
    """As run_to_host, also storing the per-point value of slot s to the device
    buffer field_ptrs[s] ([n_jobs, ny*nx] float32; None = skip)."""
    ws = np.empty((self.n_cells, 4), np.float64)
    w = np.empty((self.n_cells, 4), np.float64)
    ptrs = (c_void_p * 4)(*[c_void_p(p or 0) for p in field_ptrs])
    check(self.ctx.lib.wbx_crps_plan_run_fields(
        self.ctx.handle, self.handle, ws.ctypes.data, w.ctypes.data,
        SPACE_HOST, ptrs))
    return ws, w

  def run_to_device(self, ws_ptr: int, w_ptr: int, accumulate: bool = False):
    check(self.ctx.lib.wbx_crps_plan_run(
        self.ctx.handle, self.handle, c_void_p(ws_ptr), c_void_p(w_ptr),
        SPACE_DEVICE, 1 if accumulate else 0))

  def close(self):
    if self.handle:
      self.ctx.lib.wbx_crps_plan_destroy(self.ctx.handle, self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass
