"""The binned kernel's host-compiled reduction schedule (csrc/det_bins3.cuh),
checked on the CPU: the tables `wbx_bins_schedule_tables` returns are replayed
by a NumPy emulation of what the kernel and its finalize do with them (slot
4-sums, segmented warp scan with static segment heads, class -> segment
lists) and compared with a direct per-class weighted sum.  No GPU is touched.

The classes stand for the sets of bins a grid point belongs to: the reference
multiplies the statistic by every bin mask inside xr.dot
(aggregation.py:320-335, masks from binning.py:22-49)."""

import ctypes

import numpy as np
import pytest

from weatherbenchx_b200 import _cabi

THREADS = 512


def _tables(cmap, n_classes, ny, nx, part):
  lib = _cabi.load_library()
  slab = ny * nx
  n_parts = -(-slab // part)
  desc = np.zeros(n_parts * THREADS * 2, np.uint32)
  seg_base = np.zeros(n_parts + 1, np.int32)
  class_ptr = np.zeros(n_classes + 1, np.int32)
  class_segs = np.zeros(n_parts * 64, np.int32)
  total = ctypes.c_int32()
  rc = lib.wbx_bins_schedule_tables(
      cmap.ctypes.data, n_classes, ny, nx, part, desc.ctypes.data,
      seg_base.ctypes.data, class_ptr.ctypes.data, class_segs.ctypes.data,
      ctypes.byref(total))
  return rc, dict(desc=desc.reshape(n_parts, THREADS, 2),
                  seg_base=seg_base, class_ptr=class_ptr,
                  class_segs=class_segs, total=total.value)


def _replay(tab, values, weights, part, per_element):
  """Kernel + finalize on the host: per-class sums of `values` ([slab]) with
  the element weights `weights` ([slab]; row-uniform unless per_element)."""
  slab = values.size
  n_parts = tab['desc'].shape[0]
  seg_sums = np.zeros(tab['total'])
  seg_written = np.zeros(tab['total'], np.int64)
  covered = np.zeros(slab, np.int64)
  for s in range(n_parts):
    e_lo = s * part
    length = min(part, slab - e_lo)
    a = tab['desc'][s, :, 0].astype(np.int64)          # [thread]
    b = tab['desc'][s, :, 1].astype(np.int64)
    quad = np.stack([a & 0x3ff, (a >> 10) & 0x3ff], 1)  # [thread, slot]
    sel = np.stack([(a >> 20) & 0xf, (a >> 24) & 0xf], 1)
    first_lane, closes, misc, seg = b & 31, (b >> 5) & 1, (b >> 6) & 1, b >> 7
    assert (quad[sel != 0] * 4 + 3 < length).all()
    # slot value: selected elements of the quad (exact zeros elsewhere)
    idx = e_lo + 4 * quad[..., None] + np.arange(4)
    idx = np.minimum(idx, slab - 1)
    take = (sel[..., None] >> np.arange(4)) & 1
    np.add.at(covered, idx[take == 1], 1)
    v = np.where(take == 1, values[idx], 0.0)
    if per_element:
      slot = (v * weights[idx]).sum(-1)
    else:   # the weight of the quad's first element serves the whole quad
      slot = v.sum(-1) * weights[idx[..., 0]]
    acc = slot.sum(-1).reshape(THREADS // 32, 32)       # the thread's two slots
    first_lane = first_lane.reshape(-1, 32)
    closes = closes.reshape(-1, 32)
    misc = misc.reshape(-1, 32)
    seg = seg.reshape(-1, 32)
    assert (misc == misc[:, :1]).all()                  # warp-uniform
    lane_id = np.arange(32)
    # warps with lane-granular segments: segmented inclusive scan
    scan = acc.copy()
    delta = 1
    while delta < 32:
      up = np.zeros_like(scan)
      up[:, delta:] = scan[:, :-delta]
      scan = scan + np.where(lane_id - delta >= first_lane, up, 0.0)
      delta *= 2
    # other warps: butterfly over the 8 lanes of a group, scan over the groups
    regular = misc[:, 0] == 0
    used = (sel != 0).any(-1).reshape(-1, 32)
    assert (first_lane[regular][used[regular]] % 8 == 0).all()
    group_sum = acc.reshape(-1, 4, 8).sum(-1)           # [warp, group]
    head = first_lane.reshape(-1, 4, 8)[..., 0] >> 3
    g = np.arange(4)
    v1 = group_sum.copy()
    v1[:, 1:] += np.where(g[1:] - 1 >= head[:, 1:], group_sum[:, :-1], 0.0)
    v2 = v1.copy()
    v2[:, 2:] += np.where(g[2:] - 2 >= head[:, 2:], v1[:, :-2], 0.0)
    by_group = np.repeat(v2, 8, axis=1)
    # (a regular warp flags every lane of the closing group: each stores one
    # accumulator; one of them stands for the segment here)
    reg_close = closes.reshape(-1, 4, 8)
    assert (reg_close[regular] == reg_close[regular][..., :1]).all()
    assert not closes[~used].any()                      # unused lanes close nothing
    lane_val = np.where(misc == 1, scan, by_group)
    closing = (closes == 1) & ((misc == 1) | (lane_id % 8 == 7))
    rec = tab['seg_base'][s] + seg[closing]
    assert rec.max(initial=-1) < tab['seg_base'][s + 1]
    np.add.at(seg_sums, rec, lane_val[closing])
    np.add.at(seg_written, rec, 1)
    # the two slots of a thread, and the groups of a segment, share a class
  assert (covered == 1).all()         # every element is in exactly one slot
  assert (seg_written == 1).all()     # every segment is closed by one lane
  n_classes = tab['class_ptr'].size - 1
  out = np.zeros(n_classes)
  seen = np.zeros(tab['total'], np.int64)
  for c in range(n_classes):
    segs = tab['class_segs'][tab['class_ptr'][c]:tab['class_ptr'][c + 1]]
    assert (np.diff(segs) > 0).all()
    seen[segs] += 1
    out[c] = seg_sums[segs].sum()
  assert (seen == 1).all()            # every segment belongs to one class
  return out


def _maps(rng, ny, nx):
  yield 'single', np.zeros((ny, nx), np.int64)
  bands = (np.arange(ny) * 3 // ny)[:, None]
  land = np.kron(rng.random(((ny + 7) // 8, (nx + 10) // 11)) > 0.6,
                 np.ones((8, 11), bool))[:ny, :nx]
  yield 'coast', bands * 2 + land
  region = ((np.arange(nx) >= 37) & (np.arange(nx) <= 90))[None, :] & (
      np.arange(ny) > ny // 3)[:, None]
  yield 'edges', land * 4 + region * 2 + (np.arange(nx) == 37)[None, :]
  yield 'noise', rng.integers(0, 7, (ny, nx))


@pytest.mark.parametrize('shape', [(32, 64), (128, 256), (96, 146), (240, 484),
                                   (721, 1440)])
@pytest.mark.parametrize('per_element', [False, True])
def test_schedule_replay_matches_direct_sums(shape, per_element):
  ny, nx = shape
  if (ny * nx) % 16:
    pytest.skip('binned plans need slab % 16 == 0')
  if not per_element and nx % 4:
    pytest.skip('row weights need rows that are a multiple of four long')
  rng = np.random.default_rng(ny * 7 + nx)
  w_y = rng.uniform(0.2, 1.0, ny)
  w_x = rng.uniform(0.2, 1.0, nx) if per_element else None
  values = rng.normal(size=ny * nx)
  for name, raw in _maps(rng, ny, nx):
    _, inv = np.unique(raw, return_inverse=True)
    cmap = inv.reshape(-1).astype(np.uint8)
    n_classes = int(cmap.max()) + 1
    w = (w_y[:, None] * (w_x[None, :] if per_element else np.ones((1, nx)))
         ).reshape(-1)
    want = np.bincount(cmap, weights=values * w, minlength=n_classes)
    served = 0
    # (the 0.25 degree grid takes seconds per replay: its own part size and
    # one small one)
    for part in ((3520, 1024, 512) if ny * nx > 500_000 else
                 (4096, 3520, 1024, 64)):
      rc, tab = _tables(cmap, n_classes, ny, nx, part)
      if rc != 0:      # too many boundary quads for this part size
        assert name in ('noise', 'edges', 'coast') and part > 1024, (name, part)
        continue
      served += 1
      got = _replay(tab, values, w, part, per_element)
      np.testing.assert_allclose(got, want, rtol=0,
                                 atol=1e-12 * np.abs(values).sum(),
                                 err_msg=f'{name} part {part}')
    assert served >= 2, name


def test_schedule_rejects_bad_requests():
  lib = _cabi.load_library()
  cmap = np.zeros(64, np.uint8)
  buf = np.zeros(8192, np.uint32)
  out = ctypes.c_int32()
  tail = (buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data,
          ctypes.byref(out))
  call = lib.wbx_bins_schedule_tables
  assert call(cmap.ctypes.data, 1, 4, 16, 24, *tail) != 0     # part % 16
  assert call(cmap.ctypes.data, 1, 4, 16, 8192, *tail) != 0   # part > 4096
  bad = np.full(64, 3, np.uint8)
  assert call(bad.ctypes.data, 1, 4, 16, 64, *tail) != 0      # class >= n
  assert call(cmap.ctypes.data, 1, 4, 16, 64, *tail) == 0
  assert out.value == 1
