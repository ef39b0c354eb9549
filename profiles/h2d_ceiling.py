"""Host->device copy ceiling of the box, per GPU and for N processes at once.

  python profiles/h2d_ceiling.py                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
      --master-addr 127.0.0.1 --master-port 29531 profiles/h2d_ceiling.py

Plain pinned-memory `cudaMemcpyAsync` (torch `copy_(non_blocking=True)` from a
`pin_memory=True` tensor, and the same from a write-combined `cudaHostAlloc`
buffer), timed with CUDA events on the copy stream after a barrier, max over
ranks.  Legs:

  all      every rank copies at once (what bench.py's e2e step does)
  solo r   only rank r copies (per-GPU link rate)
  pair 0,k ranks 0 and k copy at once (shows which GPUs share an uplink)

Prints one JSON object; `aggregate_GBps` of the `all` leg is the number
bench.py's `e2e.h2d_ceiling_gbs` has to be compared with.
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

BYTES = int(os.environ.get('H2D_BYTES', str(830_592_000)))  # one headline step
REPS = int(os.environ.get('H2D_REPS', '6'))


def cudart():
  for name in ('libcudart.so.12', 'libcudart.so'):
    try:
      return ctypes.CDLL(name)
    except OSError:
      continue
  import glob
  base = os.path.dirname(torch.__file__)
  for path in glob.glob(os.path.join(base, '..', 'nvidia', 'cuda_runtime',
                                     'lib', 'libcudart.so*')):
    return ctypes.CDLL(path)
  return None


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  n = BYTES // 4
  dst = torch.empty(n, dtype=torch.float32, device=dev)
  src = torch.empty(n, dtype=torch.float32, pin_memory=True)
  src.fill_(1.0)
  stream = torch.cuda.Stream(dev)
  flag = torch.zeros(1, device=dev)

  def barrier():
    if world > 1:
      dist.all_reduce(flag)
    torch.cuda.synchronize()

  def timed(copy_fn, active):
    """max-over-ranks seconds of REPS copies by the active ranks."""
    best = None
    for _ in range(2):
      barrier()
      e0 = torch.cuda.Event(enable_timing=True)
      e1 = torch.cuda.Event(enable_timing=True)
      with torch.cuda.stream(stream):
        e0.record()
        if active:
          for _ in range(REPS):
            copy_fn()
        e1.record()
      stream.synchronize()
      ms = e0.elapsed_time(e1) if active else 0.0
      t = torch.tensor([ms], dtype=torch.float64, device=dev)
      if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
      best = float(t.item()) if best is None else min(best, float(t.item()))
    return best * 1e-3

  def whole():
    dst.copy_(src, non_blocking=True)

  slab = 721 * 1440
  def slabs():
    for lo in range(0, n - slab + 1, slab):
      dst[lo:lo + slab].copy_(src[lo:lo + slab], non_blocking=True)

  out = {'world': world, 'bytes_per_copy': BYTES, 'reps': REPS, 'legs': {}}

  def leg(name, copy_fn, ranks):
    active = rank in ranks
    sec = timed(copy_fn, active)
    total = len(ranks) * BYTES * REPS
    out['legs'][name] = {
        'ranks': list(ranks), 'seconds': sec,
        'aggregate_GBps': total / sec / 1e9,
        'per_gpu_GBps': total / sec / 1e9 / len(ranks)}

  everyone = list(range(world))
  leg('all', whole, everyone)
  leg('all_4MB_slabs', slabs, everyone)
  for r in range(world):
    leg(f'solo_{r}', whole, [r])
  for k in range(1, world):
    leg(f'pair_0_{k}', whole, [0, k])
  if world >= 8:
    leg('even', whole, [0, 2, 4, 6])
    leg('first4', whole, [0, 1, 2, 3])

  # write-combined pinned memory (no CPU cache snooping on the DMA reads)
  rt = cudart()
  if rt is not None:
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(BYTES),
                          ctypes.c_uint(4))  # cudaHostAllocWriteCombined
    if rc == 0:
      rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_size_t, ctypes.c_int,
                                     ctypes.c_void_p]
      ctypes.memset(ptr, 0, BYTES)
      def wc():
        rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ptr,
                           ctypes.c_size_t(BYTES), 1,
                           ctypes.c_void_p(stream.cuda_stream))
      leg('all_write_combined', wc, everyone)
      rt.cudaFreeHost(ptr)
  # two copy streams per rank (does one DMA queue leave bandwidth unused?)
  stream2 = torch.cuda.Stream(dev)
  half = n // 2
  def two_streams():
    dst[:half].copy_(src[:half], non_blocking=True)
    with torch.cuda.stream(stream2):
      dst[half:].copy_(src[half:], non_blocking=True)
    stream.wait_stream(stream2)
  leg('all_two_streams', two_streams, everyone)
  # device->host at the same time as host->device (full duplex?)
  back = torch.empty(n, dtype=torch.float32, pin_memory=True)
  def duplex():
    dst.copy_(src, non_blocking=True)
    with torch.cuda.stream(stream2):
      back.copy_(dst, non_blocking=True)
    stream.wait_stream(stream2)
  leg('all_h2d_with_d2h', duplex, everyone)

  if rank == 0:
    try:
      out['cpu_count'] = os.cpu_count()
      out['affinity'] = len(os.sched_getaffinity(0))
    except Exception:  # pylint: disable=broad-except
      pass
    print(json.dumps(out))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  sys.exit(main())
