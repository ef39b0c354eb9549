"""Chunks of (init_time, lead_time) that drive an evaluation.

Same interface as /root/reference/weatherbenchX/time_chunks.py:37-202
(``TimeChunks``, ``TimeChunkOffsets``, ``iter_with_chunk_offsets``): the chunk
list is the product of the init-time chunks and the lead-time chunks, lead
times vary fastest, and every chunk knows its offsets within the full time
axes so that per-chunk results can be put back in place
(pipeline.py; beam_pipeline.py:121-138 in the reference).
"""

from __future__ import annotations

import dataclasses
from typing import Iterator, Optional, Union

import numpy as np

# (init_times, lead_times); lead_times is an array of exact lead times or a
# slice for a lead-time interval.
TimeChunk = tuple


@dataclasses.dataclass(frozen=True)
class TimeChunkOffsets:
  init_time: int
  lead_time: int


def _split(values: np.ndarray, size: int) -> list:
  return [values[i:i + size] for i in range(0, len(values), size)]


class TimeChunks:
  """Iterable of (init_times, lead_times) chunks.

  Args:
    init_times: array of np.datetime64.
    lead_times: array of np.timedelta64 (exact lead times) or a slice of
      np.timedelta64 with start and stop (an interval, end inclusive; no step).
    init_time_chunk_size: chunk length along init_time; None / 0 = one chunk.
    lead_time_chunk_size: chunk length along lead_time; None / 0 = one chunk;
      must be None / 0 for a slice.
  """

  def __init__(self, init_times: np.ndarray,
               lead_times: Union[np.ndarray, slice],
               init_time_chunk_size: Optional[int] = None,
               lead_time_chunk_size: Optional[int] = None):
    for name, size in (('init_time_chunk_size', init_time_chunk_size),
                       ('lead_time_chunk_size', lead_time_chunk_size)):
      if size is not None and size < 0:
        raise ValueError(f'{name}={size} but should be non-negative or None')
    init_times = np.asarray(init_times).astype('datetime64[ns]')
    self._init_time_chunk_size = init_time_chunk_size or len(init_times)
    self._init_time_chunks = _split(init_times,
                                    max(self._init_time_chunk_size, 1))
    if isinstance(lead_times, slice):
      if lead_times.start is None or lead_times.stop is None:
        raise ValueError('Slice start and stop must be specified.')
      if lead_times.step is not None:
        raise ValueError('Slice step must be None.')
      if lead_time_chunk_size:
        raise ValueError('Chunking in lead time not compatible for slice.')
      self._lead_time_chunks = [lead_times]
      self._lead_time_chunk_size = lead_time_chunk_size
    elif isinstance(lead_times, np.ndarray):
      lead_times = lead_times.astype('timedelta64[ns]')
      self._lead_time_chunk_size = lead_time_chunk_size or len(lead_times)
      self._lead_time_chunks = _split(lead_times,
                                      max(self._lead_time_chunk_size, 1))
    else:
      raise ValueError('Lead times must be either np.ndarray or slice.')
    self._init_times = init_times
    self._lead_times = lead_times

  @property
  def init_times(self) -> np.ndarray:
    return self._init_times

  @property
  def lead_times(self) -> Union[np.ndarray, slice]:
    return self._lead_times

  @property
  def init_time_chunk_size(self) -> int:
    return self._init_time_chunk_size

  @property
  def lead_time_chunk_size(self):
    return self._lead_time_chunk_size

  def __len__(self) -> int:
    return len(self._init_time_chunks) * len(self._lead_time_chunks)

  def __getitem__(self, index: int) -> TimeChunk:
    if index < 0 or index >= len(self):
      raise IndexError(f'TimeChunks index out of range: {index}')
    n_lead = len(self._lead_time_chunks)
    return (self._init_time_chunks[index // n_lead],
            self._lead_time_chunks[index % n_lead])

  def __iter__(self) -> Iterator[TimeChunk]:
    return (self[i] for i in range(len(self)))

  def offsets(self, index: int) -> TimeChunkOffsets:
    """Offsets of chunk ``index`` within the full init / lead time arrays."""
    n_lead = len(self._lead_time_chunks)
    return TimeChunkOffsets(
        init_time=self._init_time_chunk_size * (index // n_lead),
        lead_time=(self._lead_time_chunk_size or 0) * (index % n_lead))

  def iter_with_chunk_offsets(self) -> Iterator[tuple]:
    """Yields (TimeChunkOffsets, (init_chunk, lead_chunk))."""
    for index in range(len(self)):
      yield self.offsets(index), self[index]
