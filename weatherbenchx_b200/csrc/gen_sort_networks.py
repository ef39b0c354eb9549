"""Generates sort_networks.inc: fixed-size sorting networks for the CRPS sort
kernel (crps.cu, MFIX = 50 / 51).

    python weatherbenchx_b200/csrc/gen_sort_networks.py

Batcher's odd-even merge sort on index lists of ARBITRARY length (halves split
at n // 2, merges of unequal runs): 395 compare-exchanges for n = 50 and 408
for n = 51, against 421 / 433 that ptxas leaves of the power-of-two network for
64 wires with +inf padding.  The sorted order ends up in a fixed permutation of
the wires, emitted as WBX_SORT<n>_RANKS (rank, wire) so the consumer reads
x[wire] with literal indices.  Every network is checked here on all 0/1 inputs
of a random sample plus random permutations (zero-one principle, sampled).
"""

import os
import random


def odd_even_merge(a, b, ces, tail=None):
  """Merges the sorted runs a and b.  With `tail` (the outermost merge) the
  last layer -- compare-exchanges between ADJACENT ranks whose outputs nothing
  reads again -- is recorded in `tail` as (rank of the minimum, wire, wire)
  as well, so that a consumer can stop before it (emit_split)."""
  if not a or not b:
    return a + b
  if len(a) == 1 and len(b) == 1:
    ces.append((a[0], b[0]))
    if tail is not None:
      tail.append((0, a[0], b[0]))
    return [a[0], b[0]]
  even = odd_even_merge(a[0::2], b[0::2], ces)
  odd = odd_even_merge(a[1::2], b[1::2], ces)
  out = [even[0]]
  ei, oi = 1, 0
  while ei < len(even) and oi < len(odd):
    ces.append((odd[oi], even[ei]))
    if tail is not None:
      tail.append((len(out), odd[oi], even[ei]))
    out += [odd[oi], even[ei]]
    ei += 1
    oi += 1
  return out + odd[oi:] + even[ei:]


def network(n, tail=None):
  ces = []

  def sort(idx, top=False):
    if len(idx) <= 1:
      return idx
    h = len(idx) // 2
    return odd_even_merge(sort(idx[:h]), sort(idx[h:]), ces,
                          tail if top else None)

  order = sort(list(range(n)), top=True)
  return ces, order


def check(ces, order, n, trials=3000):
  rng = random.Random(n)
  for t in range(trials):
    if t % 2:
      x = [rng.randint(0, 1) for _ in range(n)]
    else:
      x = [rng.random() for _ in range(n)]
    want = sorted(x)
    for a, b in ces:
      if x[a] > x[b]:
        x[a], x[b] = x[b], x[a]
    if [x[i] for i in order] != want:
      raise AssertionError(f'network for n={n} does not sort')


def emit(n):
  ces, order = network(n)
  check(ces, order, n)
  lines = [f'// n = {n}: {len(ces)} compare-exchanges',
           f'#define WBX_SORT{n}_SIZE {len(ces)}',
           f'#define WBX_SORT{n}(CE) \\']
  row = []
  for a, b in ces:
    row.append(f'CE({a}, {b})')
    if len(row) == 6:
      lines.append('  ' + ' '.join(row) + ' \\')
      row = []
  lines.append('  ' + ' '.join(row))
  lines.append(f'#define WBX_SORT{n}_RANKS(R) \\')
  row = []
  for rank, wire in enumerate(order):
    row.append(f'R({rank}, {wire})')
    if len(row) == 8:
      lines.append('  ' + ' '.join(row) + ' \\')
      row = []
  lines.append('  ' + ' '.join(row))
  return '\n'.join(lines) + '\n'


def moment_terms(tail, singles, n):
  """sum_q (2q + 1 - n) x_(q) over the wires BEFORE the last layer, written
  with differences only (identical members give exactly zero, as the
  reference's float64 sum does).  A pair of adjacent ranks (q, q + 1) on wires
  (a, b) contributes (2q + 2 - n) (a + b) + |a - b| whichever way round its
  two values are; terms of opposite weight are coupled, what is left over is
  anchored on the rank-0 wire.  Returns (couples, single_couples, anchored
  pairs, anchored singles, anchor wire)."""
  pairs = [(2 * q + 2 - n, a, b) for q, a, b in tail]
  ones = [(2 * q + 1 - n, w) for q, w in singles]
  couples, one_couples = [], []
  rest = []
  by_m = {m: (a, b) for m, a, b in pairs}
  assert len(by_m) == len(pairs)
  for m, a, b in pairs:
    if m > 0 and -m in by_m:
      couples.append((m,) + by_m[-m] + (a, b))
    elif m < 0 and -m in by_m:
      pass
    elif m != 0:
      rest.append((m, a, b))
  by_c = {c: w for c, w in ones}
  rest_ones = []
  for c, w in ones:
    if c > 0 and -c in by_c:
      one_couples.append((c, by_c[-c], w))
    elif c < 0 and -c in by_c:
      pass
    elif c != 0:
      rest_ones.append((c, w))
  anchor = None
  if rest or rest_ones:
    zero = [w for c, w in rest_ones if c == 1 - n]
    assert zero, 'the rank-0 wire must be a single'
    anchor = zero[0]
    rest_ones = [(c, w) for c, w in rest_ones if w != anchor]
    assert (1 - n) + sum(c for c, _ in rest_ones) + 2 * sum(
        m for m, _, _ in rest) == 0
  return couples, one_couples, rest, rest_ones, anchor


def check_split(head, tail, terms, n, trials=2000):
  couples, one_couples, rest, rest_ones, anchor = terms
  rng = random.Random(1000 + n)
  for _ in range(trials):
    x = [rng.choice((0.0, 1.0, rng.random())) for _ in range(n)]
    srt = sorted(x)
    want = sum((2 * q + 1 - n) * v for q, v in enumerate(srt))
    for a, b in head:
      if x[a] > x[b]:
        x[a], x[b] = x[b], x[a]
    got = sum(abs(x[a] - x[b]) for _, a, b in tail)
    got += sum(m * ((x[a2] + x[b2]) - (x[a1] + x[b1]))
               for m, a1, b1, a2, b2 in couples)
    got += sum(c * (x[w2] - x[w1]) for c, w1, w2 in one_couples)
    got += sum(m * (x[a] + x[b] - 2 * x[anchor]) for m, a, b in rest)
    got += sum(c * (x[w] - x[anchor]) for c, w in rest_ones)
    if abs(got - want) > 1e-9 * max(1.0, abs(want)):
      raise AssertionError(f'split moment for n={n}: {got} != {want}')


def emit_split(n):
  """The same network as WBX_SORT<n>, cut before its last layer:
  WBX_SORT<n>_HEAD(CE) with CE(index, a, b), WBX_SORT<n>_TAIL(P) with
  P(wire, wire) for the unsorted adjacent pairs (their |a - b| enters the
  moment) and WBX_SORT<n>_MOMENT(D2, D1, A2, A1) for the signed part:
    D2(m, a1, b1, a2, b2)  m ((x[a2] + x[b2]) - (x[a1] + x[b1]))
    D1(c, w1, w2)          c (x[w2] - x[w1])
    A2(m, a, b, w0)        m (x[a] + x[b] - 2 x[w0])
    A1(c, w, w0)           c (x[w] - x[w0])"""
  tail = []
  ces, order = network(n, tail)
  head = ces[:len(ces) - len(tail)]
  assert ces[len(head):] == [(a, b) for _, a, b in tail]
  in_tail = {w for _, a, b in tail for w in (a, b)}
  singles = [(q, w) for q, w in enumerate(order) if w not in in_tail]
  for q, a, b in tail:
    assert order[q] == a and order[q + 1] == b
  terms = moment_terms(tail, singles, n)
  check_split(head, tail, terms, n)
  couples, one_couples, rest, rest_ones, anchor = terms

  def rows(items, per):
    out, row = [], []
    for it in items:
      row.append(it)
      if len(row) == per:
        out.append('  ' + ' '.join(row) + ' \\')
        row = []
    if row:
      out.append('  ' + ' '.join(row))
    else:
      out[-1] = out[-1][:-2]
    return out

  lines = [f'// n = {n} without the last layer: {len(head)} compare-exchanges,'
           f' {len(tail)} adjacent pairs, {len(singles)} single wires',
           f'#define WBX_SORT{n}_HEAD_SIZE {len(head)}',
           f'#define WBX_SORT{n}_HEAD(CE) \\']
  lines += rows([f'CE({i}, {a}, {b})' for i, (a, b) in enumerate(head)], 5)
  lines.append(f'#define WBX_SORT{n}_TAIL(P) \\')
  lines += rows([f'P({a}, {b})' for _, a, b in tail], 8)
  lines.append(f'#define WBX_SORT{n}_MOMENT(D2, D1, A2, A1) \\')
  lines += rows(
      [f'D2({m}, {a1}, {b1}, {a2}, {b2})' for m, a1, b1, a2, b2 in couples] +
      [f'D1({c}, {w1}, {w2})' for c, w1, w2 in one_couples] +
      [f'A2({m}, {a}, {b}, {anchor})' for m, a, b in rest] +
      [f'A1({c}, {w}, {anchor})' for c, w in rest_ones], 4)
  return '\n'.join(lines) + '\n'


def main():
  here = os.path.dirname(os.path.abspath(__file__))
  body = ('// GENERATED by gen_sort_networks.py -- do not edit.\n'
          '// CE(a, b): x[a] <- min, x[b] <- max.  R(rank, wire): x[wire] is '
          'the rank-th smallest.\n\n' + emit(50) + '\n' + emit(51) + '\n' + emit_split(50) + '\n' +
          emit_split(51))
  with open(os.path.join(here, 'sort_networks.inc'), 'w') as f:
    f.write(body)
  print('wrote sort_networks.inc')


if __name__ == '__main__':
  main()
