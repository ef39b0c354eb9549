"""Turns ncu outputs brought back in gpurun_out/ into the text summaries kept
under profiles/ (launch list per kernel + key raw metrics of the top kernel).

  python profiles/summarize.py <tag>     # e.g. r1_b
reads  gpurun_out/launches_<tag>.csv, gpurun_out/prof_det_<tag>.ncu-rep
writes profiles/launches_<tag>.txt,  profiles/ncu_det_<tag>.txt
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'sm__inst_executed.sum',
    'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second',
    'dram__cycles_elapsed.avg.per_second',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
]


def launches(tag, pattern='launches'):
  path = os.path.join(ROOT, 'gpurun_out', f'{pattern}_{tag}.csv')
  rows = [r for r in csv.reader(
      l for l in open(path) if not l.startswith('=='))]
  hdr = rows[0]
  ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
  agg = collections.OrderedDict()
  for r in rows[1:]:
    if len(r) > vi:
      agg.setdefault(r[ki], []).append(float(r[vi].replace(',', '')) / 1e3)
  total = sum(sum(v) for v in agg.values())
  out = [f'# ncu --metrics gpu__time_duration.sum --clock-control none '
         f'(cold-cache, serialised; compare SHARES)  tag={tag}',
         f'{"kernel":90s} {"n":>4s} {"mean_us":>10s} {"min_us":>10s} '
         f'{"max_us":>10s} {"share":>7s}']
  for k, v in agg.items():
    out.append(f'{k[:90]:90s} {len(v):4d} {sum(v)/len(v):10.2f} {min(v):10.2f} '
               f'{max(v):10.2f} {100*sum(v)/total:6.1f}%')
  dst = os.path.join(ROOT, 'profiles', f'{pattern}_{tag}.txt')
  open(dst, 'w').write('\n'.join(out) + '\n')
  print('\n'.join(out))


def raw(tag, name='prof_det', out_name='ncu_det'):
  rep = os.path.join(ROOT, 'gpurun_out', f'{name}_{tag}.ncu-rep')
  txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(txt)))
  hdr, units = rows[0], rows[1]
  out = [f'# ncu --set full --clock-control none, tag={tag}']
  for r in rows[2:]:
    out.append('--- ' + r[hdr.index('Kernel Name')])
    for w in WANT:
      if w in hdr:
        i = hdr.index(w)
        out.append(f'  {w:72s} {r[i]:>16s} {units[i]}')
  # stall summary from the source page
  src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(src)))
  hdr, body = None, []
  for r in rows:
    if r and r[0] == 'Kernel Name':
      if hdr is not None:
        break
      continue
    if r and r[0] == 'Address':
      hdr = r
      continue
    if hdr:
      body.append(r)
  if hdr:
    si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
    stalls = [i for i, h in enumerate(hdr)
              if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {hdr[i]: sum(int(r[i]) for r in body) for i in stalls}
    tot = sum(int(r[si]) for r in body)
    out.append(f'  warp-instructions executed: {sum(int(r[ii]) for r in body)}'
               f'  stall samples: {tot}')
    out.append('  top stall reasons: ' + ', '.join(
        f'{k}={100*v/max(tot,1):.0f}%' for k, v in
        sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    out.append('  hottest SASS (samples, executions, instruction):')
    for r in sorted(body, key=lambda r: -int(r[si]))[:12]:
      out.append(f'    {int(r[si]):6d} {int(r[ii]):9d}  {r[hdr.index("Source")].strip()}')
  dst = os.path.join(ROOT, 'profiles', f'{out_name}_{tag}.txt')
  open(dst, 'w').write('\n'.join(out) + '\n')
  print('\n'.join(out))


if __name__ == '__main__':
  tag = sys.argv[1]
  if os.path.exists(os.path.join(ROOT, 'gpurun_out', f'launches_{tag}.csv')):
    launches(tag)
  if os.path.exists(os.path.join(ROOT, 'gpurun_out', f'prof_det_{tag}.ncu-rep')):
    raw(tag)
