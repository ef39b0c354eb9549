"""Prints key raw metrics, stall mix and opcode mix of every kernel in an
ncu report:  python profiles/ncu_inspect.py gpurun_out/x.ncu-rep"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
for r in rows[2:]:
  print('===', r[hdr.index('Kernel Name')][:100])
  for w in WANT:
    if w in hdr:
      print(f'  {w:72s} {r[hdr.index(w)]:>16s} {rows[1][hdr.index(w)]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernels, hdr, body, name = [], None, [], None
for r in rows:
  if r and r[0] == 'Kernel Name':
    if hdr is not None:
      kernels.append((name, hdr, body))
    name, hdr, body = r[1], None, []
    continue
  if r and r[0] == 'Address':
    hdr = r
    continue
  if hdr:
    body.append(r)
if hdr:
  kernels.append((name, hdr, body))
for name, hdr, body in kernels:
  si, ii, so = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
  stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
  agg = {hdr[i]: sum(int(r[i]) for r in body) for i in stalls}
  tot = sum(int(r[si]) for r in body)
  print('=== source:', name[:100])
  print('  warp instr', sum(int(r[ii]) for r in body), 'samples', tot)
  print('  stalls', [(k, round(100 * v / max(tot, 1))) for k, v in
                     sorted(agg.items(), key=lambda kv: -kv[1])[:7]])
  ops = collections.Counter()
  for r in body:
    t = r[so].split()
    ops[t[1] if t[0].startswith('@') else t[0]] += int(r[ii])
  print('  ops', ops.most_common(14))
  for r in sorted(body, key=lambda r: -int(r[si]))[:8]:
    print(f'    {int(r[si]):6d} {int(r[ii]):9d}  {r[so].strip()[:80]}')
