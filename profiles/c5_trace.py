"""Timeline of the chunk driver on the C5 leg (one GPU): runs bench.py's C5 leg
with WBX_PIPELINE_TRACE and summarises how the phases of the lanes overlap.

  python profiles/c5_trace.py [lanes]
"""
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lanes = sys.argv[1] if len(sys.argv) > 1 else '2'
trace = os.path.join(ROOT, 'gpurun_out', f'c5_trace_l{lanes}.jsonl')
if os.path.exists(trace):
  os.remove(trace)
env = dict(os.environ, WBX_PIPELINE_TRACE=trace)
subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '3',
                '--warmup', '3', '--no-suite', '--no-cpu-baseline',
                '--c5-lanes', lanes], env=env, check=True,
               stdout=subprocess.DEVNULL)
rows = [json.loads(l) for l in open(trace)]
# runs are separated by long gaps: split on gaps > 50 ms, keep runs with the
# deterministic suite's chunk count
rows.sort(key=lambda r: r['start'])
runs, cur = [], [rows[0]]
for r in rows[1:]:
  if r['start'] - max(x['end'] for x in cur) > 0.05:
    runs.append(cur)
    cur = []
  cur.append(r)
runs.append(cur)
for i, run in enumerate(runs):
  t0 = min(r['start'] for r in run)
  t1 = max(r['end'] for r in run)
  chunks = len({r['chunk'] for r in run})
  per = collections.defaultdict(float)
  for r in run:
    per[(r['thread'], r['phase'])] += r['end'] - r['start']
  print(f'run {i}: {chunks} chunks, {1e3 * (t1 - t0):.1f} ms wall, '
        f'{1e3 * (t1 - t0) / chunks:.2f} ms per chunk')
  for (thread, phase), sec in sorted(per.items()):
    print(f'   {thread:28s} {phase:11s} {1e3 * sec:8.1f} ms total '
          f'{1e3 * sec / chunks:6.2f} ms per chunk of the run')
  if i == 1:   # the timed deterministic run: print the first chunks' timeline
    for r in sorted(run, key=lambda r: r['start'])[:36]:
      print(f"      {1e3 * (r['start'] - t0):8.2f} -> {1e3 * (r['end'] - t0):8.2f} ms  "
            f"chunk {r['chunk']:3d} {r['phase']:11s} {r['thread']}")
