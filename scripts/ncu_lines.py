"""Aggregate the warp-stall samples of an ncu --import-source capture per CUDA source line.
    python scripts/ncu_lines.py gpurun_out/prof_physics.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
tot = 0
lines = []
for r in rows[3:]:
    if len(r) >= 8 and r[0].isdigit():
        num = lambda v: int(v) if v.strip().lstrip('-').isdigit() else 0
        s, inst = num(r[6]), num(r[7])
        lines.append((int(r[0]), r[1][:110], s, inst))
        tot += s
print('total samples', tot)
for ln, src, s, inst in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f'{ln:4d} {s:7d} {100 * s / tot:5.1f}% inst={inst:10d}  {src}')
if len(sys.argv) > 3:      # regions: name:a-b,...
    for spec in sys.argv[3].split(','):
        name, rng = spec.split(':')
        a, b = map(int, rng.split('-'))
        print(f'{name:12s} {100 * sum(x[2] for x in lines if a <= x[0] <= b) / tot:5.1f}%')
