"""us per launch of the last windowed and the last whole-frame detector call in an ncu launch list of scripts/tag_time_probe.py
(ncu --metrics gpu__time_duration.sum --csv --log-file LIST python scripts/tag_time_probe.py)."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(lines))
hdr = r[0]; ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
rows = [(x[ki][:44], float(x[vi].replace(',', '')) / 1000, x[gi]) for x in r[1:]]
first = [i for i, x in enumerate(rows) if 'minmax' in x[0] or 'tile_max' in x[0]]
half = len(first) // 2
for name, start in (("search windows", first[half - 1]), ("whole frames", first[-1])):
    print(f"# {name}")
    tot = 0.0
    for j, x in enumerate(rows[start:start + 16]):
        if j > 0 and ('minmax' in x[0] or 'tile_max' in x[0]): break
        if 'at::' in x[0]: break
        print(f"{x[0]:46s}{x[1]:8.1f} {x[2]}"); tot += x[1]
    print(f"sum {tot:.1f} us")
