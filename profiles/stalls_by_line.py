"""Warp-stall samples of an ncu report aggregated by CUDA source line (via nvdisasm line info of the same object file).

    python profiles/stalls_by_line.py <report.ncu-rep> <object.o> <kernel-name-substring> [top_n]
"""
import csv, io, re, subprocess, sys, tempfile, os, collections

def first_kernel_block(rows):
    """Source-page CSV of a multi-kernel report repeats a ("Kernel Name", ...) row + header per kernel: keep the first block."""
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    return rows[:starts[1]] if len(starts) > 1 else rows

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = first_kernel_block(list(csv.reader(io.StringIO(sass))))
h = rows[1]
ia, isamp, isrc, iex = h.index('Address'), h.index('Warp Stall Sampling (All Samples)'), h.index('Source'), h.index('Instructions Executed')
inst = [(r[isrc], int(r[isamp] or 0), int(r[iex] or 0)) for r in rows[2:] if len(r) > isamp]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for ln in dis.splitlines():
    if ln.startswith('.text.'):
        on = kname in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
if len(lines) != len(inst):
    print(f"# warning: {len(lines)} disassembled instructions vs {len(inst)} in the report", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
for (src, s, ex), loc in zip(inst, lines):
    agg[loc][0] += s
    agg[loc][1] += ex
tot = sum(v[0] for v in agg.values())
print(f"# {rep}: {tot} warp-stall samples, by source line (file:line  samples  share  warp-instructions executed)")
srcs = {}
for (loc, (s, ex)) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    text = ''
    if loc:
        path = os.path.join(os.path.dirname(os.path.abspath(obj)), '..', 'csrc', loc[0])
        if os.path.isfile(path):
            srcs.setdefault(path, open(path).read().splitlines())
            text = srcs[path][loc[1] - 1].strip()[:100] if loc[1] - 1 < len(srcs[path]) else ''
    print(f"{str(loc[0]) + ':' + str(loc[1]) if loc else '?':28s} {s:8d} {100 * s / max(tot, 1):5.1f}%  {ex:10d}  {text}")
