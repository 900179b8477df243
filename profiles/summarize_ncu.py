"""Turn an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_dec_r1.ncu-rep > profiles/r1_euler_fwd_tc_kernel.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def first_kernel_block(rows):
    """Source-page CSV of a multi-kernel report repeats a ("Kernel Name", ...) row + header per kernel: keep the first block."""
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    return rows[:starts[1]] if len(starts) > 1 else rows

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
print(f"# ncu summary of {rep}")
print(f"kernel: {m['Kernel Name'][1]}   grid {m.get('launch__grid_size', ('', '?'))[1]} x block {m.get('launch__block_size', ('', '?'))[1]}")
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']
for k in keys:
    if k in m:
        print(f"{k:75s} {m[k][1]:>18s} {m[k][0]}")
print("\n# warp stall reasons (warps stalled per issue-active cycle)")
st = [(h, float(v)) for h, (u, v) in m.items() if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and v]
for h, v in sorted(st, key=lambda t: -t[1])[:10]:
    print(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = first_kernel_block(list(csv.reader(io.StringIO(src))))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
mix = collections.Counter()
for r in data:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[idx['Source']].strip())
    op = s.split()[0] if s else '?'
    mix[op.split('.')[0] if not op.startswith(('UTC', 'LDTM', 'STTM', 'UTMA', 'UBLKCP', 'SYNCS', 'MUFU')) else '.'.join(op.split('.')[:2])] += int(r[idx['Instructions Executed']] or 0)
print("\n# SASS instruction mix (warp-level instructions executed)  — UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA")
for op, n in mix.most_common(24):
    print(f"{op:28s} {n:12d}")
print(f"\n# top stall sites by sampled warp stalls (total samples {tot})")
for i in sorted(range(len(data)), key=lambda i: -int(data[i][idx['# Samples']] or 0))[:18]:
    r = data[i]
    print(f"{int(r[idx['# Samples']]):7d} {100 * int(r[idx['# Samples']]) / tot:5.1f}%  x{r[idx['Instructions Executed']]:>10s}  {r[idx['Source']].strip()[:90]}")
