import csv,re,collections,sys
rows=list(csv.reader(open('gpurun_out/raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sass__inst_executed_local_loads','dram__bytes_read.sum ','dram__bytes_write.sum ','gpu__dram_throughput.avg.pct','sm__cycles_elapsed.avg ','launch__registers']
for h,u,v in zip(hdr,units,vals):
    if any(k in h+' ' for k in want) or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0)>0.3): print(f"{h:95s} {u:10s} {v}")
rows=list(csv.reader(open('gpurun_out/src.csv')))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=rows[2:]
unit=int(sys.argv[1]) if len(sys.argv)>1 else 780800
tot=sum(int(r[idx['# Samples']] or 0) for r in data)
seg_s=0; seg_i=0
print("samples-before-marker  warp-instr/warp-step  marker")
for i,r in enumerate(data):
    src=r[idx['Source']].strip()
    s=int(r[idx['# Samples']] or 0); n=int(r[idx['Instructions Executed']] or 0)
    mark=None
    if 'SYNCS.PHASECHK' in src: mark='WAIT'
    elif 'LDTM' in src: mark='LDTM'
    elif 'STTM' in src: mark='STTM'
    elif 'BAR.SYNC' in src: mark='BAR'
    elif 'SYNCS.ARRIVE' in src: mark='ARRIVE'
    elif 'USETMAXREG' in src: mark='SETMAXREG'
    elif 'UTCHMMA' in src: mark='MMA'
    elif 'UTMA' in src or 'UBLKCP' in src: mark='TMA'
    if mark:
        if seg_s>250 or (mark=='WAIT' and (s>100 or seg_s>100)):
            print(f"{seg_s:7d} ({100*seg_s/tot:4.1f}%) {seg_i/unit:8.1f}  -> {mark} [{src[:64]}] own {s} exec {n}")
        seg_s=0; seg_i=0
    seg_s+=s; seg_i+=n
