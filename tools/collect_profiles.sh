#!/bin/bash
# After a `tools/gpu_visit.sh <tag> ... ncu` visit: turn the .ncu-rep files into the committed text summaries + the traffic file bench.py reads.
#   bash tools/collect_profiles.sh <tag>
tag=$1
commit=$(git rev-parse --short HEAD)
for k in fwd_dw fwd_philox enc_fwd bwd_tc enc_bwd_sweep heads_fwd; do
  rep=gpurun_out/${tag}_$k.ncu-rep
  [ -f $rep ] || continue
  python profiles/summarize_ncu.py $rep > profiles/${tag}_$k.txt
done
python - "$tag" "$commit" <<'PY'
import csv, io, json, os, subprocess, sys
tag, commit = sys.argv[1], sys.argv[2]
names = {'fwd_dw': ('euler_fwd_tc_kernel<1,0>', 204800, 61), 'fwd_philox': ('euler_fwd_tc_kernel<0,0>', 204800, 61),
         'enc_fwd': ('enc_fwd_tc_kernel', 21504, 21), 'bwd_tc': ('euler_bwd_tc_kernel<0>', 204800, 61),
         'enc_bwd_sweep': ('enc_bwd_sweep_kernel<0>', 21504, 21), 'heads_fwd': ('heads_fwd_kernel', 204800, 60)}
out = {"_comment": "DRAM traffic per launch from ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum); bench.py copies the "
                   "matching entry into roofline.traffic", "commit": commit}
for k, (name, rows, steps) in names.items():
    rep = f'gpurun_out/{tag}_{k}.ncu-rep'
    if not os.path.isfile(rep):
        continue
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    m = dict(zip(r[0], zip(r[1], r[2])))
    def val(key):
        u, v = m[key]
        v = float(v.replace(',', ''))
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    out[name] = {"rows": rows, "steps": steps, "dram_bytes": int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum')),
                 "source": f"profiles/{tag}_{k}.txt"}
json.dump(out, open(f'profiles/{tag}_traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
PY
