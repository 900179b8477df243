"""SASS evidence of the Blackwell-native path: per kernel of the shipped library, counts of tcgen05 (UTC*MMA, LDTM, STTM), TMA (UTMALDG,
UTMASTG, UBLKCP), mbarrier (SYNCS) and MUFU instructions.

    python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'trajsde_b200/lib/libtrajsde_b200.so'
commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|MUFU\.\w+|HMMA|SYNCS)\b")
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur:
        for k in pat.findall(line):
            cnt[cur]["MUFU.other" if k.startswith("MUFU.") and k != "MUFU.TANH" else k] += 1
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.TANH", "MUFU.other", "HMMA"]
print(f"# cuobjdump -sass {lib}  (commit {commit}): instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path (none)")
print(f"{'kernel':72s} " + " ".join(f"{c:>10s}" for c in cols))
names = subprocess.run(['c++filt'], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
for (k, c), name in zip(cnt.items(), names):
    name = re.sub(r"trajsde::\(anonymous namespace\)::", "", name).split("(")[0][-72:]
    print(f"{name:72s} " + " ".join(f"{c.get(x, 0):>10d}" for x in cols))
