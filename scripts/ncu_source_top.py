"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line: python scripts/ncu_source_top.py rep [kernel-regex] [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if rx:
    cmd += ["--kernel-name", "regex:" + rx]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.defaultdict(lambda: [0, 0, 0, ''])
cur_file, hdr, ci = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr, ci = r, {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) - 3:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def f(k):
        try:
            return float(r[ci[k]] or 0)
        except Exception:
            return 0.0
    key = (cur_file, ln)
    agg[key][0] += f('# Samples')
    agg[key][1] += f('stall_barrier')
    agg[key][2] += f('stall_long_sb')
    agg[key][3] = r[1]
tot = sum(v[0] for v in agg.values())
print('total samples', tot)
for (fn, ln), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{v[0]:6.0f} {100*v[0]/max(tot,1):5.1f}% bar={v[1]:5.0f} lsb={v[2]:5.0f}  {fn}:{ln}: {v[3].strip()[:90]}")
