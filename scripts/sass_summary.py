"""Per-kernel counts of the SASS mnemonics that prove which hardware paths libqmcb.so uses (B200_PROFILING.md):
UTMALDG / UBLKCP (TMA loads), UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (TMEM access), UTCBAR (tcgen05 commit), DMMA
(FP64 tensor pipe), HMMA (mma.sync), SYNCS (mbarrier), plus registers per thread.  python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qmcpack_b200", "libqmcb.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTMALDG", "UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "DMMA", "HMMA", "SYNCS", "LDGSTS", "BAR.SYNC", "NANOSLEEP"]
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
print("# cuobjdump -sass %s  (arch: %s)" % (os.path.relpath(so, ROOT), ", ".join(sorted(set(re.findall(r"arch = (\S+)", sass))))))
print("# %-110s %5s %6s  %s" % ("kernel", "regs", "instr", "mnemonic counts"))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for blk, dem in zip(blocks, names):
    mangled = blk.split("\n", 1)[0].strip()
    body = blk
    cnt = collections.Counter()
    ninstr = 0
    for line in body.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        ninstr += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
    short = re.sub(r"\(.*", "", dem.replace("qmcb::", ""))
    short = re.sub(r"^void ", "", short)
    r = regs.get(mangled, ("?", "?"))
    print("%-112s %5s %6d  %s" % (short[:112], r[0], ninstr, " ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k])))
