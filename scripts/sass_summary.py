"""Per-kernel counts of the SASS mnemonics that identify the tensor / copy paths (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk, 1-D TMA), UTMALDG (tensor-map TMA),
HMMA (mma.sync), LDGSTS (cp.async).   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gennbv_b200", "libgennbv_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "HMMA", "LDGSTS", "SYNCS"]
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for n in names:
            if re.search(r"\b" + n + r"[\.\s]", line):
                counts[cur][n] += 1
print("kernel".ljust(64) + "".join(n.rjust(9) for n in names))
for k, c in counts.items():
    if sum(c[n] for n in names[:8]):
        print(k[:63].ljust(64) + "".join(str(c[n]).rjust(9) for n in names))
