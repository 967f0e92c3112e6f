"""Per-kernel opcode histogram of libosq_b200.so (cuobjdump -sass): the Blackwell-native evidence the profiling guide asks for
(tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG / UBLKCP, mbarrier -> SYNCS, cluster barriers ...).
Run in the build container: python scripts/sass_histogram.py > profiles/r02_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "outlier_suppression_b200", "libosq_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
INTEREST = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTMACCTL", "SYNCS",
            "UCGABAR", "ACQBULK", "REDUX", "MATCH", "VOTE", "SHFL", "ATOMS", "ATOMG", "RED", "HMMA", "IMMA", "MUFU", "LDG", "STG", "LDS", "STS",
            "BAR", "ERRBAR", "MEMBAR", "FENCE", "CCTL"]
cur, hist, arch = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s*\.target\s+(\S+)", line) or re.search(r"arch = (sm_\w+)", line)
    if m and cur is None:
        arch["all"] = m.group(1)
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur]["_total"] += 1
        op = m.group(1)
        for k in INTEREST:
            if op == k or op.startswith(k):
                hist[cur][k] += 1
                break
head = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.strip()
print("# SASS opcode histogram per kernel of outlier_suppression_b200/libosq_b200.so")
print("# ELF images: " + " ".join(head.split()))
print("# columns: kernel | SASS instructions | " + " ".join("%s" % k for k in INTEREST))
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
for name, h in hist.items():
    cols = " ".join("%s=%d" % (k, h[k]) for k in INTEREST if h[k])
    print("%-64s %6d  %s" % (demangle(name)[-64:], h["_total"], cols))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("# library totals: " + " ".join("%s=%d" % (k, tot[k]) for k in INTEREST if tot[k]))
