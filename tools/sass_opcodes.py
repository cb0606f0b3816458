"""Per-kernel SASS opcode census of librd_b200.so: the Blackwell-native evidence (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma (UTCHMMA kind::f16, UTCQMMA kind::f8f6f4), LDTM/STTM = tcgen05.ld/st,
UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU.* = the XU pipe.
    python tools/sass_opcodes.py [path/to/librd_b200.so] > profiles/r2_sass_opcodes.txt      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ribodetector_b200", "librd_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU", "HMMA", "UCGABAR",
         "LDGSTS", "REDUX", "ATOM", "RED", "BAR", "F2FP", "FFMA")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", txt)
if names:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out)) if len(out) == len(names) else {}
print("# SASS opcode census of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(lib, ROOT))
print("# columns: kernel | total instructions | watched opcodes (with modifiers) : count")
cur, counts, total = None, None, 0


def flush():
    if cur is None:
        return
    name = demangle.get(cur, cur)
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    name = name.replace("(int)", "")
    name = re.sub(r">\(.*", ">", name) if "<" in name else re.sub(r"\(.*", "", name)
    name = {"lstm_tc_kernel<0, 4>": "lstm_tc_kernel<M_FAST>", "lstm_tc_kernel<1, 4>": "lstm_tc_kernel<M_EXACT>",
            "lstm_tc_kernel<2, 4>": "lstm_tc_kernel<M_MIXED>"}.get(name, name)
    items = ["%s:%d" % kv for kv in sorted(counts.items()) if kv[0].split(".")[0] in WATCH]
    print("%-46s %6d  %s" % (name[:46], total, "  ".join(items)))


for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        total += 1
        base = op.split(".")[0]
        if base in ("MUFU", "UTCHMMA", "UTCQMMA", "UTCBAR", "UBLKCP", "LDTM", "STTM"):
            counts[".".join(op.split(".")[:3])] += 1
        else:
            counts[base] += 1
flush()
