"""Static evidence from the built library (no GPU needed): per kernel, the count of the SASS mnemonics that show which
hardware path it uses (B200_PROFILING.md, "What proves a Blackwell-native kernel") plus registers / shared memory from
cuobjdump's resource usage.  Usage: python tools/sass_evidence.py > profiles/<round>_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dummynode4graphlearning_b200", "csrc", "libdn4gl.so")
WATCH = [("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA\b"), ("LDTM/STTM (tcgen05.ld/st)", r"\b(LDTM|STTM)\b"),
         ("UBLKCP (cp.async.bulk)", r"\bUBLKCP\b"), ("UTMALDG/UTMASTG (TMA tensor)", r"\bUTMA(LDG|STG)\b"),
         ("SYNCS (mbarrier)", r"\bSYNCS\b"), ("FADD2/FMUL2/FFMA2 (packed fp32)", r"\b(FADD2|FMUL2|FFMA2)\b"),
         ("LDG.E.128 / STG.E.128", r"\b(LDG|STG)\.E(\.\w+)*\.128\b"), ("LDS.128", r"\bLDS(\.\w+)*\.128\b"),
         ("HMMA (legacy mma.sync)", r"\bHMMA\b"), ("ATOMG/RED float (nondeterministic adds)", r"\b(ATOMG|RED)\.E\.ADD\.F32")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in ln:
            continue
        counts[cur]["_instr"] += 1
        for label, pat in WATCH:
            if re.search(pat, ln):
                counts[cur][label] += 1
    names = demangle(list(counts))
    print("# SASS evidence for %s (sm_100a), %d kernels" % (os.path.relpath(SO, ROOT), len(counts)))
    print("# columns: instructions | registers | static shared bytes | watched mnemonics (count)")
    tot = collections.Counter()
    for k, c in sorted(counts.items(), key=lambda kv: names[kv[0]]):
        short = re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", ""))
        short = re.sub(r"^void ", "", short)
        reg, sh = usage.get(k, (-1, -1))
        marks = ", ".join("%s x%d" % (label, c[label]) for label, _ in WATCH if c[label])
        print("%-58s %6d | %3d | %6d | %s" % (short[:58], c["_instr"], reg, sh, marks or "-"))
        for label, _ in WATCH:
            if c[label]:
                tot[label] += 1
    print("\n# kernels using each path")
    for label, _ in WATCH:
        print("%-45s %d" % (label, tot[label]))


if __name__ == "__main__":
    main()
