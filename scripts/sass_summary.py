"""SASS summary of liblitridge.so (no GPU needed): per kernel the instruction count and the counts of the mnemonics
that prove what the kernel runs on (tcgen05 MMAs, TMEM loads, TMA loads, mbarriers, shared / global vector accesses,
local-memory spills).

    python scripts/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "litcoder_core_b200", "liblitridge.so")

GROUPS = collections.OrderedDict([
    ("tcgen05 MMA", r"^UTC[A-Z]*MMA"),
    ("TMEM ld", r"^LDTM"),
    ("TMEM alloc", r"^UTCATOM|^UTCALLOC|^UTCRELINQ"),
    ("commit / tc barrier", r"^UTCBAR"),
    ("TMA load", r"^UTMALDG"),
    ("TMA prefetch / store", r"^UTMAPF|^UTMASTG"),
    ("mbarrier", r"^SYNCS"),
    ("LDS/STS", r"^LDS|^STS"),
    ("LDG/STG .128", r"^(LDG|STG)\S*\.128"),
    ("LDG/STG other", r"^(LDG|STG)"),
    ("L2 prefetch", r"^CCTL|^PREFETCH|^LDG\S*\.LTC"),
    ("local (spill)", r"^LDL|^STL"),
    ("FP64", r"^D(ADD|MUL|FMA|SETP)"),
    ("ATOM/RED", r"^ATOM|^RED"),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for g, pat in GROUPS.items():
                if re.match(pat, op):
                    cur[g] += 1
                    break
    names = demangle(list(kernels))
    print("# SASS summary of liblitridge.so (sm_100a; `python scripts/sass_summary.py`, cuobjdump %s)\n" %
          subprocess.run(["cuobjdump", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1])
    print("Counts are static instructions per kernel (all template instances listed).  `UTC*MMA` = tcgen05.mma "
          "(`.2CTA` forms in the cta_group::2 instances), `LDTM` = tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor loads, "
          "`SYNCS` = mbarrier operations, `UTCBAR` = tcgen05.commit.  There is no `UTMASTG`: the store epilogue goes "
          "through shared memory with ordinary 16-byte stores (DESIGN.md 5).\n")
    cols = list(GROUPS)
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["total"]):
        name = names.get(k, k)
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        print("| `%s` | %d | %s |" % (name[:100], c["total"], " | ".join(str(c[g]) if c[g] else "" for g in cols)))


if __name__ == "__main__":
    main()
