"""SASS instruction summary of the shipped library's kernels (no GPU needed):
   python scripts/sass_summary.py autompc_b200/lib/libampc_b200.so profiles/r02_sass_summary.md
Counts the Blackwell-specific mnemonics that prove the tcgen05 / TMEM / bulk-copy path per kernel instantiation."""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "DMMA",
        "DFMA", "FFMA", "F2FP", "LDS", "STS", "LDG", "STG", "BAR", "MUFU"]


def clean(name):
    name = re.sub(r"\([^()]*\)$", "", name.strip())                 # argument list
    name = re.sub(r"\((?:int|bool)\)", "", name)                    # (int)2 -> 2
    return name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")


def main():
    lib, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", ln)
        if m and fn:
            per[fn][m.group(1)] += 1
            per[fn]["_total"] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTCBAR", "UBLKCP", "SYNCS", "F2FP"):
                per[fn]["_" + full] += 1
    demangle = subprocess.run(["cu++filt"] + list(per.keys()), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS instruction summary of %s" % lib, "",
             "`cuobjdump -sass`, static instruction counts per kernel (sm_100a).  UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = "
             "tcgen05.ld / .st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops.", "",
             "| kernel | total | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for (fn, c), name in zip(per.items(), demangle):
        short = clean(name)
        lines.append("| `%s` | %d | " % (short[:90], c["_total"]) + " | ".join(str(c[k]) for k in KEYS) + " |")
    lines += ["", "## variants of the Blackwell-specific instructions in the headline instantiation", ""]
    for (fn, c), name in zip(per.items(), demangle):
        if clean(name).endswith("mppi_rollout_tc_kernel<2, 24, 1, 1, 0>"):
            lines.append("`%s`:" % clean(name))
            for k, v in sorted(c.items()):
                if k.startswith("_") and k != "_total":
                    lines.append("* %s: %d" % (k[1:], v))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, "(%d kernels)" % len(per))


if __name__ == "__main__":
    main()
