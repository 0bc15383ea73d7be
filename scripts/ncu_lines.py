"""Attribute the stall samples of an ncu report to CUDA source lines.

    python scripts/ncu_lines.py gpurun_out/X.ncu-rep <kernel substring> [cubin/object with -lineinfo [mangled substring]]

ncu's `--page source --csv` lists SASS instructions with their sampling counts but without source lines; nvdisasm
--print-line-info lists the same SASS with `//## File "...", line N` markers.  The two listings are joined by
instruction index inside the kernel (same binary).  Prints the hottest source lines with their dominant stall reasons.
"""
import collections
import csv
import re
import subprocess
import sys


def sass_lines(obj, kernel_sub):
    """[(line_no, sass_text)] of the first function whose name contains kernel_sub."""
    out = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout
    cur_fn, cur_line, res, active = None, None, [], False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            cur_fn = m.group(1)
            if active and res:
                break
            active = kernel_sub in cur_fn
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((cur_line, m.group(2).strip()))
    return res


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    obj = sys.argv[3] if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    # the report may hold several kernels: take the block whose "Kernel Name" row matches
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and ksub in r[1]:
            start = i
            break
    if start is None:
        raise SystemExit("kernel not found in the report")
    hdr = rows[start + 1]
    ix = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            data.append(r)
    stall_cols = [k for k in hdr if k.startswith("stall_") and "(" not in k]
    samples = [int(r[ix["# Samples"]]) for r in data]
    tot = sum(samples)
    print("instructions %d, samples %d" % (len(data), tot))
    if obj is None:
        return
    sl = sass_lines(obj, sys.argv[4] if len(sys.argv) > 4 else ksub)
    print("nvdisasm instructions %d" % len(sl))
    if len(sl) != len(data):
        print("WARNING: instruction counts differ; joining by index anyway")
    per_line = collections.defaultdict(lambda: [0, collections.Counter(), 0])
    for i, r in enumerate(data):
        line = sl[i][0] if i < len(sl) else None
        e = per_line[line]
        e[0] += samples[i]
        e[2] += int(r[ix["Instructions Executed"]])
        for k in stall_cols:
            v = int(r[ix[k]])
            if v:
                e[1][k] += v
    src = {}
    for (line, (s, st, ex)) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:40]:
        text = ""
        if line:
            f, n = line
            if f not in src:
                try:
                    import glob
                    cand = glob.glob("autompc_b200/csrc/" + f)
                    src[f] = open(cand[0]).read().splitlines() if cand else []
                except OSError:
                    src[f] = []
            if 0 < n <= len(src[f]):
                text = src[f][n - 1].strip()[:90]
        print("%5.1f%%  %-22s exec=%-9d %-90s %s" % (100.0 * s / max(tot, 1), "%s:%d" % line if line else "?", ex, text,
                                                   dict(st.most_common(2))))


if __name__ == "__main__":
    main()
