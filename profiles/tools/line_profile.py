"""Per-source-line executed-instruction profile of one kernel.

Joins the SASS page of an ncu report (instructions executed per SASS instruction) with
`nvdisasm -g` line information of the same kernel in the built library.

    python profiles/tools/line_profile.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def ncu_sass(report):
    if report.endswith(".csv"):      # page exported on the GPU box: ncu -i rep --page source --csv > file
        raw = open(report).read()
    else:
        raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    i_src, i_inst, i_thr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    out = []
    for r in rows[2:]:
        if len(r) <= i_thr or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break
            continue
        out.append((r[i_src].strip(), int(r[i_inst] or 0), int(r[i_thr] or 0)))
    return out


def disasm_lines(cubin, kernel):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith(".text.") and kernel in l and l.endswith(":"):
            start = i
            break
    assert start is not None, "kernel not found"
    cur = None
    out = []
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((m.group(2).strip(), cur))
    return out


def main():
    report, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    a = ncu_sass(report)
    b = disasm_lines(cubin, kernel)
    print(f"# ncu sass rows {len(a)}, nvdisasm instructions {len(b)}")
    n = min(len(a), len(b))
    per = defaultdict(lambda: [0, 0])
    total = 0
    for k in range(n):
        src, inst, thr = a[k]
        line = b[k][1]
        key = (line[0], line[1]) if line else ("?", 0)
        per[key][0] += inst
        per[key][1] += 1
        total += inst
    print(f"# total warp-instructions {total}")
    for key, (inst, cnt) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * inst / total:6.2f}%  {inst:12d}  sass={cnt:4d}  {key[0]}:{key[1]}")


if __name__ == "__main__":
    main()
