"""Turn ncu outputs (gpurun_out/) into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches_c1_cartpole.txt
    python profiles/summarize.py full gpurun_out/prof.ncu-rep      > profiles/r1_full_c1_cartpole.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
    "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_xu.sum",
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14])
    total = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {len(rows)} launches, {total / 1e3:.1f} us total")
    print(f"# (cold-cache, serialised: compare SHARES, not absolutes)")
    print(f"{'share':>7} {'count':>6} {'avg_us':>9}  kernel")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * ns / total:6.2f}% {cnt:6d} {ns / cnt / 1e3:9.2f}  {name}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print(f"## {d[hdr.index('Kernel Name')][:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:70s} {d[i]:>16s} {units[i]}")
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        print(f"{'dram traffic per launch (read + write)':70s} {float(d[i_r]) + float(d[i_w]):16.3f} {units[i_r]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        iex = hdr.index("Instructions Executed")
        body = []
        for r in rows[2:]:
            if r and r[0] == "Kernel Name":
                break
            body.append(r)
        tot = sum(int(r[iex]) for r in body if len(r) > iex and r[iex].isdigit())
        mix = Counter()
        for r in body:
            if len(r) > iex and r[iex].isdigit() and int(r[iex]) > 0:
                toks = r[1].strip().split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                mix[op.split(".")[0]] += int(r[iex])
        print(f"## SASS executed (first kernel): {tot} warp-instructions")
        for k, v in mix.most_common(18):
            print(f"{k:12s} {100 * v / tot:6.2f}%")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
