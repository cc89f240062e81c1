"""DRAM traffic per step of every bench workload (run on the GPU box): one light ncu pass per
workload (dram__bytes_read.sum + dram__bytes_write.sum, gpu__time_duration.sum, instruction counts) of the
step / rollout kernels in steady state, written to gpurun_out/traffic_r2.json -- bench.py's
roofline.traffic and frac_physical read the committed copy, profiles/traffic.json.

    python tools/collect_traffic.py [workload ...]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402

METRICS = ("dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,"
           "launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,"
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")


def main():
    names = sys.argv[1:] or [w for w in WORKLOADS if w not in ("c2_frozenlake8", "c5_bridge_rollout8", "c5_bridge_rollout100")]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    result = {}
    for wl in names:
        rollout = "rollout" in wl
        steps, skip = (8, 4) if rollout else (70, 60)
        per_step = 2 if WORKLOADS[wl].get("hetero") else 1
        cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "3",
               "--no-cpu-baseline", "--no-table", "--e2e-steps", "2"]
        plain = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        if plain.returncode != 0:
            result[wl] = {"error": plain.stderr[-300:]}
            continue
        log = os.path.join(out_dir, f"r2_metrics_{wl}.csv")
        pat = "rollout_kernel|nsgym_spec_.*rollout" if rollout else "step_kernel|step_het_kernel|nsgym_spec_.*step"
        subprocess.run(["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", f"regex:{pat}", "-s",
                        str(skip * per_step), "-c", str(2 * per_step), "--csv", "--log-file", log] + cmd,
                       cwd=ROOT, capture_output=True, text=True)
        rows = list(csv.reader(open(log)))
        hdr, agg = None, {}
        for r in rows:
            if r and r[0] == "ID":
                hdr = r
                continue
            if hdr and len(r) == len(hdr):
                d = dict(zip(hdr, r))
                agg.setdefault(d["Metric Name"], []).append((d["Kernel Name"][:70], float(d["Metric Value"].replace(",", ""))))
        n_launch = len(agg.get("gpu__time_duration.sum", []))
        n_steps = max(n_launch // per_step, 1)
        tot = lambda k: sum(v for _, v in agg.get(k, []))  # noqa: E731
        result[wl] = {
            "traffic_bytes_per_step": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / n_steps,
            "read": tot("dram__bytes_read.sum") / n_steps, "write": tot("dram__bytes_write.sum") / n_steps,
            "ncu_us_per_step": tot("gpu__time_duration.sum") / n_steps / 1e3,
            "inst_executed_per_step": tot("smsp__inst_executed.sum") / n_steps,
            "kernels": sorted({k for k, _ in agg.get("gpu__time_duration.sum", [])}),
            "per_launch": {k: [v for _, v in vs[:per_step]] for k, vs in agg.items()
                           if k not in ("dram__bytes_read.sum", "dram__bytes_write.sum")},
            "envs": 1 << WORKLOADS[wl]["log2_envs"],
        }
        print(wl, json.dumps(result[wl])[:300], flush=True)
    with open(os.path.join(out_dir, "traffic_r2.json"), "w") as f:
        json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
