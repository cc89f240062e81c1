"""gpurun_out/traffic_r2.json (tools/collect_traffic.py, run on the GPU box) -> the committed summaries:
profiles/traffic.json (DRAM bytes per step; bench.py's roofline.traffic / frac_physical read it) and
profiles/r2_kernel_metrics.json (per-kernel metrics of the same pass).

    python tools/publish_traffic.py
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = json.load(open(os.path.join(ROOT, "gpurun_out", "traffic_r2.json")))
traffic, metrics = {}, {}
for wl, d in src.items():
    if "error" in d:
        continue
    traffic[wl] = d["traffic_bytes_per_step"]
    pl = d["per_launch"]
    rollout_k = 32 if "rollout32" in wl else 1
    metrics[wl] = {
        "kernels": d["kernels"], "envs": d["envs"],
        "dram_bytes_per_env": round(d["traffic_bytes_per_step"] / d["envs"], 2),
        "ncu_us_per_step": d["ncu_us_per_step"],
        "warp_instructions_per_warp_step": round(d["inst_executed_per_step"] / (d["envs"] / 32) / rollout_k, 1),
        "registers": pl.get("launch__registers_per_thread"),
        "issue_active_pct": pl.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": pl.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "threads_per_inst": pl.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "fp64_pipe_pct": pl.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": pl.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    }
traffic["_unit"] = ("bytes per step (dram__bytes_read.sum + dram__bytes_write.sum of the step's kernel launch(es) in "
                    "steady state, ncu metrics pass of tools/collect_traffic.py, round 2: profiles/r2_kernel_metrics.json)")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
json.dump(metrics, open(os.path.join(ROOT, "profiles", "r2_kernel_metrics.json"), "w"), indent=1)
for wl, m in metrics.items():
    print(f"{wl:28s} {m['kernels'][0][:34]:34s} {m['dram_bytes_per_env']:7.1f} B/env {m['ncu_us_per_step']:8.1f} us "
          f"{m['warp_instructions_per_warp_step']:7.1f} instr/warp-step regs {m['registers']} issue {m['issue_active_pct']} warps {m['warps_active_pct']}")
