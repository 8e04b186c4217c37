"""Turn gpurun_out/ artefacts of a measurement run into the committed summaries under profiles/.

    python tools/summarize_profiles.py [round_tag]      # default r01
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")

# launch list -> shares
rows = [r for r in csv.reader(open(os.path.join(src, tag + "_launches.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
total = sum(a[1] for a in agg.values())
top = [{"kernel": k[:100], "launches": n, "total_us": round(t / 1e3, 1), "share": round(t / total, 4)}
       for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]]

# full capture -> key metrics
raw = subprocess.run(["ncu", "-i", os.path.join(src, tag + "_gp_tile.ncu-rep"), "--page", "raw", "--csv"],
                     stdout=subprocess.PIPE, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
names, units, vals = rr[0], rr[1], rr[2]
m = {h: (v, u) for h, u, v in zip(names, units, vals)}


def num(key):
    return float(m[key][0].replace(",", ""))


def to_bytes(key):
    v, u = m[key]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v.replace(",", ""))
          for h, v in zip(names, vals)
          if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v.replace(",", "").replace(".", "").isdigit()}
ts = sum(stalls.values()) or 1.0
dur = num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3}[m["gpu__time_duration.sum"][1]]
summary = {
    "kernel": "gp_tile_kernel<3,false>",
    "source": "ncu --set full --clock-control none (caches flushed per pass), tools/profile_sweep.py: "
              "C2, 65536 points, M=500, 2 factors",
    "duration_ms": dur,
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
    "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "algorithmic_hbm_bytes_per_launch": 65536 * 25 + 2 * (129024 * 8 + 500 * 4 * 8),
    "registers_per_thread": int(num("launch__registers_per_thread")),
    "dynamic_smem_kb": num("launch__shared_mem_per_block_dynamic"),
    "tensor_pipe_active_pct_of_elapsed": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
    "warp_stall_samples_pct": {k: round(100 * v / ts, 2)
                               for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]},
    "launch_list_top": top,
}
json.dump(summary, open(os.path.join(dst, tag + "_gp_tile_kernel_ncu.json"), "w"), indent=1)
for name in (tag + "_bench.json", tag + "_bench_reference.json", tag + "_launches.csv",
             tag + "_bench_extra.jsonl"):
    if os.path.exists(os.path.join(src, name)):
        shutil.copy(os.path.join(src, name), os.path.join(dst, name))
b = json.load(open(os.path.join(dst, tag + "_bench.json")))
print("value %.4g pts/s | %.3f ms/step | e2e %.4g | kernel %.3f ms = %.1f%% of fp64 peak | cpu %.3g"
      % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["kernel_ms"],
         100 * b["roofline"]["frac"], b["cpu_baseline"]["value"]))
print("ncu: %.3f ms, tensor pipe %.1f%%, dram %.2f MB, kernel share %.1f%%"
      % (dur, summary["tensor_pipe_active_pct_of_elapsed"], summary["dram_bytes_per_launch"] / 1e6,
         100 * top[0]["share"]))
