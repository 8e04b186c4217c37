"""gpurun_out/r02_* -> committed summaries under profiles/ (round 2).

    python tools/summarize_profiles_r02.py
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = os.environ.get("SLB_TAG", "r02")      # r02: first half of round 2, r02b: second half
src, dst, tag = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles"), "r02"


def launch_shares(path):
    with open(path) as fh:
        rows = [r for r in csv.reader(l for l in fh if not l.startswith("==")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        t = t / 1e3 if r[ui] == "ns" else t * 1e3 if r[ui] == "ms" else t
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    return [{"kernel": k[:110], "launches": n, "total_us": round(t, 1), "avg_us": round(t / n, 2),
             "share": round(t / total, 4)} for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]]


def ncu_summary(rep, label, extra=None, match=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rr[0], rr[1], rr[2]
    if match is not None:          # a report holding several kernels: the first row whose name matches
        ki = names.index("Kernel Name")
        vals = next(r for r in rr[2:] if match in r[ki])
    m = {h: (v, u) for h, u, v in zip(names, units, vals)}

    def num(key, default=None):
        try:
            return float(m[key][0].replace(",", ""))
        except (KeyError, ValueError):
            return default

    def to_bytes(key):
        v, u = m[key]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    stalls = {}
    for h, v in zip(names, vals):
        if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v.replace(",", ""))
            except ValueError:
                pass
    ts = sum(stalls.values()) or 1.0
    dur = num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[m["gpu__time_duration.sum"][1]]
    out = {
        "kernel": m.get("Kernel Name", (label,))[0], "what": label,
        "source": "ncu --set full --clock-control none --import-source on (cold caches, serialised)",
        "duration_ms": dur,
        "grid": m.get("launch__grid_size", ("",))[0], "block": m.get("launch__block_size", ("",))[0],
        "registers_per_thread": num("launch__registers_per_thread"),
        "dynamic_smem_kb": num("launch__shared_mem_per_block_dynamic"),
        "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
        "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
        "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
        "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_active_pct_of_elapsed": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "achieved_occupancy_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
        "warp_stall_samples_pct": {k: round(100 * v / ts, 2)
                                   for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]},
    }
    if extra:
        out.update(extra)
    return out


done = []
STAGES = TAG + "_stages3.ncu-rep"          # one capture of the three stages of the default step (r02b)
for rep, label, name, match in (
        ((TAG + "_filter_mean_kernel.ncu-rep"), "filter stage 1 (fp32 screening mean where it applies, else the fp64 mean kernel) on the C2 grid", (TAG + "_filter_mean_kernel_ncu.json"), "filter_mean"),
        ((TAG + "_filter_head_kernel.ncu-rep"), "filter stage 2 (head-subset variance bound, fp64 means of the entries the screened box leaves open) on list A of the C2 grid", (TAG + "_filter_head_kernel_ncu.json"), "filter_head"),
        ((TAG + "_gp_tile_kernel.ncu-rep"), "refine pass (row- and factor-split 32-point tiles) on list B of the C2 grid", (TAG + "_refine_tile_kernel_ncu.json"), "gp_tile_kernel"),
        ((TAG + "_gp_tile_full.ncu-rep"), "full posterior for every point of the C2 grid (filter off): gp_tile_kernel<3, 64>", (TAG + "_gp_tile_kernel_ncu.json"), None)):
    path = os.path.join(src, rep)
    if not os.path.exists(path) and match is not None and os.path.exists(os.path.join(src, STAGES)):
        path = os.path.join(src, STAGES)
    else:
        match = None
    if not os.path.exists(path):
        print("missing", rep)
        continue
    json.dump(ncu_summary(path, label, match=match), open(os.path.join(dst, name), "w"), indent=1)
    cmd = ["ncu", "-i", path, "--page", "details"] + (["--kernel-name", "regex:" + match] if match else [])
    det = subprocess.run(cmd, stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(dst, name.replace("_ncu.json", "_ncu_details.txt")), "w").write(det)
    done.append(name)

if os.path.exists(os.path.join(src, (TAG + "_launches.csv"))):
    json.dump(launch_shares(os.path.join(src, (TAG + "_launches.csv"))),
              open(os.path.join(dst, (TAG + "_launch_shares.json")), "w"), indent=1)
for name in os.listdir(src):
    if name.startswith(TAG + "_") and (name.endswith(".json") or name.endswith(".jsonl") or name.endswith("launches.csv")
                                    or name.endswith("racecheck.txt")) and "call" not in name:
        shutil.copy(os.path.join(src, name), os.path.join(dst, name))
print("summaries:", done)
for name in ((TAG + "_bench.json"), (TAG + "_bench_n8.json")):
    p = os.path.join(dst, name)
    if os.path.exists(p):
        b = json.loads(open(p).read().strip().splitlines()[-1])
        print(name, "value %.4g pts/s | %.4f ms/step | e2e %.4g | roofline %.3f (%s)"
              % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel"][:30]))
