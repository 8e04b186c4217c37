"""Markdown of DESIGN.md section 6.1 from the committed measurement files under profiles/.

    python tools/design_table.py > /tmp/r02_table.md
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = os.environ.get("SLB_TAG", "r02")      # r02: first half of round 2, r02b: second half
P = os.path.join(ROOT, "profiles")


def last_json(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def jsonl(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return []
    return [json.loads(l) for l in open(path) if l.startswith("{")]


out = []
b = last_json((TAG + "_bench.json"))
if b:
    r, rf, f, e = b["roofline"], b["roofline_full_posterior"], b["filter"], b["e2e"]
    st = r.get("stage_ms", {})
    out += ["| Quantity (N=1, `profiles/%s_bench.json`) | Value |" % TAG, "|---|---|",
            "| `value` (device-resident, whole default step) | **%.3g points/s** (%.4f ms/step, spread %.4f–%.4f, %d launches/step) |"
            % (b["value"], b["ms_per_step"], b["ms_per_step_spread"]["min"], b["ms_per_step_spread"]["max"], b["gpu_launches"] // b["steps"]),
            "| `e2e` (host buffers in/out every step: %.2f MB H2D GP tables, %.1f KB D2H safe set + key/stats) | **%.3g points/s** (%.4f ms/step) |"
            % (e["h2d_bytes_per_step"] / 1e6, e["d2h_bytes_per_step"] / 1e3, e["value"], e["ms_per_step"]),
            "| filter: decided by mean + prior bound / by the head-subset bound / refined by the full posterior | %.1f%% / %.1f%% / %.2f%% |"
            % (100 * f["decided_by_mean_and_prior_bound"], 100 * f["decided_by_head_rank_bound"], 100 * f["refined_by_full_posterior"]),
            "| stage times, L2 flushed (mean / head / refine incl. its launches) | %.1f / %.1f / %.1f µs |"
            % (1e3 * st.get("mean", 0), 1e3 * st.get("head", 0), 1e3 * st.get("refine", 0)),
            "| `roofline` = %s (stage 1: %s) | %.2f TFLOP/s algorithmic (F_B = D·M·(3d_in+4+E_exp), E_exp=1) = **%.1f%% of %.1f TFLOP/s** (%s); %.3g exp/s = %.0f%% of the %s |"
            % (r["kernel"].split(":")[0], r.get("stage1", "fp64 mean"), r["achieved"], 100 * r["frac"], r["peak"],
               "fp32 FFMA peak 148 x 128 x 2 x 1.965 GHz" if r.get("stage1") == "fp32 screening" else "measured fp64 peak",
               r.get("exp_per_s", 0), 100 * r.get("exp_frac", 0),
               "SFU rate (16 ex2 per clock and SM)" if r.get("stage1") == "fp32 screening" else "exp-only microbenchmark"),
            "| `roofline_full_posterior` = `gp_tile_kernel<3,64>` over the whole grid (filter off) | %.3f ms → %.1f TFLOP/s algorithmic = **%.1f%% of the fp64 DMMA peak** |"
            % (rf["kernel_ms"], rf["achieved"], 100 * rf["frac"]),
            "| the same step with the filter off (round-1 path) | %.3g points/s (%.3f ms/step) |"
            % (b["full_posterior"]["value"], b["full_posterior"]["ms_per_step"]),
            "| parity of the timed configuration (safe set, c_max vs oracle, %d points) | %d mismatches, c_max equal: %s |"
            % (b["parity"]["points"], b["parity"]["mismatches"], b["parity"]["c_max_equal"]),
            "| CPU baseline (oracle, %d BLAS threads = fastest variant) | %.3g points/s |"
            % (b["cpu_baseline"]["cores"], b["cpu_baseline"]["value"]),
            "| clocks during the timed region | %s MHz of %s, reasons %s |"
            % (b["clocks"]["sm_mhz"], b["clocks"]["sm_max_mhz"], b["clocks"]["reasons"]), ""]
ref = last_json((TAG + "_bench_reference.json"))
if ref:
    out += ["`--impl reference` arm (same oracle, %d threads, %s steps): %.3g points/s." % (
        ref["cpu_baseline"]["cores"], ref["steps"], ref["value"]), ""]
rows = []
for name, label in (((TAG + "_bench.json"), "1, weak (256² per GPU)"), ((TAG + "_bench_n2.json"), "2, weak"),
                    ((TAG + "_bench_n4.json"), "4, weak"), ((TAG + "_bench_n8.json"), "8, weak"),
                    ((TAG + "_bench_n1_strong.json"), "1, strong (one 2048² grid)"),
                    ((TAG + "_bench_n2_strong.json"), "2, strong"), ((TAG + "_bench_n8_strong.json"), "8, strong")):
    d = last_json(name)
    if d:
        par = d.get("parity")
        rows.append("| %s | %.3g | %.4f | %.3g | %.3f | %s | %s |" % (
            label, d["value"], d["ms_per_step"], d["e2e"]["value"], d["full_posterior"]["ms_per_step"],
            ("%d mismatches" % par["mismatches"]) if par else "n/a (> 2^20 points)",
            "peer memory" if "peer" in d["exchange"] else d["exchange"][:24]))
if rows:
    out += ["| GPUs, scaling | points/s (default step) | ms/step | e2e points/s | ms/step, filter off | parity vs oracle | key exchange |",
            "|---|---|---|---|---|---|---|"] + rows + [""]
ex = jsonl((TAG + "_bench_extra.jsonl"))
if ex:
    out += ["Secondary measurements (`tools/bench_extra.py`, `profiles/" + TAG + "_bench_extra.jsonl`, one B200):", ""]
    for d in ex:
        k = d["bench"]
        if k == "c5_gp_sweep":
            out.append("* C5 %s, M=%d: default step %.3f ms = %.3g points/s (refined %.2f%%); full posterior kernel %.3f ms = %.1f TF = %.0f%% of peak"
                       % (d["grid"], d["M"], d["update_safe_set_ms"], d["points_per_s"], 100 * d["refined_frac"],
                          d["full_posterior_kernel_ms"], d["full_posterior_tflops"], 100 * d["full_posterior_frac_of_fp64_peak"]))
        elif k == "deterministic_linear_sweep":
            out.append("* deterministic LinearSystem sweep %s, %s: %.3f ms = %.3g points/s = %.0f GB/s algorithmic = **%.1f%% of the HBM roofline**"
                       % (d["grid"], d["kernel"], d["kernel_ms"], d["points_per_s"], d["hbm_algorithmic_gbs"], 100 * d["hbm_frac_of_measured"]))
        elif k == "deterministic_sweep":
            out.append("* deterministic pendulum-plant sweep %s: %.3f ms = %.3g points/s (%.1f%% of HBM: ten fp64 sin per point)"
                       % (d["grid"], d["kernel_ms"], d["points_per_s"], 100 * d["hbm_frac_of_measured"]))
        elif k == "bellman_value_iteration":
            out.append("* C3 value iteration %s, M=%d: %.3f ms per sweep = %.3g state updates/s (%.3g exp/s)"
                       % (d["grid"], d["M"], d["ms_per_sweep"], d["state_updates_per_s"], d["exp_per_s"]))
        elif k == "discrete_policy_optimization":
            out.append("* discrete_policy_optimization %s, %d actions, M=%d: factored %.2f ms vs %.2f ms per-action sweeps (%.1fx), same greedy action on %.1f%% of the states"
                       % (d["grid"], d["actions"], d["M"], d["factored_ms"], d["per_action_sweeps_ms"], d["speedup"], 100 * d["same_greedy_action_frac"]))
        elif k == "c2_shared_factor":
            out.append("* C2 with one shared Cholesky factor: default step %.3f ms = %.3g points/s" % (d["update_safe_set_ms"], d["points_per_s"]))
        elif k == "c4_cartpole_update_safe_set":
            out.append("* C4 at 1-GPU size (32^4, M=2000, four factors): %.1f ms per update_safe_set = %.3g points/s" % (d["ms_per_sweep"], d["points_per_s"]))
        elif k == "notebook_pendulum_kernels":
            out.append("* the reference's 2001x1501 pendulum experiment, its own kernels, M=%d: update_safe_set %.2f ms = %.3g points/s" % (d["M"], d["update_safe_set_ms"], d["points_per_s"]))
    out.append("")
dist = jsonl((TAG + "_bench_extra_dist_n8.jsonl"))
if dist:
    out += ["Stated-size configurations on 8 B200 (`tools/bench_extra_dist.py`, `profiles/" + TAG + "_bench_extra_dist_n8.jsonl`; grid sharded by contiguous index range, peer-memory key exchange):", ""]
    for d in dist:
        out.append("* %s (N=%d points, M=%d, %d factors): **%.2f ms** per update_safe_set = %.3g points/s; filter %.1f%% / %.1f%% / %.2f%% refined"
                   % (d["bench"], d["grid_points"], d["M"], d["factors"], d["ms_per_update_safe_set"], d["points_per_s"],
                      100 * d["filter"]["prior"], 100 * d["filter"]["head"], 100 * d["filter"]["refined"]))
    out.append("")
for name in ((TAG + "_filter_mean_kernel_ncu.json"), (TAG + "_filter_head_kernel_ncu.json"), (TAG + "_refine_tile_kernel_ncu.json"), (TAG + "_gp_tile_kernel_ncu.json")):
    d = last_json(name) if False else (json.load(open(os.path.join(P, name))) if os.path.exists(os.path.join(P, name)) else None)
    if d:
        out.append("* ncu `%s`: %.1f µs, fp64 pipe %.0f%%, tensor pipe %.0f%%, issue slots %.0f%%, DRAM %.2f MB, %d registers, stalls %s"
                   % (name, 1e3 * d["duration_ms"], d["fp64_pipe_active_pct"] or 0, d["tensor_pipe_active_pct_of_elapsed"] or 0,
                      d["issue_active_pct"] or 0, d["dram_bytes_per_launch"] / 1e6, d["registers_per_thread"] or 0,
                      ", ".join("%s %.0f%%" % kv for kv in list(d["warp_stall_samples_pct"].items())[:4])))
print("\n".join(out))
