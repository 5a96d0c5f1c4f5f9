"""Copy the artefacts of tools/gpu_deliver.sh from gpurun_out/ into profiles/ and derive the summaries (developer tool)."""
import collections, csv, json, os, shutil, subprocess
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
for src, dst in [("bench.json", "r02_bench_fused_bf16_B256.json"), ("bench_ref.json", "r02_bench_reference_arm.json"),
                 ("bench_b1024.json", "r02_bench_fused_bf16_B1024.json"), ("bench_fp16.json", "r02_bench_fused_fp16_B256.json"),
                 ("timeline.txt", "r02_stage_timeline_fused_bf16_B256.txt"), ("launches.csv", "r02_launches_fused_bf16_B256.csv"),
                 ("layerwise_b4096.txt", "r02_event_profile_layerwise_bf16_B4096.txt"),
                 ("bench_cfg3.json", "r02_bench_fused_bf16_B256_cfg3.json"), ("bench_c5_1gpu.json", "r02_bench_c5_stl_sd_euler100_B4096_1gpu.json"),
                 ("timeline_B1024.txt", "r02_stage_timeline_fused_bf16_B1024.txt")]:
    if os.path.isfile(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
raw = subprocess.run(["ncu", "-i", os.path.join(G, "r02_full_forward.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_dim_x", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
cols = [c for c in want if c in h]
with open(os.path.join(P, "r02_ncu_full_one_forward_fused_bf16_B256.csv"), "w") as f:
    w = csv.writer(f); w.writerow(cols); w.writerow([units[h.index(c)] for c in cols])
    for r in rows[2:]:
        w.writerow([r[h.index(c)] for c in cols])
ix = {c: h.index(c) for c in cols}
mult = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
agg = {}
for r in rows[2:]:
    k = "k_chain" if "k_chain" in r[ix["Kernel Name"]] else "k_attn"
    a = agg.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "tensor_pipe_active_pct_time_weighted": 0.0})
    t = float(r[ix["gpu__time_duration.sum"]])
    a["launches"] += 1; a["us"] += t
    a["dram_bytes"] += float(r[ix["dram__bytes_read.sum"]]) * mult[units[ix["dram__bytes_read.sum"]]] + float(r[ix["dram__bytes_write.sum"]]) * mult[units[ix["dram__bytes_write.sum"]]]
    a["tensor_pipe_active_pct_time_weighted"] += float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) * t
for a in agg.values():
    a["dram_bytes_per_launch"] = a["dram_bytes"] / a["launches"]; a["tensor_pipe_active_pct_time_weighted"] /= a["us"]
json.dump({"source": "ncu --set full --clock-control none, the 19 kernels of one forward of the bench workload (B=256 bf16 fused): "
                     "profiles/r01_ncu_full_one_forward_fused_bf16_B256.csv; dram__bytes_read.sum + dram__bytes_write.sum "
                     "(activation writes stay in the 126 MB L2)", "kernels": agg}, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
print(json.dumps(agg, indent=1))
lr = list(csv.reader(open(os.path.join(P, "r02_launches_fused_bf16_B256.csv"))))
hi = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
hh = lr[hi]; jx = {n: i for i, n in enumerate(hh)}
tt = collections.defaultdict(float); cnt = collections.Counter()
for r in lr[hi + 1:]:
    if len(r) == len(hh) and r[jx["Metric Name"]] == "gpu__time_duration.sum":
        k = r[jx["Kernel Name"]][:32]; tt[k] += float(r[jx["Metric Value"]].replace(",", "")); cnt[k] += 1
tot = sum(tt.values())
for k in tt:
    print(cnt[k], k, round(tt[k] / 1e6, 2), "ms", round(100 * tt[k] / tot, 1), "%")
for f in ("r02_bench_fused_bf16_B256.json", "r02_bench_fused_bf16_B1024.json", "r02_bench_fused_fp16_B256.json", "r02_bench_reference_arm.json"):
    d = json.loads(open(os.path.join(P, f)).read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), (round(d["roofline"]["achieved"], 2), round(d["roofline"]["frac"], 4)) if "roofline" in d else "", d.get("cpu_baseline", {}).get("value"))

# the standalone GroupNorm+FiLM+SiLU pass: ncu --set full of the first six GN launches of a layer-wise forward at B=4096
gn_rep = os.path.join(G, "r02_gn_tma_b4096.ncu-rep")
if os.path.isfile(gn_rep):
    raw = subprocess.run(["ncu", "-i", gn_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    cols = [c for c in want if c in h]
    with open(os.path.join(P, "r02_ncu_gn_pass_B4096.csv"), "w") as f:
        w = csv.writer(f); w.writerow(cols); w.writerow([units[h.index(c)] for c in cols])
        for r in rows[2:]:
            w.writerow([r[h.index(c)] for c in cols])
    print("gn pass:", [(r[h.index("gpu__time_duration.sum")], r[h.index("dram__bytes_read.sum")], r[h.index("dram__bytes_write.sum")]) for r in rows[2:]])
