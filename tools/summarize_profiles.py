"""Dev tool: turn the ncu captures under gpurun_out/ into the small text summaries committed under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

def launch_list(path, dst, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    t = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0][:90]
        t[name][0] += 1
        t[name][1] += float(r[iv].replace(",", ""))
    unit = rows[1][hdr.index("Metric Unit")]
    tot = sum(v[1] for v in t.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n# source: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# launches profiled: {sum(v[0] for v in t.values())}, total {tot:.0f} {unit}\n")
        f.write(f"{'share':>7} {'launches':>8} {'total':>12} {'avg':>10}  kernel\n")
        for name, (n, s) in sorted(t.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100*s/tot:6.2f}% {n:8d} {s:12.0f} {s/n:10.1f}  {name}\n")
    print(open(dst).read())

def full_summary(rep, dst, title, extra=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    with open(dst, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}), one row per profiled launch\n{extra}")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                vals = [r[i] for r in rows[2:] if len(r) > i]
                f.write(f"{w} [{units[i]}]: {' | '.join(v[:70] for v in vals)}\n")
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[2:]:
            if len(r) > i_w:
                sc = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                f.write(f"traffic (dram read+write) per launch: {(float(r[i_r]) * sc[units[i_r]] + float(r[i_w]) * sc[units[i_w]]) / 1e6:.1f} MB\n")
    print(open(dst).read())

def traffic_of(rep):
    """mean dram read+write bytes per profiled launch of a --set full capture"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    hdr, units = rows[0], rows[1]
    i_r, i_w, i_t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    sc = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    tot = [float(r[i_r]) * sc[units[i_r]] + float(r[i_w]) * sc[units[i_w]] for r in rows[2:] if len(r) > i_w]
    dur = [float(r[i_t]) for r in rows[2:] if len(r) > i_t]
    return {"dram_bytes_per_launch": sum(tot) / len(tot), "launches_profiled": len(tot),
            "ncu_duration_" + units[i_t]: sum(dur) / len(dur), "source": os.path.basename(rep)}

g = os.path.join(ROOT, "gpurun_out")
launch_list(os.path.join(g, f"launches_default_{tag}.csv"), os.path.join(OUT, f"{tag}_launches_bench_default.txt"),
            "bench.py --steps 1 --warmup 3 (default net, 4096 games x 100 sims): per-kernel device time")
full_summary(os.path.join(g, f"tree_default_{tag}.ncu-rep"), os.path.join(OUT, f"{tag}_ncu_tree_wave_kernel.txt"),
             "tree_wave_kernel inside bench.py (default net, 4096 games x 100 sims/move)")
full_summary(os.path.join(g, f"tree_stub_{tag}.ncu-rep"), os.path.join(OUT, f"{tag}_ncu_tree_wave_kernel_stub.txt"),
             "tree_wave_kernel inside bench.py --net stub (tree + rules only, 4096 games x 100 sims/move)")
full_summary(os.path.join(g, f"rules_{tag}.ncu-rep"), os.path.join(OUT, f"{tag}_ncu_rules_mask_step_kernel.txt"),
             "rules_mask_step_kernel, bench.py --workload rules (4 Mi states per launch)")
big = os.path.join(g, f"tree_stub16k_{tag}.ncu-rep")
if os.path.exists(big):
    full_summary(big, os.path.join(OUT, f"{tag}_ncu_tree_wave_kernel_stub16k.txt"),
                 "tree_wave_kernel inside bench.py --net stub --games 16384 --nodes 640 (4x the headline's games per launch)")
traffic = {"tree_wave_kernel": {"default": traffic_of(os.path.join(g, f"tree_default_{tag}.ncu-rep")),
                                "stub": traffic_of(os.path.join(g, f"tree_stub_{tag}.ncu-rep")),
                                **({"stub16k": traffic_of(big)} if os.path.exists(big) else {})},
           "rules_mask_step_kernel": {"rules": traffic_of(os.path.join(g, f"rules_{tag}.ncu-rep"))}}
json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
rep = os.path.join(g, f"net_convs_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    with open(os.path.join(OUT, f"{tag}_ncu_net_convolutions.txt"), "w") as f:
        f.write("# the policy/value net's 512-channel convolutions inside bench.py (cuDNN/CUTLASS library kernels, fp16 in / fp32 accumulate)\n"
                "# source: ncu --set full --clock-control none (net_convs_{tag}.ncu-rep), one column per profiled launch\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w} [{units[i]}]: {' | '.join(r[i][:60] for r in rows[2:] if len(r) > i)}\n")
    print(open(os.path.join(OUT, f"{tag}_ncu_net_convolutions.txt")).read())
for name in ("bench_default", "bench_stub", "bench_rules", "bench_reference", "bench_sims200"):
    src = os.path.join(g, f"{tag}_{name}.json")
    if os.path.exists(src):
        open(os.path.join(OUT, f"{tag}_{name}.json"), "w").write(open(src).read())
for name in ("trace_default", "trace_stub"):
    src = os.path.join(g, f"{tag}_{name}.log")
    if os.path.exists(src):
        open(os.path.join(OUT, f"{tag}_{name}.txt"), "w").write("# tools/trace_wave.py 4096 " + name.split("_")[1] + ": per-warp %globaltimer phases of the wave kernel\n" + open(src).read())
