"""Turns an .ncu-rep into the text summaries committed under profiles/ (run here, no GPU needed)."""
import csv, io, json, subprocess, sys

rep, out_prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_register_spilling"]
summary = []
with open(out_prefix + "_summary.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none, source report {rep.split('/')[-1]}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        f.write(f"\n== {name}\n")
        rec = {"kernel": name}
        for k in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                f.write(f"{k} = {v} {u}\n")
                rec[k] = v + " " + u
        stalls = []
        for i, k in enumerate(hdr):
            if "smsp__average_warps_issue_stalled_" in k and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), k.split("stalled_")[1].split("_per_")[0]))
                except ValueError:
                    pass
        f.write("warp stalls per issue: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:10]) + "\n")
        summary.append(rec)
json.dump(summary, open(out_prefix + "_summary.json", "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
top = subprocess.run([sys.executable, __file__.replace("summarize_ncu.py", "ncu_top.py"), "/tmp/_src.csv", "25"], capture_output=True, text=True).stdout
open(out_prefix + "_source_top.txt", "w").write("# ncu source page: most-sampled SASS instructions, stall mix, executed opcode mix\n" + top)
print("wrote", out_prefix + "_summary.txt")
