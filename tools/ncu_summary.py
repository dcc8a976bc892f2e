"""Selected raw metrics of every kernel in one or more .ncu-rep files -> JSON (what profiles/*_ncu_summary.json holds).
usage: python tools/ncu_summary.py out.json rep1.ncu-rep [rep2.ncu-rep ...]   (developer tool)"""
import csv, io, json, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "").strip()
        k = {m: {"unit": u[m], "value": d[m]} for m in WANT if m in d}
        k["stall_cycles_per_issue"] = {m: d[m] for m in hdr if "issue_stalled" in m and m.endswith("per_issue_active.ratio") and float(d[m] or 0) >= 0.05}
        k["source"] = rep.split("/")[-1]
        out[name] = k
json.dump(out, open(sys.argv[1], "w"), indent=1)
for n, k in out.items():
    print(n, k["gpu__time_duration.sum"]["value"], k["gpu__time_duration.sum"]["unit"], "dram r/w", k["dram__bytes_read.sum"]["value"], k["dram__bytes_write.sum"]["value"], k["dram__bytes_read.sum"]["unit"])
