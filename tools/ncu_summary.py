"""Summarise an .ncu-rep (first kernel) into a small text/JSON: python tools/ncu_summary.py rep.ncu-rep out.txt"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
res = {}
for i, h in enumerate(hdr):
    if h in keep or h in stall:
        res[h] = (vals[i], units[i])
with open(out, "w") as f:
    f.write(f"# ncu summary of {rep}\n")
    for k in keep:
        if k in res:
            f.write(f"{k:75s} {res[k][0]} {res[k][1]}\n")
    f.write("# stall reasons (% of active warps), top 8\n")
    st = sorted(((float(res[h][0].replace(',', '')), h) for h in stall if h in res), reverse=True)[:8]
    for v, h in st:
        f.write(f"{h:75s} {v:.2f}\n")
    try:
        rd = float(res["dram__bytes_read.sum"][0].replace(',', '')); wr = float(res["dram__bytes_write.sum"][0].replace(',', ''))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * mult[res["dram__bytes_read.sum"][1]] + wr * mult[res["dram__bytes_write.sum"][1]]
        f.write(f"dram_bytes_per_launch {tot:.0f}\n")
    except Exception as e:
        f.write(f"# dram total unavailable: {e}\n")
print(open(out).read())
