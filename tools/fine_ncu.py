"""profiles/fine_ncu.json from an `ncu --set full` capture of k_fine: DRAM bytes and warp instructions per launch, tagged with
the hash of the k_fine.cu they were measured on and the workload, so that bench.py quotes them only for THIS build.
usage: python tools/fine_ncu.py gpurun_out/prof_fine_X.ncu-rep "paris-like-30k 4096x4096 MSAA16" """
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(k):
    v, u = d[k]
    return float(v.replace(",", "")) * mult.get(u, 1)


out = {"k_fine_sha16": hashlib.sha256(open(os.path.join(ROOT, "vello_b200", "csrc", "k_fine.cu"), "rb").read()).hexdigest()[:16],
       "workload": workload, "kernel": d["Kernel Name"][0],
       "dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
       "dram_read": int(num("dram__bytes_read.sum")), "dram_write": int(num("dram__bytes_write.sum")),
       "warp_instructions": int(num("smsp__inst_executed.sum")), "duration_us_under_ncu": num("gpu__time_duration.sum") / 1e3 if d["gpu__time_duration.sum"][1] == "ns" else float(d["gpu__time_duration.sum"][0].replace(",", "")),
       "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
       "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
       "lanes_per_instruction": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
       "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "fine_ncu.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
