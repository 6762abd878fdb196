"""profiles/ncu_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two kernels bench.py reports a
roofline for, READ FROM the `ncu --set full` captures under gpurun_out/ (tools/gpu_r2_final.sh makes them at the bench shapes).  bench.py
loads this file at run time; a shape without a capture reports `traffic: null`.
    python tools/ncu_traffic.py gpurun_out/r14_att64_b16.ncu-rep:attention_d64:16x4429x24 gpurun_out/r14_gemm_mlp_up.ncu-rep:gemm_mlp_up:65536x6144x1536"""
import json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(out_path)) if os.path.exists(out_path) else {}
for spec in sys.argv[1:]:
    rep, kernel, shape = spec.split(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw"], capture_output=True, text=True).stdout
    vals = {}
    for line in raw.splitlines():
        m = re.match(r"\s*(dram__bytes_(?:read|write)\.sum|gpu__time_duration\.sum)\s+(\S+)\s+([\d.,]+)", line)
        if m:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[m.group(2)]
            vals[m.group(1)] = float(m.group(3).replace(",", "")) * mult
    table.setdefault(kernel, {})[shape] = {"dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
                                           "dram_read": vals["dram__bytes_read.sum"], "dram_write": vals["dram__bytes_write.sum"],
                                           "ncu_duration_ms": vals.get("gpu__time_duration.sum"), "capture": os.path.basename(rep)}
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(table, indent=1))
