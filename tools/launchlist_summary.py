"""gpurun_out/launches_bN.csv (tools/gpu_launchlist.sh) -> the markdown table of profiles/rNN_launch_list_one_step_bN.md (last full step)."""
import collections, csv, re, sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches_b8.csv"
with open(path) as f:
    recs = list(csv.DictReader([l for l in f if not l.startswith("==")]))
names = [x["Kernel Name"] for x in recs]
fs = [i for i, n in enumerate(names) if "final_step" in n]
seg = recs[fs[-2] + 1: fs[-1] + 1]


def ms(x):
    v = float(x["Metric Value"].replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(x["Metric Unit"], 1e-6)


agg = collections.defaultdict(lambda: [0, 0.0])
for x in seg:
    n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")
    agg[(n, x["Grid Size"])][0] += 1
    agg[(n, x["Grid Size"])][1] += ms(x)
tot = sum(v[1] for v in agg.values())
print(f"{len(seg)} launches, {tot:.2f} ms summed.\n")
print("| kernel | grid | launches | ms | share | us / launch |\n|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k[0]} | {k[1]} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] * 1000:.1f} |")
by = collections.defaultdict(float)
for k, v in agg.items():
    by[k[0]] += v[1]
print("\nBy kernel: " + ", ".join(f"{k.replace('ffb::', '')} {100 * v / tot:.1f} %" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))

# kernel-family shares for bench.py's `time_share_of_step` (profiles/launch_shares.json), when asked: python tools/launchlist_summary.py x.csv --shares out.json
if "--shares" in sys.argv:
    import json
    fam = collections.defaultdict(float)
    for k, v in by.items():
        key = "attention" if "attention" in k else "gemm" if "gemm" in k else "ln_modulate" if "ln_modulate" in k else "other"
        fam[key] += v
    with open(sys.argv[sys.argv.index("--shares") + 1], "w") as f:
        json.dump({"source": "ncu launch list of one denoise step at the bench batch (tools/launchlist_summary.py; per-launch times are cold-cache and "
                             "serialised at full clocks - shares, not absolutes)", "launches_per_step": len(seg), "ms_summed": round(tot, 3),
                   **{k: f"{100 * v / tot:.1f} %" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}}, f, indent=1)
