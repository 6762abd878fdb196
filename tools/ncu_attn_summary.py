"""Summarises an `ncu --set full --import-source on` capture of an attention kernel into markdown (profiles/): duration, pipe utilisation,
issue statistics, DRAM traffic, warp-stall breakdown, opcode mix and the hottest instructions.  Usage: python tools/ncu_attn_summary.py x.ncu-rep"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw"], capture_output=True, text=True).stdout
def metric(name):
    for line in raw.splitlines():
        parts = line.split()
        if parts and parts[0] == name:
            return parts[-1], (parts[-2] if len(parts) > 2 else "")
    return None, None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; st = collections.Counter(); op = collections.Counter(); ops = collections.Counter(); top = []
for k, r in enumerate(data):
    try:
        n = int(r[idx["# Samples"]]); ex = int(r[idx["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    tot += n
    for s in stalls:
        st[s[6:]] += int(r[idx[s]] or 0)
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]].strip())
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += ex; ops[o] += n
    top.append((n, r[idx["Source"]].strip(), ex))
print(f"# ncu --set full: `{kernel}` ({rep.split('/')[-1]})\n")
print("| metric | value |\n|---|---|")
for name in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
             "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "smsp__inst_executed.sum"):
    v, u = metric(name)
    if v is not None:
        print(f"| `{name}` | {v} {u} |")
print(f"\nWarp-stall samples ({tot} total):\n\n| reason | share |\n|---|---|")
for s, v in st.most_common(10):
    if v:
        print(f"| {s} | {100 * v / tot:.1f} % |")
tex = sum(op.values())
print("\nOpcode mix (executed warp instructions) and where the samples sit:\n\n| opcode | executed | share | samples |\n|---|---|---|---|")
for o, v in op.most_common(14):
    print(f"| {o} | {v / 1e6:.1f} M | {100 * v / tex:.1f} % | {100 * ops[o] / tot:.1f} % |")
print("\nHottest instructions:\n\n| samples | executed | instruction |\n|---|---|---|")
for n, s, ex in sorted(top, reverse=True)[:12]:
    print(f"| {100 * n / tot:.2f} % | {ex / 1e6:.2f} M | `{s[:90]}` |")
