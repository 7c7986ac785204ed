"""Turn gpurun_out/*.ncu-rep + launches_bench.csv into tracked summaries under profiles/ (run here, no GPU needed)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


traffic = {}
try:
    traffic = json.load(open(os.path.join(OUT, f"{ROUND}_traffic.json")))   # keep entries whose .ncu-rep is no longer in gpurun_out/
except Exception:
    pass
PREFIX = "full_" if ROUND == "r01" else f"{ROUND}_full_"
for rep in sorted(f for f in os.listdir(os.path.join(ROOT, "gpurun_out")) if f.startswith(PREFIX) and f.endswith(".ncu-rep")):
    m = raw(os.path.join(ROOT, "gpurun_out", rep))
    name = rep[len(PREFIX):-len(".ncu-rep")]
    for suffix in ("_r01e", "_r01"):
        if name.endswith(suffix):
            name = name[: -len(suffix)]
    lines = [f"# ncu --set full summary: {name} ({ROUND})", "", f"kernel: `{m.get('Kernel Name', ('?', ''))[0]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEEP:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    stalls = sorted(((float(v[0]), k) for k, v in m.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v[0] not in ("", "n/a")), reverse=True)[:6]
    lines += ["", "Top warp stall reasons (warps per issue-active cycle):", ""] + [f"- {k.split('issue_stalled_')[1].split('_per_issue')[0]}: {v:.3f}" for v, k in stalls]
    open(os.path.join(OUT, f"{ROUND}_ncu_{name}.md"), "w").write("\n".join(lines) + "\n")
    def tobytes(key):
        v, u = m[key]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    key = {"saw_svf_mix": "saw_svf", "noise_svf_mix": "noise_svf", "saw_svf_voices": "saw_svf+voices", "fm_mix": "fm", "subdry": "subtractive_dry", "subdry_st": "subtractive_dry",
           "fdn": "subtractive", "conv_tc": "conv", "net_mix": "net"}.get(name, name)
    traffic[key] = {"dram_bytes_per_launch": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"), "samples_per_launch": 16384,
                    "warp_inst_per_launch": float(m["smsp__inst_executed.sum"][0]) if "smsp__inst_executed.sum" in m else None,
                    "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in m else None,
                    "note": "one 16384-sample launch of the fused voice kernel (bench steps are 3 such launches)"}
json.dump(traffic, open(os.path.join(OUT, f"{ROUND}_traffic.json"), "w"), indent=1)

# launch list of the bench command: per-kernel totals and shares
src = os.path.join(ROOT, "gpurun_out", "launches_bench.csv" if ROUND == "r01" else f"{ROUND}_launches_bench.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = {}
    for r in rows[1:]:
        try:
            tot.setdefault(r[ki].split("(")[0][:120], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    allsum = sum(sum(v) for v in tot.values())
    lines = [f"# ncu launch list of `python bench.py --steps 2 --warmup 3` ({ROUND}): gpu__time_duration.sum per kernel", "",
             "(cold-cache, serialised: compare shares, not absolutes)", "", "| kernel | launches | total (ns) | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v):.0f} | {sum(v) / allsum:.1%} |")
    open(os.path.join(OUT, f"{ROUND}_launches_bench.md"), "w").write("\n".join(lines) + "\n")
print("wrote", sorted(os.listdir(OUT)))
