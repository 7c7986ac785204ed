"""Markdown table of the multi-GPU bench lines (gpurun_out/mg2_*.json, mg8_*_nN.json, f_bench_*.json) for DESIGN.md §7 / profiles/."""
import glob, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
def add(name, n, path):
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        rows.setdefault(name, {})[n] = d
    except Exception:
        pass
G = os.path.join(ROOT, "gpurun_out")
for f in glob.glob(os.path.join(G, "mg8_*_n*.json")):
    m = re.match(r"mg8_(.*)_n(\d+)\.json", os.path.basename(f)); add(m.group(1), int(m.group(2)), f)
for f in glob.glob(os.path.join(G, "mg2_*.json")):
    add(os.path.basename(f)[4:-5], 2, f)
for name, f in (("saw_svf", "f_bench_saw_svf.json"), ("subtractive_weak", "f_bench_subtractive.json"), ("subtractive_strong", "f_bench_subtractive.json")):
    add(name, 1, os.path.join(G, f))
for f in glob.glob(os.path.join(G, "final_bench_*.json")):
    nm = os.path.basename(f)[12:-5]
    for name in ((nm,) if nm not in ("subtractive", "net") else (nm + "_weak", nm + "_strong")):
        add(name, 1, f)
print("| configuration | N | voices / GPU | value (Gsample/s) | e2e (Gsample/s) | ms per step (1 s of audio) | x N=1 |")
print("|---|---|---|---|---|---|---|")
for name in sorted(rows):
    base = rows[name].get(1)
    for n in sorted(rows[name]):
        d = rows[name][n]
        sp = f"{d['value'] / base['value']:.2f}" if base else "—"
        print(f"| {name.replace('_', ' ')} | {n} | {d['config']['voices_per_gpu']} | {d['value'] / 1e3:.1f} | {d['e2e']['value'] / 1e3:.1f} | {d['ms_per_step']:.2f} | {sp} |")
