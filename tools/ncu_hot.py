"""Per-SASS-instruction stall samples of an ncu report (needs -lineinfo): prints the hottest instructions and an opcode histogram."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci, si, ei = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ci]), float(r[ei]), r[si].strip(), k))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data); ex = sum(d[1] for d in data)
print(f"total samples {tot:.0f}, warp-instructions {ex:.0f}, SASS lines {len(data)}")
for s, e, src, k in sorted(data, reverse=True)[:top]:
    print(f"{s / tot:6.2%} samples  {e / ex:6.2%} exec  #{k:5d}  {src[:100]}")
ops = {}
for s, e, src, k in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    a = ops.setdefault(op, [0.0, 0.0]); a[0] += s; a[1] += e
print("--- by opcode (samples share, executed share)")
for op, (s, e) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{op:10s} {s / tot:6.2%} {e / ex:6.2%}")
