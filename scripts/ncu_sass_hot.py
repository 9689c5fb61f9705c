"""Development tool: top stall sites of one launch from `ncu --page source --print-source sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "warp instructions", tot_inst)
ops = {}
for r in data:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    a = ops.setdefault(op, [0, 0])
    a[0] += int(r[ix["Instructions Executed"]] or 0); a[1] += int(r[ix["# Samples"]] or 0)
print("by opcode (inst, samples):")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
    print("  %-10s %10d %6.1f%%  samples %6.1f%%" % (k, v[0], 100.0 * v[0] / tot_inst, 100.0 * v[1] / tot))
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i in sorted(top):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5d %5.1f%% %-60s %s" % (i, 100.0 * s / tot, r[ix["Source"]].strip()[:60], " ".join("%s=%d" % (n, v) for v, n in why if v)))
