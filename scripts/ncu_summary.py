"""Summarise an `ncu --set full` report of one kernel into profiles/<name>.txt and .json.

    python scripts/ncu_summary.py gpurun_out/r01_hist_full.ncu-rep profiles/hist_full "capture description"

The .json carries the per-launch averages bench.py quotes (`roofline.traffic`); the .txt is the
per-launch table a reader can check against the launch list.
"""
import csv
import json
import subprocess
import sys

rep, out, capture = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    if name not in ix:
        return None
    v = r[ix[name]].replace(",", "")
    try:
        x = float(v)
    except ValueError:
        return None
    u = units[ix[name]]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3,
             "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
    return x * scale


cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd_B"), ("dram__bytes_write.sum", "dram_wr_B"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__inst_executed_op_shared_atom.sum", "atoms_inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts")]
lines = []
tot = {c: 0.0 for _, c in cols}
n = 0
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("qr::", "")
    grid = r[ix["Grid Size"]]
    vals = {c: val(r, m) for m, c in cols}
    lines.append("%-44s grid %-14s " % (name[:44], grid) + " ".join("%s=%s" % (c, ("%.1f" % v if v is not None else "-")) for c, v in vals.items()))
    for c, v in vals.items():
        tot[c] += v or 0.0
    n += 1
avg = {c: tot[c] / max(n, 1) for c in tot}
summary = {"capture": capture, "launches": n, "us_per_launch": round(avg["us"], 2),
           "dram_bytes_per_launch": int(avg["dram_rd_B"] + avg["dram_wr_B"]),
           "dram_read_bytes_per_launch": int(avg["dram_rd_B"]), "dram_write_bytes_per_launch": int(avg["dram_wr_B"]),
           "dram_gbs_under_ncu": round((avg["dram_rd_B"] + avg["dram_wr_B"]) / max(avg["us"], 1e-9) / 1e3, 1),
           "avg": {c: round(v, 2) for c, v in avg.items()}}
with open(out + ".json", "w") as f:
    json.dump(summary, f, indent=1)
with open(out + ".txt", "w") as f:
    f.write("# %s\n# source: %s (ncu --set full --clock-control none)\n" % (capture, rep))
    f.write("\n".join(lines) + "\n")
    f.write("# averages over %d launches: %s\n" % (n, json.dumps(summary["avg"])))
print(json.dumps(summary))
