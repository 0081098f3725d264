"""profiles/traffic.json from an ncu --set full report: DRAM bytes (read + write) per launch of each blend kernel.
usage: python tools/ncu_traffic.py report.ncu-rep config_key"""
import csv, io, json, os, subprocess, sys
rep, key = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = "blend_fwd" if "forward" in d["Kernel Name"] else "blend_bwd" if "backward" in d["Kernel Name"] else None
    if name is None:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[m]) * scale[units[hdr.index(m)]]
    res[name] = tot
p = os.path.join(ROOT, "profiles", "traffic.json")
allr = json.load(open(p)) if os.path.exists(p) else {}
allr[key] = res
allr["_source"] = "ncu --set full --clock-control none, one launch per kernel, bytes = dram__bytes_read.sum + dram__bytes_write.sum"
json.dump(allr, open(p, "w"), indent=1)
print(json.dumps(allr))
