"""Per-source-line summary of an ncu report: stall samples and warp instructions executed per CUDA source line
(SASS rows are attributed to the innermost file:line ncu maps them to).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, os, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, fname = None, "?"
acc = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8 or r[0] == "":
        continue   # SASS rows (already summed into their source row by ncu)
    d = dict(zip(hdr[4:], r[4:]))
    try:
        key = (fname, int(r[0]))
        s_, i_ = int(d["# Samples"]), int(d["Instructions Executed"])
    except ValueError:
        continue
    o = acc.setdefault(key, [0, 0, r[1].strip()[:100]])
    o[0] += s_; o[1] += i_
ts = sum(v[0] for v in acc.values()) or 1
ti = sum(v[1] for v in acc.values()) or 1
print(f"total samples {ts}, total warp instructions {ti}")
for (fn, ln), (s_, i_, src) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*i_/ti:5.1f}% inst {100*s_/ts:5.1f}% smp  {fn}:{ln:<4} {src}")
