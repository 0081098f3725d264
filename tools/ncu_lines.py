"""Per-source-line summary of an ncu report: samples and instructions executed per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8 or r[0] == "":
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        lines.append((int(d["# Samples"]), int(d["Instructions Executed"]), int(r[0]), r[1].strip()[:110]))
    except ValueError:
        pass
ts = sum(l[0] for l in lines) or 1
ti = sum(l[1] for l in lines) or 1
print(f"total samples {ts}, total warp instructions {ti}")
for s_, i_, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*s_/ts:5.1f}% smp {100*i_/ti:5.1f}% inst  L{ln:<4} {src}")
