"""Registers / spills / shared memory of every kernel of a .cu file (ptxas -v, sm_100a), one line per kernel.
    python tools/ptxas_report.py hier_slam_b200/csrc/blend_fwd.cu [substring]"""
import re, subprocess, sys
src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
       "-Xptxas", "-v", "-c", src, "-o", "/dev/null"]
err = subprocess.run(cmd, capture_output=True, text=True).stderr
err = subprocess.run(["c++filt"], input=err, capture_output=True, text=True).stdout
name = None
for ln in err.splitlines():
    m = re.search(r"Compiling entry function '(.*)' for", ln)
    if m:
        name = re.sub(r"\(.*", "", m.group(1)).replace("void ", "")
        spill = ""
    elif "spill" in ln and name:
        s = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        spill = f"stack {s.group(1)} spill {s.group(2)}/{s.group(3)}" if s else ln.strip()
    elif "Used" in ln and name:
        u = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", ln)
        if flt in name:
            print(f"{name:70s} regs {u.group(1):>3s}  smem {u.group(2) or 0:>6}  {spill}")
        name = None
