"""torch.profiler breakdown of one c3 tracking iteration (ours)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import tracking_bench as tb
from hier_slam_b200.scene import CONFIGS
import diff_gaussian_rasterization as ours
cfg = CONFIGS["c2"]
tb.run(ours, cfg, 1, 5)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tb.run(ours, cfg, 1, 10)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
