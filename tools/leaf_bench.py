"""Leaf cross-entropy (1x1 conv S -> L + CE, scripts/hierslam.py:975-984): the tcgen05 / TMEM / TMA pixel pass
(csrc/leaf_loss_tc.cu) vs the mma.sync generation (csrc/leaf_loss.cu), with and without the weight gradient, CUDA events.
One JSON line per shape; `tensor_frac` = 3xTF32 algorithmic flops / time / measured TF32 peak (tools/tf32_peak.py)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hier_slam_b200 import _lib, losses
lib = _lib.load()
TF32_PEAK = float(os.environ.get("HS_TF32_PEAK_TFLOPS", "717.7"))
g = torch.Generator().manual_seed(0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (S, L, hh, ww) in ((74, 550, 480, 640), (26, 102, 680, 1200), (16, 41, 480, 640)):
    sem = torch.randn(S, hh, ww, generator=g).cuda()
    w = (0.3 * torch.randn(L, S, generator=g)).cuda(); b = torch.randn(L, generator=g).cuda()
    lab = torch.randint(0, L, (hh, ww), generator=g).int().cuda()
    loss = torch.zeros((), device="cuda"); grad = torch.empty_like(sem)
    row = dict(shape=[S, hh, ww], classes=L)
    for legacy in (False, True):
        losses.LEAF_KERNEL = "mma_sync" if legacy else "tcgen05"
        for wg in (False, True):
            fn = lambda: losses._run_leaf(lib, sem, lab, w, b, 1.0, hh * ww, loss, grad, False, wg)
            row[("mma_sync" if legacy else "tcgen05") + ("_with_wgrad_ms" if wg else "_pixel_ms")] = round(t(fn), 4)
    losses.LEAF_KERNEL = "auto"
    K = (S + 1 + 15) // 16 * 16
    chunks = (L + 63) // 64
    flops = 2.0 * hh * ww * K * 64 * chunks * 2          # logits + dX, one product each (3xTF32 counts once)
    row["pixel_pass_gflop"] = round(flops / 1e9, 2)
    row["tensor_frac_of_measured_tf32_peak"] = round(flops / (row["tcgen05_pixel_ms"] * 1e-3) / 1e12 / TF32_PEAK, 4)
    row["tensor_frac_counting_3_products"] = round(3 * flops / (row["tcgen05_pixel_ms"] * 1e-3) / 1e12 / TF32_PEAK, 4)
    print(json.dumps(row), flush=True)
