"""Measures the dense TF32 tensor throughput of this GPU the way MEASURED_PEAKS.json measures BF16: torch.matmul (cuBLAS)
on 8192^3 with allow_tf32, best of 10 (burst) and back to back for ~3 s (sustained); also fp32 SIMT for reference.
Prints one JSON line (used as the denominator of the tensor-roofline fraction of the tcgen05 / mma.sync TF32 kernels)."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k = 0
while time.perf_counter() - t0 < 3.0:
    for _ in range(10): a @ b
    k += 10
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / k
torch.backends.cuda.matmul.allow_tf32 = False
for _ in range(2): a @ b
torch.cuda.synchronize()
e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
f32 = e0.elapsed_time(e1)
fl = 2 * n ** 3
print(json.dumps({"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl / sus / 1e9, "fp32_simt_tflops": fl / f32 / 1e9,
                  "how": "torch.matmul 8192^3, allow_tf32 (cuBLAS): best of 10 / back to back for 3 s; fp32 without tf32",
                  "gpu": torch.cuda.get_device_name()}))
