"""Library GEMM peaks not in MEASURED_PEAKS.json (fp64 / tf32 / fp32), via torch.matmul (cuBLAS)."""
import json
import torch

def run(dtype, n, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device='cuda', dtype=dtype); b = torch.randn(n, n, device='cuda', dtype=dtype)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * n ** 3 / best / 1e9

out = {"fp64_tflops": run(torch.float64, 4096), "tf32_tflops": run(torch.float32, 8192, True),
       "fp32_tflops": run(torch.float32, 8192, False), "gpu": torch.cuda.get_device_name(0)}
cb = torch.randn(64, 2048, 2048, device='cuda', dtype=torch.complex64)
print(json.dumps(out))
