"""Probe: cuBLAS DGEMM / TF32 / BF16 throughput on this GPU (roofline denominators that
MEASURED_PEAKS.json does not carry).  Not part of the product path."""
import torch, json, time
def bench(dtype, n, reps=10, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * n**3 / best * 1e-9
out = {}
for n in (4096, 8192):
    out[f"fp64_{n}"] = bench(torch.float64, n)
out["fp32_8192"] = bench(torch.float32, 8192)
out["tf32_8192"] = bench(torch.float32, 8192, tf32=True)
out["bf16_8192"] = bench(torch.bfloat16, 8192)
# skinny shape like the QP iteration: (B x n) @ (n x n), n = 4480
for B in (1024, 4096, 16384):
    a = torch.randn(B, 4480, device="cuda", dtype=torch.float64); m = torch.randn(4480, 4480, device="cuda", dtype=torch.float64)
    for _ in range(3): a @ m
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ m; e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    out[f"fp64_qp_B{B}"] = 2 * B * 4480**2 / best * 1e-9
print(json.dumps(out))
