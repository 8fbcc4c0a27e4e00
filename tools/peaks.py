"""Measured int8 / bf16 GEMM throughput on this box (cuBLASLt through torch), the same burst protocol
MEASURED_PEAKS.json used for bf16: the denominator candidates for the tcgen05 int8 kernel."""
import json
import torch

dev = "cuda:0"
res = {}
N = 8192
for name, mk, mm, flops in [
    ("bf16", lambda: (torch.randn(N, N, device=dev, dtype=torch.bfloat16), torch.randn(N, N, device=dev, dtype=torch.bfloat16)),
     lambda a, b: a @ b, 2.0 * N ** 3),
    ("int8", lambda: (torch.randint(0, 2, (N, N), device=dev, dtype=torch.int8), torch.randint(0, 2, (N, N), device=dev, dtype=torch.int8).t().contiguous().t()),
     lambda a, b: torch._int_mm(a, b), 2.0 * N ** 3),
]:
    try:
        a, b = mk()
        for _ in range(3):
            mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); mm(a, b); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name + "_tflops_burst"] = flops / (best * 1e-3) / 1e12
    except Exception as e:  # noqa
        res[name + "_error"] = repr(e)[:200]
print(json.dumps(res))
