"""Roofline denominators measured on THIS box, in the same lease as the kernels they are compared with.

  gemm_peaks(device)  dense GEMM throughput of the vendor library through torch (cuBLASLt): bf16 (torch.matmul), int8
                      (torch._int_mm) and block-scaled fp4 (torch._scaled_mm on float4_e2m1fn_x2 with e8m0 scales per
                      32 elements = the tcgen05 kind::mxf4 path rf_f4x2_kernel uses; nvfp4 with e4m3 scales per 16 as
                      a second candidate).  "burst" = best of 10 single GEMMs (a kernel timed alone, e.g. the 0.6 ms
                      cfg2 contraction), "sustained" = back-to-back GEMMs for ~2 s (a seconds-long kernel, cfg5),
                      the protocol MEASURED_PEAKS.json used for bf16.
  pipe_peaks()        issue rates of the CUDA-core pipes (tools/microbench.cu): DADD/DMUL for the bit-exact ESS kernel
                      (separate multiply and add, so the ceiling is the FP64 issue rate, half the DFMA rating), POPC
                      for the popcount kernel.

Used by bench.py (rank 0, before the timed workload); `python tools/peaks.py` prints the JSON.  Library GEMMs are
measurement references here, never on the product path.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time_gemm(torch, fn, flops, sustain_s=2.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(sustain_s * 1e3 / best))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return {"burst": flops / (best * 1e-3) / 1e12, "sustained": flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
            "burst_ms": best, "sustained_reps": reps}


def gemm_peaks(device="cuda:0", n=8192, sustain_s=2.0):
    import torch
    res = {"n": n, "torch": torch.__version__, "unit": "TFLOP/s (2*N^3 per GEMM)"}
    flops = 2.0 * n ** 3
    with torch.cuda.device(device):
        try:
            a = torch.randn(n, n, device=device, dtype=torch.bfloat16)
            b = torch.randn(n, n, device=device, dtype=torch.bfloat16)
            res["bf16"] = _time_gemm(torch, lambda: a @ b, flops, sustain_s)
            del a, b
        except Exception as e:  # noqa: BLE001
            res["bf16_error"] = repr(e)[:900]
        try:
            a = torch.randint(0, 2, (n, n), device=device, dtype=torch.int8)
            b = torch.randint(0, 2, (n, n), device=device, dtype=torch.int8).t().contiguous().t()
            res["int8"] = _time_gemm(torch, lambda: torch._int_mm(a, b), flops, sustain_s)
            del a, b
        except Exception as e:  # noqa: BLE001
            res["int8_error"] = repr(e)[:900]
        # block-scaled fp4: packed e2m1 pairs, A [M][K/2] row-major, B [K/2][N] column-major, unit scales.  Two sizes
        # (the library's fp4 GEMM is still climbing at 8192^3); the best of them is the denominator.
        for name, sdt, blk, sval in (("fp4_mx_e8m0_block32", "float8_e8m0fnu", 32, 127), ("fp4_nv_e4m3_block16", "float8_e4m3fn", 16, 0x38)):
            for m in (n, 2 * n):
                try:
                    f4 = torch.float4_e2m1fn_x2
                    a = torch.randint(0, 256, (m, m // 2), device=device, dtype=torch.uint8).view(f4)
                    b = torch.randint(0, 256, (m, m // 2), device=device, dtype=torch.uint8).view(f4).t()
                    rup = lambda x, k: (x + k - 1) // k * k  # noqa: E731
                    numel = rup(m, 128) * rup(m // blk, 4)
                    sa = torch.full((numel,), sval, device=device, dtype=torch.uint8).view(getattr(torch, sdt))
                    sb = torch.full((numel,), sval, device=device, dtype=torch.uint8).view(getattr(torch, sdt))
                    fn = lambda: torch._scaled_mm(a, b, sa, sb, out_dtype=torch.bfloat16)  # noqa: E731
                    fn()
                    r = _time_gemm(torch, fn, 2.0 * m ** 3, sustain_s)
                    r["n"] = m
                    res.setdefault(name + "_by_n", {})[str(m)] = {k: r[k] for k in ("burst", "sustained")}
                    if name not in res or r["burst"] > res[name]["burst"]:
                        keep = dict(r)
                        if name in res:
                            keep["sustained"] = max(r["sustained"], res[name]["sustained"])
                        res[name] = keep
                    elif r["sustained"] > res[name]["sustained"]:
                        res[name]["sustained"] = r["sustained"]
                    del a, b, sa, sb
                except Exception as e:  # noqa: BLE001
                    res[name + "_error"] = repr(e)[:900]
                    break
        torch.cuda.empty_cache()
    return res


def pipe_peaks(device_index=0):
    exe = os.path.join(ROOT, "tools", "microbench")
    if not os.path.exists(exe):
        return {"error": "tools/microbench not built (__graft_entry__.build() compiles it)"}
    try:
        env = dict(os.environ)  # the binary measures device 0 of what it sees: show it this process's GPU only
        vis = [v for v in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if v]
        env["CUDA_VISIBLE_DEVICES"] = vis[device_index] if device_index < len(vis) else str(device_index)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
        d = json.loads(r.stdout)
        keep = {k: d[k] for k in ("gpu", "sms", "clock_khz", "popc", "dadd", "dmul", "dfma", "dmul_dadd_pair_instrs") if k in d}
        keep["fp64_issue_tflops"] = d["dmul_dadd_pair_instrs"]["ops_per_s"] / 1e12  # one DMUL or DADD per issue slot = one flop
        return keep
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:900]}


if __name__ == "__main__":
    out = {"gemm": gemm_peaks(), "pipes": pipe_peaks()}
    json.dump(out, sys.stdout, indent=1)
    print()
