#!/usr/bin/env python
"""What a library GEMM makes of the 728-wide trunk shape (M = 32 x 32 x 32 pixels, N = K = 728, FP16 in / FP32 accumulate):
the yardstick for the hand-written pointwise kernel's 42 us.  Measurement aid only -- nothing in the product calls a library GEMM."""
import torch

def main():
    M, N, K = 32 * 32 * 32, 728, 728
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    b = torch.randn(K, N, device="cuda", dtype=torch.float16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for label, fn in (("matmul [M,K]x[K,N]", lambda: a @ b), ("matmul [M,K]x[N,K]^T", lambda: a @ b.t().contiguous().t())):
        for _ in range(20): fn()
        ts = []
        for _ in range(30):
            flush.zero_()
            t0.record(); fn(); t1.record(); torch.cuda.synchronize()
            ts.append(t0.elapsed_time(t1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        print(f"{label}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
    # back to back (inputs L2-warm, launch overlap), 50 calls
    for _ in range(5): a @ b
    t0.record()
    for _ in range(50): a @ b
    t1.record(); torch.cuda.synchronize()
    print(f"50 back to back: {t0.elapsed_time(t1) * 1e3 / 50:.1f} us each")

if __name__ == "__main__":
    main()
