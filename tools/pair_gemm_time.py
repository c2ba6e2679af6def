"""Time the three stage-2 linear shapes on gemm_pair (2-CTA f16f8) vs gemm_tc npass=2.  python tools/pair_gemm_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bevgen_b200 import ops

dev = "cuda"
M = 16 * 1792
for (N, K, gelu) in [(3072, 1024, False), (4096, 1024, True), (1024, 4096, False)]:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    ap = ops.pack_act_f16f8_scaled(a); wp = ops.pack_linear_f16f8(w)
    au = ops.pack_act_f16f8(a); wu = ops.pack_f16f8(w)
    out = torch.empty(M, N, device=dev)
    o16, opair = torch.empty(M, N, dtype=torch.float16, device=dev), torch.empty(M, 2 * N, dtype=torch.uint8, device=dev)
    def pair():
        if gelu: ops.linear_f16f8(ap[0], ap[1], wp[0], wp[1], wp[2], M, N, K, bias=bias, gelu=True, out_f16=o16, out_pair=opair)
        else: ops.linear_f16f8(ap[0], ap[1], wp[0], wp[1], wp[2], M, N, K, bias=bias, out_f32=out)
    def tc():
        if gelu: ops.gemm_tc(a_hi=au[0], a_lo=au[1], a_dims=(1, 1, M, K), b_hi=wu[0], b_lo=wu[1], k=K, n_cols=N, out_w=M, ldc=N, bias=bias,
                             out_hi=o16, out_lo=opair, flags=ops.GF_GELU | ops.GF_OUT_F16F8, bn=128, npass=2, lo_scale=wu[2])
        else: ops.gemm_tc(a_hi=au[0], a_lo=au[1], a_dims=(1, 1, M, K), b_hi=wu[0], b_lo=wu[1], k=K, n_cols=N, out_w=M, ldc=N, bias=bias,
                          out_f32=out, bn=128, npass=2, lo_scale=wu[2])
    for name, fn in (("pair", pair), ("gemm_tc", tc)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"N={N} K={K} gelu={gelu} {name}: {ms*1e3:.0f} us  {2.0*M*N*K/ms/1e9:.0f} TFLOP/s algorithmic", flush=True)
    if not gelu:
        pair(); o1 = out.clone(); tc(); torch.cuda.synchronize()
        print("   max |pair - gemm_tc| =", (o1 - out).abs().max().item())
