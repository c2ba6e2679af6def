import sys, torch
sys.path.insert(0, '/root/repo')
from bevgen_b200 import ops
d = 1024
B, Bp = 16, 16
for n_out, K in [(4096, 1024), (3072, 1024), (1024, 4096)]:
    w = torch.randn(n_out, K, device="cuda") * 0.02
    w_hi, w_lo = ops.split_planes(w, 3)
    x = torch.randn(Bp, K, device="cuda")
    x_hi, x_lo = ops.split_planes(x, 3)
    for kk in (256, 512, 1024):
        if K % kk: continue
        ks = K // kk
        part = torch.zeros(ks, Bp, n_out, device="cuda")
        f = lambda: ops.gemm_tc(a_hi=w_hi, a_lo=w_lo, a_dims=(1, 1, n_out, K), b_hi=x_hi, b_lo=x_lo, k=kk, n_cols=B, a_c_zstride=kk,
                                b_k_zstride=kk, z_inner=ks, out_w=n_out, out_zi_stride=Bp * n_out, ldc=n_out, out_f32=part,
                                flags=ops.GF_OUT_T, bn=16, npass=3)
        for _ in range(3): f()
        torch.cuda.synchronize()
        # flush L2 between launches by touching a big buffer
        big = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
        ts = []
        for _ in range(10):
            big.fill_(1.0)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) * 1e3)
        ts.sort()
        ref = (x.double() @ w.double().t())
        err = (part.sum(0).double() - ref).abs().max().item()
        print(f"n_out={n_out} K={K} kk={kk} ks={ks} ctas={ks * n_out // 128}: median {ts[5]:.1f} us  min {ts[0]:.1f} us  err {err:.1e}")
