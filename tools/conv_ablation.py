"""Ablation timing of the fused conv kernel on the dominant VQGAN layer (128->128 3x3 at 256x256, 96 images)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from bevgen_b200 import ops

N, H, W, C = 96, 256, 256, 128
dev = "cuda"
x = torch.randn(N, H, W, C, device=dev)
res = torch.randn(N, H, W, C, device=dev)
w = torch.randn(9 * C + 128, C, device=dev) / (9 * C) ** 0.5
w_hi, w_lo = ops.split_planes(w, 3)
b = torch.randn(C, device=dev)
aff = torch.randn(N, C, 2, device=dev)
out = torch.empty(N, H, W, C, device=dev)
sums = torch.empty(N * 64, dtype=torch.float64, device=dev)
flops = 2.0 * N * H * W * C * C * 9

import os
w16, w8pair, lo_scale = ops.pack_f16f8(w)

def t(label, npass=3, two_cta=True, dbg=0, **kw):
    args = dict(affine=aff, swish=True, residual=res, gn_sums=sums)
    args.update(kw)
    os.environ["BEVGEN_CONV_DBG"] = str(dbg)
    if npass == 2:
        f = lambda: ops.conv3x3_fused_f16f8(x, w16, w8pair, lo_scale, C, b, out, **args)
    else:
        f = lambda: ops.conv3x3_fused(x, w_hi, w_lo if npass == 3 else None, C, b, out, npass=npass, two_cta=two_cta, **args)
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    print(f"{label:44s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s algorithmic")

t("full (affine+swish, residual, stats) 2cta")
t("full, 1cta", two_cta=False)
t("no affine/swish (identity prologue)", affine=None, swish=False)
t("no residual", residual=None)
t("no stats", gn_sums=None)
t("bare (identity prologue, no residual, no stats)", affine=None, swish=False, residual=None, gn_sums=None)
t("bf16 full", npass=1)
t("bf16 bare", npass=1, affine=None, swish=False, residual=None, gn_sums=None)

for npass in (2, 3, 1):
    print(f"--- kernel-internal ablations (BEVGEN_CONV_DBG), npass={npass}; results are wrong by construction")
    t("full", npass=npass)
    t("1: producers skip the global fetch", npass=npass, dbg=1)
    t("2: producers skip transform + smem stores", npass=npass, dbg=2)
    t("3: producers do nothing", npass=npass, dbg=3)
    t("4: epilogue drains TMEM only", npass=npass, dbg=4)
    t("8: no MMA issue", npass=npass, dbg=8)
    t("7: producers + epilogue idle (MMA + weights)", npass=npass, dbg=7)
    t("12: no MMA, epilogue idle (producers + weights)", npass=npass, dbg=12)
    t("11: no MMA, producers idle (epilogue + weights)", npass=npass, dbg=11)
    t("15: weights ring only", npass=npass, dbg=15)
os.environ["BEVGEN_CONV_DBG"] = "0"
