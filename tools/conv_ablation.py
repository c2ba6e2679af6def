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
w16b, w8pairb, lo_scaleb = ops.pack_f16f8_block(w)

def t(label, npass=3, two_cta=True, dbg=0, block16=False, **kw):
    args = dict(affine=aff, swish=True, residual=res, gn_sums=sums)
    args.update(kw)
    os.environ["BEVGEN_CONV_DBG"] = str(dbg)
    if npass == 2 and block16:
        f = lambda: ops.conv3x3_fused_f16f8(x, w16b, w8pairb, lo_scaleb, C, b, out, block16=True, **args)
    elif npass == 2:
        f = lambda: ops.conv3x3_fused_f16f8(x, w16, w8pair, lo_scale, C, b, out, **args)
    else:
        f = lambda: ops.conv3x3_fused(x, w_hi, w_lo if npass == 3 else None, C, b, out, npass=npass, two_cta=two_cta, block16=block16, **args)
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    print(f"{label:44s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s algorithmic")

if len(sys.argv) < 2: t("full (affine+swish, residual, stats) 2cta")
if len(sys.argv) < 2: t("full, 1cta", two_cta=False)
if len(sys.argv) < 2: t("no affine/swish (identity prologue)", affine=None, swish=False)
if len(sys.argv) < 2: t("no residual", residual=None)
if len(sys.argv) < 2: t("no stats", gn_sums=None)
if len(sys.argv) < 2: t("bare (identity prologue, no residual, no stats)", affine=None, swish=False, residual=None, gn_sums=None)
if len(sys.argv) < 2: t("bf16 full", npass=1)
if len(sys.argv) < 2: t("bf16 bare", npass=1, affine=None, swish=False, residual=None, gn_sums=None)

if len(sys.argv) > 1 and sys.argv[1] == "block":
    for npass in (2, 3, 1):
        t(f"npass={npass} conv_fused2 (128-pixel tiles)", npass=npass)
        t(f"npass={npass} conv_fused3 (16x16 blocks)", npass=npass, block16=True)
        for dbg, label in [(8, "no MMA"), (24, "no MMA, no weights"), (7, "MMA + weights only"), (3, "producers idle"), (4, "epilogue drains only")]:
            t(f"   block16 {dbg}: {label}", npass=npass, block16=True, dbg=dbg)
    os.environ["BEVGEN_CONV_DBG"] = "0"
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "pe":
    print("--- producer / epilogue interplay, f16f8 kernel, MMA issue off (8); 16 = no weight traffic")
    for dbg, label in [(0, "full"), (8, "no MMA"), (24, "no MMA, no weights"), (24 | 4, "no MMA/weights, epilogue idle (P only)"),
                       (24 | 3, "no MMA/weights, producers idle (E only)"), (24 | 2, "no MMA/weights, P fetch only + E"),
                       (24 | 1, "no MMA/weights, P transform only + E"), (24 | 7, "nothing (loop skeleton)"),
                       (24 | 4 | 2, "P fetch only"), (24 | 4 | 1, "P transform only")]:
        t(f"{dbg}: {label}", npass=2, dbg=dbg)
    t("24|3, no residual: E without residual", npass=2, dbg=24 | 3, residual=None)
    t("24|3, no stats: E without stats", npass=2, dbg=24 | 3, gn_sums=None)
    t("24, no residual: P + E without residual", npass=2, dbg=24, residual=None)
    t("24, no swish/affine", npass=2, dbg=24, affine=None, swish=False)
    os.environ["BEVGEN_CONV_DBG"] = "0"
    sys.exit(0)
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
