"""One dominant-layer conv (128->128 3x3 at 256x256, 96 images) launched a few times: the target of `ncu --set full` captures.
usage: python tools/conv_one.py [f16f8|fp32x3|bf16] [launches]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from bevgen_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, H, W, C = 96, 256, 256, 128
dev = "cuda"
x = torch.randn(N, H, W, C, device=dev)
res = torch.randn(N, H, W, C, device=dev)
w = torch.randn(9 * C + 128, C, device=dev) / (9 * C) ** 0.5
b = torch.randn(C, device=dev)
aff = torch.randn(N, C, 2, device=dev)
out = torch.empty(N, H, W, C, device=dev)
sums = torch.empty(N * 64, dtype=torch.float64, device=dev)
kw = dict(affine=aff, swish=True, residual=res, gn_sums=sums)
if mode == "f16f8b":       # 16x16-block kernel (conv_fused3.cu)
    w16, w8pair, lo_scale = ops.pack_f16f8_block(w)
    f = lambda: ops.conv3x3_fused_f16f8(x, w16, w8pair, lo_scale, C, b, out, block16=True, **kw)
elif mode == "f16f8":
    w16, w8pair, lo_scale = ops.pack_f16f8(w)
    f = lambda: ops.conv3x3_fused_f16f8(x, w16, w8pair, lo_scale, C, b, out, **kw)
else:
    npass = 3 if mode == "fp32x3" else 1
    w_hi, w_lo = ops.split_planes(w, npass)
    f = lambda: ops.conv3x3_fused(x, w_hi, w_lo, C, b, out, npass=npass, two_cta=True, **kw)
for _ in range(reps):
    f()
torch.cuda.synchronize()
print("done", mode, reps)
