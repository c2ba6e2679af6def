"""Times the two direct image-end convs at the benchmark shape (96 x 256 x 256)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from bevgen_b200 import ops
N, H, W, C = 96, 256, 256, 128
x = torch.randn(N, H, W, C, device="cuda")
img = torch.randn(N, 3, H, W, device="cuda")
w_out = torch.randn(3, C, 3, 3, device="cuda") * 0.03
w_in = torch.randn(C, 3, 3, 3, device="cuda") * 0.2
b3, bC = torch.randn(3, device="cuda"), torch.randn(C, device="cuda")
aff = torch.randn(N, C, 2, device="cuda")
out = torch.empty(N, 3, H, W, device="cuda")
hout = torch.empty(N, H, W, C, device="cuda")
sums = torch.empty(N * 64, dtype=torch.float64, device="cuda")
def t(f, label, gb):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    print(f"{label:12s} {ms:7.3f} ms   {gb / ms:7.1f} GB/s algorithmic")
t(lambda: ops.conv_out3(x, w_out, b3, out, affine=aff, swish=True), "conv_out3", (x.numel() + out.numel()) * 4 / 1e6)
t(lambda: ops.conv_in3(img, w_in, bC, hout, gn_sums=sums), "conv_in3", (img.numel() + hout.numel()) * 4 / 1e6)
