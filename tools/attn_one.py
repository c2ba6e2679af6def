"""A few fused-attention launches at the benchmark geometry (B=16, 16 heads, L=1792): the target of `ncu --set full` captures.
usage: python tools/attn_one.py [fp32x3|bf16] [launches]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tools.stage2_perf import KW, sizes
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = 16
cfg = GPTConfig(**{**KW, "num_layers": 1})
eng = GPTEngine(synth.gpt_state_dict(sizes(cfg), seed=2), cfg, device="cuda:0", precision=prec)
d = cfg.num_embed
qkv = (torch.randn(B, cfg.gpt_block_size, 3 * d, device="cuda").bfloat16(), torch.randn(B, cfg.gpt_block_size, 3 * d, device="cuda").bfloat16() * 1e-2)
if eng.npass == 1:
    qkv = (qkv[0], None)
y = torch.randn(B, cfg.gpt_block_size, d, device="cuda")
for _ in range(reps):
    eng.attention(qkv, y, B, cfg.gpt_block_size)
torch.cuda.synchronize()
print("done", prec, reps)
