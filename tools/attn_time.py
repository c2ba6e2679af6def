"""Fused-attention time per layer at the benchmark geometry (B=16, 16 heads, L=1792), 50 launches: python tools/attn_time.py [fp32x3|bf16]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tools.stage2_perf import KW, sizes, timed
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
B = 16
cfg = GPTConfig(**{**KW, "num_layers": 1})
eng = GPTEngine(synth.gpt_state_dict(sizes(cfg), seed=2), cfg, device="cuda:0", precision=prec)
d = cfg.num_embed
qkv = (torch.randn(B, cfg.gpt_block_size, 3 * d, device="cuda").bfloat16(), torch.randn(B, cfg.gpt_block_size, 3 * d, device="cuda").bfloat16() * 1e-2)
if eng.npass == 1:
    qkv = (qkv[0], None)
y = torch.randn(B, cfg.gpt_block_size, d, device="cuda")
for _ in range(5):
    eng.attention(qkv, y, B, cfg.gpt_block_size)
ms = [timed(lambda: eng.attention(qkv, y, B, cfg.gpt_block_size), 50) for _ in range(3)]
print(f"attn_fused {prec}: {min(ms):.4f} ms / layer (B=16) [{', '.join(f'{m:.4f}' for m in ms)}]")
