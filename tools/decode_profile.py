"""Short decode run for ncu launch lists: full-size model, B=16, a few eager steps at a late position (long KV)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tools.stage2_perf import KW, sizes
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_decode import GPTSampler
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
start = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
cfg = GPTConfig(**KW)
eng = GPTEngine(synth.gpt_state_dict(sizes(cfg), seed=2), cfg, device="cuda:0", precision=prec)
_, bev, batch = synth.stage2_inputs(16, seed=0)
s = GPTSampler(eng, 16)
s.sample(bev, batch, top_k=100, seed=1, steps=3, use_graph=False)
# jump to a late step (cache contents beyond the prefill are zeros: timing only)
bev = bev.cuda(); batch = {k: v.cuda() for k, v in batch.items()}
args = s._embed_args(bev, batch)
s.step.fill_(start)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    s._step(args, 1.0, 100, False, 1, None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", int(s.step.item()))
