"""One full-size stage-2 teacher-forced forward (B=16) for ncu launch lists: python tools/forward_profile.py [f16f8|fp32x3|bf16]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tools.stage2_perf import KW, sizes
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth
prec = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
cfg = GPTConfig(**KW)
eng = GPTEngine(synth.gpt_state_dict(sizes(cfg), seed=2), cfg, device="cuda:0", precision=prec)
cam, bev, batch = synth.stage2_inputs(16, seed=0)
cam, bev = cam.cuda(), bev.cuda()
batch = {k: v.cuda() for k, v in batch.items()}
eng.forward(cam, bev, batch, sampling=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(cam, bev, batch, sampling=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
