"""One MaskGit forward at the reference config's size for ncu launch lists: python tools/maskgit_profile.py [B]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.maskgit_engine import MaskGitEngine
from oracle import synth
from tests.cases import GPT_KW, gpt_sizes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kw = {**GPT_KW, "num_layers": 14, "sparse_block_size": 1, "cam_latent_res": (14, 25), "cam_res": (224, 400)}
cfg = GPTConfig(**kw)
sd = synth.maskgit_state_dict(gpt_sizes(cfg), 14, 16, seed=1)
critic = {"weight": sd.pop("to_pred.weight"), "bias": sd.pop("to_pred.bias")}
eng = MaskGitEngine(sd, cfg, depth=14, heads=16, device="cuda:0", critic=critic, precision=sys.argv[2] if len(sys.argv) > 2 else "f16f8")
cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=0)
ids, bev = cam.reshape(B * cfg.num_cams, -1).cuda(), bev.cuda()
batch = {k: v.cuda() for k, v in batch.items()}
eng.forward(ids, bev, batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(ids, bev, batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
