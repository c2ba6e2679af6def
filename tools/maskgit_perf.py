"""MaskGit variant at the reference config's size (muse_stage_two_multi_view.yaml: 14 layers, d=1024, 16 heads, 6 x 14x25 latents,
18 iterations with the SelfCritic): one forward and one full generate.  python tools/maskgit_perf.py [B] [precision]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch


def run(B=8, prec="f16f8", steps=18):
    from bevgen_b200.gpt_config import GPTConfig
    from bevgen_b200.maskgit_engine import MaskGitEngine
    from oracle import synth          # synthetic weights / inputs only (no oracle arithmetic on this path)
    from tests.cases import GPT_KW, gpt_sizes
    kw = {**GPT_KW, "num_layers": 14, "sparse_block_size": 1, "cam_latent_res": (14, 25), "cam_res": (224, 400)}
    cfg = GPTConfig(**kw)
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 14, 16, seed=1)
    critic = {"weight": sd.pop("to_pred.weight"), "bias": sd.pop("to_pred.bias")}
    eng = MaskGitEngine(sd, cfg, depth=14, heads=16, device="cuda:0", precision=prec, critic=critic)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=0)
    ids = cam.reshape(B * cfg.num_cams, -1).cuda()
    bev = bev.cuda()
    batch = {k: v.cuda() for k, v in batch.items()}
    for _ in range(2):
        eng.forward(ids, bev, batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward(ids, bev, batch)
    e1.record(); torch.cuda.synchronize()
    fwd = e0.elapsed_time(e1) / 5
    gen = torch.Generator(device="cuda").manual_seed(0)
    eng.generate(bev, batch, timesteps=2, generator=gen)
    torch.cuda.synchronize()
    t0 = time.time()
    out = eng.generate(bev, batch, timesteps=steps, generator=gen)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    assert int(out.max()) < cfg.vocab_size
    return {"B": B, "precision": prec, "tokens_per_scene": cfg.num_img_tokens, "layers": 14, "forward_ms": fwd, "generate_s": gen_s,
            "images_per_s": B * cfg.num_cams / gen_s, "forwards_per_generate": 2 * steps,
            "note": "MaskGit variant (SURVEY 8f-1) at the reference config's size: 18 de-masking + 18 critic forwards; the reference runs 72 "
                    "(each twice for the eval-mode no-op guidance)"}


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    prec = sys.argv[2] if len(sys.argv) > 2 else "f16f8"
    res = run(B, prec)
    print(json.dumps(res))
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    json.dump(res, open(ROOT / "gpurun_out" / f"maskgit_perf_B{B}_{prec}.json", "w"))
