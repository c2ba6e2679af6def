"""Persistent decode kernel on the full-size model: failure diagnostics + the kernel's own per-phase profile.
python tools/decode_debug.py <case> [steps]   cases: plain | forced | trace | forced_trace | topk0"""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from bevgen_b200.gpt_config import GPTConfig
from bevgen_b200.gpt_decode import GPTSampler
from bevgen_b200.gpt_engine import GPTEngine
from oracle import synth
from tests.cases import GPT_FULL, gpt_sizes


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "plain"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    B = 16
    cfg = GPTConfig(**{**GPT_FULL, "num_layers": layers})
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="f16f8")
    cam, bev, batch = synth.stage2_inputs(B, seed=0)
    bd = {k: v.cuda() for k, v in batch.items()}
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx] if "forced" in case else None
    smp = GPTSampler(eng, B)
    smp.profile_phases = os.environ.get("DP_PROFILE", "1") == "1"
    kw = dict(temperature=1.0, top_k=None if (case == "topk0" or "forced" in case) else 100, seed=3, forced_tokens=forced, steps=steps)
    res = {"case": case, "steps": steps, "layers": layers}
    try:
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            out = smp.sample(bev, bd, trace_logits=("trace" in case), **kw)
            torch.cuda.synchronize()
            res[f"run{it}_s"] = time.time() - t0
        res["profile_ms"] = smp.last_profile()
        if os.environ.get("DP_DUMP") == "1" and smp._pk and "prof" in smp._pk:
            t = smp._pk["prof"][:-32].double().cpu() / 1e3 / ((steps - 1) * layers)       # us per layer-iteration, [cta][slot]
            names = smp.PROFILE_SLOTS
            q = torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0], dtype=torch.double)
            res["per_cta_us_per_layer(min,p10,p50,p90,max)"] = {names[i]: [round(float(v), 2) for v in torch.quantile(t[:, i], q)] for i in range(8)}
            body = t[:, 0] + t[:, 2] + t[:, 4] + t[:, 6]
            order = torch.argsort(body, descending=True)
            res["slowest_ctas(body us/layer)"] = [(int(i), round(float(body[i]), 1)) for i in order[:10]]
            res["fastest_ctas"] = [(int(i), round(float(body[i]), 1)) for i in order[-6:]]
        res["ms_per_token"] = res["run1_s"] * 1e3 / max(steps - 1, 1)
        res["ok"] = True
    except Exception as ex:
        res["ok"] = False
        res["error"] = repr(ex)[:200]
        res["failure(code,cta,step,layer,phase,a,b,thread)"] = smp.last_failure()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
