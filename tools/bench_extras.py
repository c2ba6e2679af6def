"""Sub-objects of the bench line (N = 1 only): the stage-1 round trip (BASELINE configs[1]) and the teacher-forced forward (configs[2])
with their own rooflines, a same-box eager-PyTorch GPU baseline, and the MaskGit variant.  Everything here runs AFTER the headline's timed
regions; it re-uses the model bench.py built."""
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from bevgen_b200 import ops  # noqa: E402


def _timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def _launch_times(fn):
    """CUDA-event time of every tcgen05 GEMM-family / attention launch of one call of fn (events on the launching stream)."""
    pairs = []

    def timer(kind, launch, flops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch()
        b.record()
        pairs.append((kind, a, b, flops))
    ops.Stats.timer = timer
    try:
        fn()
    finally:
        ops.Stats.timer = None
    torch.cuda.synchronize()
    return [(k, a.elapsed_time(b), f) for k, a, b, f in pairs]


def vqgan_round_trip(model, batch_dev, precision, sustained):
    """BASELINE configs[1]: RGB VQGAN encode -> quantise -> decode of 96 images 256x256 (VQModel.encode / decode)."""
    fs = model.first_stage_model
    x = model.get_input("image", batch_dev)
    n = x.shape[0]

    def step():
        quant, _, (_, _, idx) = fs.encode(x, None)
        return fs.decode(quant)
    step()
    ms = _timed(step, 5)
    lt = _launch_times(step)
    groups = {}
    for kind, t, f in lt:
        g = groups.setdefault((kind, round(f)), [0.0, 0])
        g[0] += t
        g[1] += 1
    (dkind, dflops), (dms, dn) = max(groups.items(), key=lambda kv: kv[1][0])
    achieved = dflops / (dms / dn / 1e3) / 1e12
    mult = {"fp32x3": 3, "f16f8": 2, "bf16": 1}[precision]
    fam_ms, fam_fl = sum(t for _, t, _ in lt), sum(f for _, _, f in lt)
    return {"images_per_s": n / ms * 1e3, "ms_per_batch": ms, "images": n,
            "roofline": {"bound": "tensor", "kernel": f"{dkind} (dominant launch shape: {dflops / 1e12:.3f} TFLOP algorithmic, {dn} launches/step, {dms / dn:.3f} ms each)",
                         "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                         "executed_mma_multiplier": mult, "family_ms_per_step": fam_ms, "family_achieved_tflops": fam_fl / fam_ms / 1e9,
                         "algorithmic_gflop_per_image": fam_fl / n / 1e9}}


def gpt_forward(model, precision, sustained, B=16):
    """BASELINE configs[2]: teacher-forced forward of the 24-layer GPT on B samples; attention TFLOP/s on allowed positions and dense-equivalent."""
    from oracle import synth
    eng = model.transformer.engine()
    cam, bev, mats = synth.stage2_inputs(B, seed=0)
    cam, bev = cam.cuda(), bev.cuda()
    mats = {k: v.cuda() for k, v in mats.items()}
    fwd = lambda: eng.forward(cam, bev, mats, sampling=True)
    fwd()
    ms = _timed(fwd, 3)
    lt = _launch_times(fwd)
    att = [t for k, t, _ in lt if k == "attn_fused"]
    att_ms = statistics.mean(att) if att else float("nan")
    fl = eng.flops_per_sample() * B
    allowed = 4.0 * B * eng.d * eng._allowed
    dense = 4.0 * B * eng.L ** 2 * eng.d
    return {"samples_per_s": B / ms * 1e3, "ms_per_batch": ms, "batch": B, "algorithmic_tflops": fl / ms / 1e9, "frac_of_sustained_bf16_peak": fl / ms / 1e9 / sustained,
            "attention": {"ms_per_layer": att_ms, "tflops_allowed_only": allowed / att_ms / 1e9, "tflops_dense_equiv": dense / att_ms / 1e9,
                          "roofline": {"bound": "tensor", "kernel": "attn_fused_kernel", "achieved": allowed / att_ms / 1e9, "peak": sustained, "unit": "TFLOP/s",
                                       "frac": allowed / att_ms / 1e9 / sustained}},
            "precision": precision}


def torch_gpu_baseline(model, batch_dev, loop_steps=3, kv_steps=48):
    """Same-box GPU reference leg (SURVEY 8d last row): the reference's eager PyTorch path (oracle restatement = the reference's modules,
    dense fp32 stand-in for the DeepSpeed ops) on THIS B200 in fp32, TF32 off and on, template scripts/inference.py:106-121.  The
    reference algorithm's generate (one full forward per token) is timed for `loop_steps` iterations at B = 16 and extrapolated; an eager
    KV-cache loop is timed for `kv_steps` tokens as the 'what plain PyTorch with a cache would give' row."""
    import numpy as np
    from bevgen_b200.gpt_config import GPTConfig  # noqa: F401
    from oracle import gpt_oracle, synth, vqgan_oracle
    out = {}
    g = np.load(ROOT / "tests" / "golden" / "vqgan_config2_rgb.npz")
    gl = np.load(ROOT / "tests" / "golden" / "gpt_full24.npz")
    cfg = model.cfg
    sd_v = {k: v.detach() for k, v in model.first_stage_model.state_dict().items()}
    sd_t = {k: v.detach() for k, v in model.transformer.state_dict().items() if "master_layout" not in k}
    geo = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in gpt_oracle.geo_from_config(cfg).items()}
    x = model.get_input("image", batch_dev)
    cam, bev, mats = synth.stage2_inputs(16, seed=0)
    cam, bev = cam.cuda(), bev.cuda()
    mats = {k: v.cuda() for k, v in mats.items()}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.no_grad(), torch.device("cuda"):
            for name, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32

                def rt():
                    quant, idx, h = vqgan_oracle.encode(x, sd_v)
                    return vqgan_oracle.decode(quant, sd_v), idx, h
                rec, idx, h = rt()
                ms_v = _timed(rt, 2)
                fwd = lambda: gpt_oracle.forward(sd_t, geo, cam, bev, mats, sampling=True)
                logits = fwd()
                ms_f = _timed(fwd, 2)
                row = {"vqgan_round_trip_images_per_s": x.shape[0] / ms_v * 1e3, "vqgan_ms_per_96_images": ms_v,
                       "vqgan_latent_max_err_vs_reference_golden": float(np.abs(h[:8, ::8].cpu().numpy() - g["h_sub"]).max()),
                       "vqgan_token_mismatches_vs_reference_golden": int((idx[:2048].cpu().numpy() != g["idx"]).sum()),
                       "gpt_forward_ms_B16": ms_f, "gpt_forward_samples_per_s": 16 / ms_f * 1e3,
                       "gpt_logits_max_err_vs_reference_golden": float(np.abs(logits[:2, gl["rows"]].cpu().numpy() - gl["logits_s"]).max())}
                # reference algorithm: full forward per generated token (B = 16 scenes at once), extrapolated
                xt = torch.full((16, cfg.num_cams, cfg.num_cam_tokens), cfg.vocab_size, dtype=torch.int64)

                def loop_iter(t=[0]):
                    j = int(cfg.forward_shuffle_idx[t[0]])
                    i, k = j // cfg.num_cam_tokens, j % cfg.num_cam_tokens
                    lg = gpt_oracle.forward(sd_t, geo, xt, bev, mats, sampling=True).view(16, cfg.num_cams, cfg.num_cam_tokens, -1)[:, i, k]
                    xt[:, i, k] = torch.multinomial(gpt_oracle.sample_probs(lg, 1.0, 100), 1)[:, 0]
                    t[0] += 1
                loop_iter()
                ms_it = _timed(loop_iter, loop_steps)
                dec96 = ms_v * 0.65                      # decode share of the round trip (252.7 of 391.3 GFLOP/image)
                gen_s = (ms_it * 1536 + ms_v + 2 * dec96 + ms_f) / 1e3
                row.update(reference_algorithm_generate_images_per_s_extrapolated=96 / gen_s, reference_algorithm_ms_per_token_step_B16=ms_it)
                out[name] = row
            # eager KV-cache loop (fp32, TF32 off): kv_steps tokens at B = 16
            torch.backends.cuda.matmul.allow_tf32 = False
            dec = gpt_oracle.KVCacheDecoder(sd_t, geo)
            g2 = torch.Generator(device="cuda").manual_seed(0)
            chooser = lambda p, t: torch.multinomial(p, 1, generator=g2)[:, 0]
            dec.run(bev, mats, steps=4, top_k=100, chooser=chooser)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dec.run(bev, mats, steps=kv_steps, top_k=100, chooser=chooser)
            b.record()
            torch.cuda.synchronize()
            out["eager_kv_cache_fp32"] = {"ms_per_token_step_B16_first_tokens": a.elapsed_time(b) / kv_steps,
                                          "note": f"prefill + first {kv_steps} tokens (short caches: a lower bound on the per-token cost of the full 1536)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["note"] = ("oracle restatement of the reference modules run with torch eager ops (cuDNN / cuBLAS) on the same B200; errors are against the "
                   "reference-minted goldens, i.e. what the reference's own path delivers at that math mode")
    return out


def run(model, batch_dev, precision, sustained, hbm):
    res = {}
    for key, fn in (("vqgan_configs1", lambda: vqgan_round_trip(model, batch_dev, precision, sustained)),
                    ("forward_configs2", lambda: gpt_forward(model, precision, sustained)),
                    ("torch_gpu_baseline", lambda: torch_gpu_baseline(model, batch_dev))):
        try:
            res[key] = fn()
        except Exception as ex:          # the headline line must still be printed
            res[key] = {"error": repr(ex)}
        torch.cuda.empty_cache()
    try:
        from tools.maskgit_perf import run as maskgit_run
        res["maskgit"] = maskgit_run(8, precision)
    except Exception as ex:
        res["maskgit"] = {"error": repr(ex)}
    return res
