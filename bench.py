#!/usr/bin/env python
"""bench.py — multi-view images/s of BEVGen's generate.py hot path (BASELINE.json configs[3] on one B200, configs[4] on N) .

Contract (task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line from rank 0.
A step = ONE `Net2NetTransformer.test_step(batch)` call (what generate.py's trainer.test runs per batch, reference
modules/stage2/cond_transformer_multi_view.py:378-384,479-544) on 16 synthetic scenes per GPU:
    batch dict (6 x 256x256 RGB + 256x256x7 BEV map + camera matrices per scene)
      -> RGB VQGAN encode (96 images) + BEV VQGAN encode (16 maps) -> teacher-forced forward + CE loss (the test/loss it logs)
      -> reconstruction decode (96 images) -> KV-cache autoregressive sampling of 16 x 1536 tokens (top-k 100, multinomial)
      -> VQGAN decode of the 96 generated images -> de-normalise  => {'gen','rec','gt'}
  value     generated images/s with the batch dict already resident in HBM (CUDA events, barrier both sides, max over ranks)
  e2e       same through the same public call with pinned HOST tensors in and `gen` + `rec` copied back to pinned host memory
            inside the timed region (the headline against the reference arm)
  roofline  the KV-cache decode loop (97 % of the step): SURVEY 8d algorithmic bytes (556 MB of bf16-equivalent weights x 1536 steps
            + bf16-equivalent KV reads = 3.33 TB per 16 scenes) / CUDA-event time of the loop, against the measured HBM copy bandwidth
  cpu_baseline / --impl reference   the reference algorithm (one full 1792-token forward per generated token, no KV cache) through the
            CPU oracle port on the host cores, a bounded number of loop iterations extrapolated to 1536 (flagged)
Sub-objects (N = 1 only): stage-1 round trip (configs[1]) and teacher-forced forward (configs[2]) with their own rooflines, a same-box
eager-PyTorch GPU baseline, the MaskGit variant.  Weak scaling: scenes are independent, 16 per rank, weights NCCL-broadcast once.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "multi_view_images_per_sec"
UNIT = "images/s"
SCENES, CAMS, RES, TOKENS = 16, 6, 256, 1536
WORKLOAD = ("configs[3] (1 GPU) / configs[4] (N GPUs): full autoregressive generate.py hot path = Net2NetTransformer.test_step on 16 scenes "
            "per GPU: 6-cam 256x256 RGB + BEV cond -> VQGAN encodes, teacher-forced loss, KV-cache sampling of 1536 tokens/scene "
            "(24-layer d=1024 GPT, top-k 100), VQGAN decode of 96 generated + 96 reconstructed images; synthetic seeded weights and inputs")
CONFIG = {"workload": WORKLOAD, "scenes_per_gpu": SCENES, "images_per_scene": CAMS, "tokens_per_scene": TOKENS,
          "gpt": "24 layers, d=1024, 16 heads, L=1792, density 1.0", "vqgan": "ch=128, ch_mult [1,1,2,2,4], 256x256 -> 16x16 latents, 1024 codes",
          "l2": "per-step working set (1.1 GB packed weights + 2.8 GB KV cache + 3.2 GB activations) exceeds the 126 MB L2; no explicit flush"}
ALGO_BYTES_8D = 24 * 11 * 1024 * 1024 * 2 + 1024 * 1024 * 2          # 556 MB of weights per decode step (bf16, head included)


def algo_bytes_decode(scenes, steps=TOKENS, nc=256, layers=24, d=1024):
    """weights once per step + bf16 KV reads 24*2*n*d*2 B per sample with n = 257 + t."""
    kv = sum(layers * 2 * (nc + 1 + t) * d * 2 for t in range(steps)) * scenes
    return ALGO_BYTES_8D * steps + kv


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1349.2), d.get("bf16_tflops", 1643.0), d.get("hbm_gbs", 6541.8), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while running (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 - 0.1 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def gpt_kw(layers=24):
    return dict(embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, num_cams=6, vocab_size=1024, cond_vocab_size=1024, hidden_size=1024,
                num_embed=1024, num_heads=16, num_layers=layers, backend="deepspeed", sparse_block_size=16, window_len=32, cam_res=(256, 256),
                cam_latent_res=(16, 16), plot=False, causal_order=True, camera_bias=True, image_embed=True, bev_embed=True,
                bev_latent_res=(16, 16), density=1.0, cam_names="NUSCENES_CAMERAS", dataset="NUSCENES")


def gpt_sizes(cfg):
    return dict(num_embed=cfg.num_embed, gpt_block_size=cfg.gpt_block_size, num_img_tokens=cfg.num_img_tokens, num_cond_tokens=cfg.num_cond_tokens,
                num_cams=cfg.num_cams, vocab_size=cfg.vocab_size, cond_vocab_size=cfg.cond_vocab_size, num_layers=cfg.num_layers)


def host_batch(scenes, rank=0, pin=True):
    """The batch dict generate.py's dataloader would hand over (SURVEY 8b schema), seeded: scene s / camera c = image 6 s + c of
    synth.image_batch(96, seed=100 + rank) - the tensor the parity goldens (tests/golden/vqgan_config2_*.npz) were minted on."""
    import torch
    from oracle import synth          # seeded inputs only
    x = synth.image_batch(scenes * CAMS, 3, RES, RES, seed=100 + rank)
    seg = (synth.image_batch(scenes, 7, RES, RES, seed=300 + rank) > 0).float()
    _, _, mats = synth.stage2_inputs(scenes, seed=rank)
    b = {"image": x.view(scenes, CAMS, 3, RES, RES).permute(0, 1, 3, 4, 2).contiguous(), "segmentation": seg.permute(0, 2, 3, 1).contiguous(),
         "intrinsics_inv": mats["intrinsics_inv"].contiguous(), "extrinsics_inv": mats["extrinsics_inv"].contiguous()}
    if pin:
        b = {k: v.pin_memory() for k, v in b.items()}
    b["sample_token"] = [f"scene{rank}_{i}" for i in range(scenes)]
    return b


def build_model(precision, dev, load_weights=True):
    """Net2NetTransformer(GPT 24 x 1024, RGB VQModel, BEV VQSegmentationModel) under the reference's import paths, seeded weights."""
    import torch
    from multi_view_generation.modules.losses.vqperceptual import DummyLoss
    from multi_view_generation.modules.stage1.vqgan import VQModel, VQSegmentationModel
    from multi_view_generation.modules.stage2.cond_transformer_multi_view import Net2NetTransformer
    from multi_view_generation.modules.transformer.mingpt_sparse import GPT, GPTConfig
    from oracle import synth          # seeded weights only
    cfg = GPTConfig(**gpt_kw())
    gpt = GPT(cfg, precision=precision)
    dd, ddb = synth.vqgan_ddconfig(), synth.vqgan_ddconfig(in_channels=7)
    fs = VQModel(dd, DummyLoss(), 1024, 256, (RES, RES), (16, 16), 256, precision=precision)
    cs = VQSegmentationModel(7, ddb, DummyLoss(), 1024, 256, (RES, RES), (16, 16), 256, precision=precision)
    if load_weights:
        gpt.load_state_dict(synth.gpt_state_dict(gpt_sizes(cfg), seed=2), strict=False)
        fs.load_state_dict(synth.vqgan_state_dict(dd, seed=1))
        cs.load_state_dict(synth.vqgan_state_dict(ddb, seed=1), strict=False)
    model = Net2NetTransformer(gpt, fs, cs, top_k=100).to(dev).eval()
    model.sample_seed = 1234
    return model


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference(loop_steps, warmup, stage1_images=2, use_reference_modules=False):
    """The reference ALGORITHM on the host cores through the CPU oracle port (oracle/gpt_oracle.py, oracle/vqgan_oracle.py; pinned to the
    reference by tests/golden): per generated token ONE full 1792-token forward of the 24-layer GPT for one scene (no KV cache,
    cond_transformer_multi_view.py:172-219) + top-k / softmax / multinomial.  `loop_steps` iterations are timed and extrapolated to the
    1536 of a scene; the stage-1 work of a scene (7 encodes, 12 decodes) is timed on `stage1_images` images and scaled."""
    import torch
    from bevgen_b200.gpt_config import GPTConfig
    from oracle import gpt_oracle, synth, vqgan_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = GPTConfig(**gpt_kw())
    geo = gpt_oracle.geo_from_config(cfg)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    _, bev, mats = synth.stage2_inputs(1, seed=0)
    dd = synth.vqgan_ddconfig()
    vsd = synth.vqgan_state_dict(dd, seed=1)
    x = synth.image_batch(stage1_images, 3, RES, RES, seed=100)
    with torch.no_grad():
        t0 = time.perf_counter()
        quant, idx, _ = vqgan_oracle.encode(x, vsd)
        t_enc = (time.perf_counter() - t0) / stage1_images
        t0 = time.perf_counter()
        vqgan_oracle.decode(quant, vsd)
        t_dec = (time.perf_counter() - t0) / stage1_images
    stage1_s = 7 * t_enc + 12 * t_dec            # 6 camera + 1 BEV encode; 6 reconstruction + 6 generated decodes
    ncam, ntok = cfg.num_cams, cfg.num_cam_tokens
    forward, kind = (lambda xt: gpt_oracle.forward(sd, geo, xt, bev, mats, sampling=True)), "port"
    if use_reference_modules:
        # In the build container (where /root/reference exists) the sample loop runs the reference's OWN GPT module, imported unmodified
        # (oracle/ref_import.py: the absent DeepSpeed block-sparse kernels replaced by the dense stand-in the goldens were minted with).
        # On the GPU box the tree is absent and the pinned port above is what runs.
        try:
            from oracle import ref_import
            if ref_import.available():
                m = ref_import.stage2()
                model = m.GPT(m.GPTConfig(**gpt_kw())).eval()
                missing, unexpected = model.load_state_dict(sd, strict=False)
                assert not unexpected and all("master_layout" in k or k == "bev_grid" for k in missing)
                forward, kind = (lambda xt: model(xt, bev, mats, sampling=True)), "reference"
        except Exception as ex:          # any import problem: the port is the documented fallback
            print(f"[bench] reference modules not usable ({ex!r}); timing the oracle port", file=sys.stderr)
    xtok = torch.full((1, ncam, ntok), cfg.vocab_size, dtype=torch.int64)
    g = torch.Generator().manual_seed(0)
    times = []
    with torch.no_grad():
        for t in range(warmup + loop_steps):
            j = int(cfg.forward_shuffle_idx[t])
            i, k = j // ntok, j % ntok
            t0 = time.perf_counter()
            logits = forward(xtok).view(1, ncam, ntok, -1)[:, i, k]
            p = gpt_oracle.sample_probs(logits, 1.0, 100)
            xtok[:, i, k] = torch.multinomial(p, 1, generator=g)[:, 0]
            dt = time.perf_counter() - t0
            if t >= warmup:
                times.append(dt)
    step_s = sum(times) / len(times)
    scene_s = stage1_s + step_s * (TOKENS + 1)          # + the teacher-forced forward of shared_step
    return {"images_per_s": CAMS / scene_s, "cores": cores, "loop_step_s": step_s, "stage1_s_per_scene": stage1_s, "scene_s_extrapolated": scene_s,
            "timed_loop_steps": len(times), "kind": kind}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup, use_reference_modules=True)
    src = ("the reference's own GPT module imported from /root/reference (DeepSpeed's absent block-sparse kernels = the dense stand-in of "
           "oracle/ref_import.py), stage 1 through the port" if r["kind"] == "reference" else "CPU oracle port (pinned to the reference by tests/golden)")
    sample = (f"{src}; reference algorithm, 1 scene: {r['timed_loop_steps']} timed iterations of the sample loop (one 24-layer 1792-token forward each, "
              f"{r['loop_step_s']:.2f} s/iteration) extrapolated x1537, + stage-1 of the scene measured on 2 images ({r['stage1_s_per_scene']:.1f} s); "
              f"torch CPU fp32, {r['cores']} threads; EXTRAPOLATED: a full scene would take {r['scene_s_extrapolated'] / 3600:.2f} h")
    line = {"impl": "reference", "metric": METRIC, "value": r["images_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["loop_step_s"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "extrapolated": True,
            "config": {**CONFIG, "parallelism": f"scene-sharded x{args.gpus} (16 scenes per GPU), weights broadcast once, no data-path collective"},
            "step_definition": "one iteration of the reference's sample loop for one scene (a bounded sample of the workload: the full step is 1536 of them per scene x 16 scenes)",
            "cpu_baseline": {"value": r["images_per_s"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["images_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ output self-check
def self_check(model, batch_dev, out, rank):
    """After the timed loop: the step's outputs against the committed reference goldens (rank 0's batch is the tensor they were minted on)."""
    import numpy as np
    import torch
    from oracle import vqgan_oracle
    res = {}
    gen, rec = out["gen"], out["rec"]
    res["shapes_ok"] = tuple(gen.shape) == (SCENES, CAMS, 3, RES, RES) and tuple(rec.shape) == tuple(gen.shape)
    res["gen_finite_in_0_1"] = bool(torch.isfinite(gen).all() and gen.min() >= 0 and gen.max() <= 1)
    if rank != 0:
        return res
    try:
        g = np.load(ROOT / "tests" / "golden" / "vqgan_config2_rgb.npz")
        gb = np.load(ROOT / "tests" / "golden" / "vqgan_config2_bev.npz")
        x, c = model.get_xc(batch_dev)
        _, zi = model.encode_to_z(x, batch_dev)
        _, ci = model.encode_to_c(c, batch_dev)
        mism = int((zi.reshape(-1)[:2048].cpu().numpy() != g["idx"]).sum())
        mism_b = int((ci.reshape(-1)[:512].cpu().numpy() != gb["idx"]).sum())
        want = vqgan_oracle.denormalize(torch.from_numpy(g["rec_sub"]))
        rec8 = rec.reshape(-1, 3, RES, RES)[:8, :, ::8, ::8].cpu()
        res.update(rgb_token_mismatches_vs_reference_golden=f"{mism}/2048", bev_token_mismatches_vs_reference_golden=f"{mism_b}/512",
                   rec_max_err_vs_reference_golden=float((rec8 - want).abs().max()) if mism == 0 else None,
                   loss_finite=bool(torch.isfinite(model.last_test_loss)), loss=float(model.last_test_loss))
        res["ok"] = bool(res["shapes_ok"] and res["gen_finite_in_0_1"] and mism <= 2 and mism_b <= 1 and res["loss_finite"]
                         and (res["rec_max_err_vs_reference_golden"] is None or res["rec_max_err_vs_reference_golden"] < 1e-3))
    except Exception as ex:      # the bench line must still be printed
        res["error"] = repr(ex)
    return res


# ------------------------------------------------------------------------------------------------ main
NCU_DECODE_TRAFFIC = 3842125670656 + 106683657984     # bytes per launch of the persistent decode kernel (ncu, see roofline.traffic_source)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16f8", choices=["f16f8", "fp32x3", "bf16"],
                    help="f16f8 / fp32x3: parity modes (fp32-equivalent split products, 1e-3 bar); bf16: single pass (fast mode, misses the bar)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1]/[2] sub-objects, the torch GPU baseline and MaskGit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from bevgen_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bcast = {"ms": 0.0, "bytes": 0}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        t = torch.ones(1, device=dev)
        dist.all_reduce(t)                     # NCCL communicator set-up + first-collective cost stays out of the broadcast stopwatch
        torch.cuda.synchronize()

    # ---- model: weights generated on rank 0 only, NCCL-broadcast (weights + integer layout buffers; the data path has no collective)
    model = build_model(args.precision, dev, load_weights=(rank == 0))
    if world > 1:
        from bevgen_b200.sharding import broadcast_module_weights
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bcast["bytes"] = broadcast_module_weights(model, src=0)
        torch.cuda.synchronize()
        bcast["ms"] = (time.perf_counter() - t0) * 1e3

    batch_host = host_batch(SCENES, rank)
    batch_dev = model.batch_to_device(batch_host)
    n_img = SCENES * CAMS
    gen_host = torch.empty((SCENES, CAMS, 3, RES, RES), dtype=torch.float32).pin_memory()
    rec_host = torch.empty_like(gen_host).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in batch_host.values() if torch.is_tensor(v))
    d2h = gen_host.numel() * 4 + rec_host.numel() * 4 + 4
    last = {}

    def step_resident():
        last["out"] = model.test_step(batch_dev, 0)

    def step_e2e():          # host tensors in (one async upload inside test_step), generated + reconstructed images and the loss back on the host
        out = model.test_step(batch_host, 0)
        gen_host.copy_(out["gen"], non_blocking=True)
        rec_host.copy_(out["rec"], non_blocking=True)
        last["loss"] = float(model.last_test_loss)          # device -> host read of the step's metric (synchronises the step)
        last["out"] = out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, w0, time.time()

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ops.Stats.reset()
    smp = model.transformer.sampler(SCENES)
    smp.timing = []                                   # the sampler records CUDA events around its decode loop (launching stream)
    ms, w0, w1 = timed(step_resident, args.steps)
    launches = ops.Stats.launches
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    value = world * n_img * args.steps / (ms / 1e3)
    decode_ms = [a.elapsed_time(b) for a, b in smp.timing]
    smp.timing = None

    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    e2e_value = world * n_img * args.steps / (ms_e2e / 1e3)
    check = self_check(model, batch_dev, last["out"], rank)
    e2e_host_ok = bool(torch.isfinite(gen_host).all()) and float(gen_host.max()) <= 1.0

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    sustained, burst, hbm, how = peaks()
    dec_ms = statistics.mean(decode_ms) if decode_ms else float("nan")
    algo = algo_bytes_decode(SCENES)
    achieved = algo / (dec_ms / 1e3) / 1e9
    packed = smp.bytes_per_batch()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"f16f8": "fp32-equivalent split products (fp16 + e4m3 residual planes / bf16x3, fp32 accumulate), fp16 KV cache",
                  "fp32x3": "bf16x3 split products (fp32-equivalent, fp32 accumulate), fp16 KV cache", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {**CONFIG, "parallelism": f"scene-sharded x{world} (16 scenes per GPU), weights broadcast once, no data-path collective"},
        "precision": args.precision,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "api": "Net2NetTransformer.test_step(host batch dict) + gen/rec/loss copied to pinned host memory", "host_result_ok": e2e_host_ok},
        "gpu_launches": launches,
        "clocks": clocks,
        "self_check": check,
        "weights_broadcast": {**bcast, "value_including_broadcast": world * n_img * args.steps / ((ms + bcast["ms"]) / 1e3)},
        "roofline": {"bound": "hbm", "kernel": smp.kernel_name, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({how})",
                     "algorithmic_bytes_per_launch": algo / TOKENS if smp.launches_per_token == 1 else algo,
                     "algorithmic_bytes": f"SURVEY 8d config 4: 556 MB weights x {TOKENS} steps + bf16 KV reads = {algo / 1e12:.3f} TB per {SCENES} scenes",
                     "decode_loop_ms": dec_ms, "ms_per_token_step": dec_ms / (TOKENS - 1), "share_of_step": dec_ms / (ms / args.steps),
                     "launches_per_token_step": smp.launches_per_token,
                     "bytes_in_the_shipped_format": packed, "achieved_in_shipped_format_GBps": packed / (dec_ms / 1e3) / 1e9,
                     "traffic": smp.ncu_traffic_bytes if smp.ncu_traffic_bytes is not None else (NCU_DECODE_TRAFFIC if smp.persistent else None),
                     "traffic_source": smp.ncu_traffic_source if smp.ncu_traffic_source is not None else
                     ("profiles/r03_decode_full_launch_ncu.csv: dram__bytes_read.sum + dram__bytes_write.sum of ONE decode_persistent_kernel launch "
                      "(1535 token steps, 16 scenes, 24 layers) = 3.842 TB + 0.107 TB, against 3.869 TB in the shipped format" if smp.persistent else None)},
    }
    if not args.no_extras and world == 1:
        from tools import bench_extras
        del last["out"]
        line.update(bench_extras.run(model, batch_dev, args.precision, sustained, hbm))
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference(2, 1)
        line["cpu_baseline"] = {"value": r["images_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": (f"reference algorithm (no KV cache), 1 scene: 2 timed iterations of the sample loop ({r['loop_step_s']:.2f} s each) "
                                           f"extrapolated x1537 + stage-1 measured on 2 images; torch CPU fp32, {r['cores']} threads; EXTRAPOLATED")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
