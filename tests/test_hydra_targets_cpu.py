"""The reference instantiates its models from YAML through Hydra `_target_` class paths (configs/model/stage_2.yaml:1-81).  Hydra is not in
this image, so a small instantiator with the same semantics (multi_view_generation/utils/instantiate.py) builds the drop-in modules from a
config of the reference's shape: same class paths, same field names, `${...}` interpolations against the experiment root."""
from pathlib import Path

import torch

from multi_view_generation.utils.instantiate import instantiate, load_yaml

CFG = Path(__file__).parent / "configs" / "stage_2_small.yaml"


def test_yaml_targets_instantiate_the_drop_in_modules():
    cfg = load_yaml(CFG)
    t = cfg["model"]["transformer"]["cfg"]
    assert t["num_cams"] == 6 and t["vocab_size"] == 1024 and t["cam_res"] == [256, 256] and t["cam_latent_res"] == [16, 16]      # interpolations
    m = instantiate(cfg["model"])
    assert type(m).__module__ == "multi_view_generation.modules.stage2.cond_transformer_multi_view"
    assert type(m.transformer).__module__ == "multi_view_generation.modules.transformer.mingpt_sparse"
    assert type(m.first_stage_model).__name__ == "VQModel" and type(m.cond_stage_model).__name__ == "VQSegmentationModel"
    assert type(m.first_stage_model.loss).__name__ == "DummyLoss"
    assert m.top_k == 20 and m.cond_stage_key == "segmentation" and m.skip_sampling is False
    assert m.cfg.gpt_block_size == 6 * 256 + 256 and m.cfg.num_img_tokens == 1536
    keys = set(m.state_dict().keys())
    for k in ("transformer.blocks.0.attention.query.weight", "transformer.blocks.1.mlp.2.bias", "transformer.x_tok_emb.weight", "transformer.camera_bias_emb", "transformer.head.weight",
              "first_stage_model.encoder.conv_in.weight", "first_stage_model.quantize.embedding.weight", "first_stage_model.post_quant_conv.bias",
              "cond_stage_model.decoder.conv_out.weight", "cond_stage_model.colorize"):
        assert k in keys, k
    assert m.cond_stage_model.encoder.conv_in.weight.shape[1] == 7
    assert not m.first_stage_model.training and not m.cond_stage_model.training       # stage-1 models are put in eval mode, as the reference's init_*_from_ckpt


def test_overrides_and_plain_values():
    cfg = load_yaml(CFG, num_cams=3)
    assert cfg["model"]["transformer"]["cfg"]["num_cams"] == 3
    gcfg = instantiate(cfg["model"]["transformer"]["cfg"], cam_names="NUSCENES_ABLATION_CAMERAS")
    assert gcfg.num_cams == 3 and gcfg.num_img_tokens == 3 * 256
    assert instantiate({"a": [1, {"b": 2}]}) == {"a": [1, {"b": 2}]}
    assert isinstance(instantiate({"_target_": "torch.nn.Linear", "in_features": 4, "out_features": 2}), torch.nn.Linear)
