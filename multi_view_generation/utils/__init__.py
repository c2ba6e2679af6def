"""utils subset: logger + `init_from_ckpt` with the reference's key-mapping semantics (utils/general.py:119-160)."""
import logging
import os

import torch


def get_pylogger(name=__name__):
    return logging.getLogger(name)


log = get_pylogger(__name__)


def init_from_ckpt(module, path, ignore_keys=list(), unfrozen_keys=list(), strict=False):
    """Load a checkpoint file: unwrap `state_dict`, strip `_forward_module.` prefixes, delete keys containing any of
    `ignore_keys`, non-strict load.  (DeepSpeed ZeRO checkpoint *directories* need deepspeed and are not supported.)"""
    if os.path.isdir(path):
        raise NotImplementedError("DeepSpeed ZeRO checkpoint directories need deepspeed's zero_to_fp32 (not available)")
    sd = torch.load(path, map_location="cpu")
    if "state_dict" in sd.keys():
        sd = sd["state_dict"]
    for k in list(sd):
        if k.startswith("_forward_module"):
            sd[k.replace("_forward_module.", "")] = sd.pop(k)
    for k in list(sd):
        if any(ik in k for ik in ignore_keys):
            log.info("Deleting key %s from state_dict.", k)
            del sd[k]
    own = module.state_dict().keys()
    for n in own:
        if n not in sd:
            print(f"Missing {n}")
    for n in sd:
        if n not in own:
            print(f"Unexpected {n}")
    module.load_state_dict(sd, strict=strict)
    log.info("Restored from %s", path)
