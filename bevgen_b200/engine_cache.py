"""Cache key / invalidation for the pre-packed weight engines held by the drop-in modules (VQModel, GPT, MaskGit).

The engines pack a module's parameters once (split planes, tap-major conv layouts, the tiled camera-bias table ...), so they
must be rebuilt whenever a weight OR a buffer (e.g. `master_layout`) changes.  `Parameter._version` alone misses writes made
through `.data` (EMA swaps, `weight.data.normal_()`, a broadcast into `p.data`), so the key is

    (device, precision, [(data_ptr, _version) of every parameter and buffer])

and three explicit hooks drop the engine as well: `load_state_dict`, `_apply` (`.cuda()/.to()/.half()`), and
`invalidate_engine()` — which `bevgen_b200.sharding.broadcast_module_weights` calls on every sub-module after the copy.
A caller that writes through `.data` on a live model must call `invalidate_engine()` itself (documented in INTEGRATION.md).
"""
import torch


def fingerprint(module: torch.nn.Module):
    ts = list(module.parameters()) + [b for b in module.buffers() if b is not None]
    return tuple((t.data_ptr(), t._version) for t in ts)


class EngineCacheMixin:
    """Mixed into an nn.Module that owns `self._engine` / `self._engine_key`."""

    def _engine_cache_key(self, device, precision):
        return (device, precision, fingerprint(self))

    def invalidate_engine(self):
        self._engine, self._engine_key = None, None
        if hasattr(self, "_samplers"):
            self._samplers = {}

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate_engine()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.invalidate_engine()
        return out


def invalidate_all(module: torch.nn.Module):
    for m in module.modules():
        if hasattr(m, "invalidate_engine"):
            m.invalidate_engine()
