"""Host-buffer round trips through the stage-1 model with the PCIe copies off the compute stream.

`VQModel.encode / decode` take device tensors; a caller holding pinned HOST images (the dataloader side of generate.py / the stage-1
reconstruction sweep) would otherwise serialise  H2D -> encode -> decode -> D2H  on one stream.  `RoundTripPipeline.submit` issues the
input copy of the NEXT batch and the result copy of the PREVIOUS one on two side streams while the current batch computes (two input
slots, events for slot reuse, `record_stream` for the result tensors), so a steady stream of batches runs at the compute rate.
torch is used for streams / events / pinned copies only.
"""
import torch


class RoundTripPipeline:
    def __init__(self, model, device, example_host: torch.Tensor):
        self.model, self.dev = model, torch.device(device)
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.slots = [torch.empty(example_host.shape, dtype=example_host.dtype, device=self.dev) for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [None, None]
        self.ev_out = None
        self.n, self.staged = 0, None

    def _stage(self, slot, x_host):
        with torch.cuda.stream(self.s_in):
            if self.ev_free[slot] is not None:
                self.s_in.wait_event(self.ev_free[slot])          # the step that read this slot has finished with it
            self.slots[slot].copy_(x_host, non_blocking=True)
            self.ev_in[slot].record(self.s_in)

    def submit(self, x_host, rec_host, idx_host, next_x_host=None):
        """One encode -> quantise -> decode round trip of the pinned batch `x_host` into the pinned `rec_host` / `idx_host`.  Pass the
        following batch as `next_x_host` to have its upload overlap this step.  Results are complete after `drain()`."""
        slot = self.n & 1
        if self.staged is not x_host:
            self._stage(slot, x_host)
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self.ev_in[slot])
        self.staged = None
        if next_x_host is not None:
            self._stage(slot ^ 1, next_x_host)
            self.staged = next_x_host
        quant, _, (_, _, idx) = self.model.encode(self.slots[slot], None)
        rec = self.model.decode(quant)
        done = torch.cuda.Event()
        done.record(main)
        self.ev_free[slot] = done
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            rec_host.copy_(rec, non_blocking=True)
            idx_host.copy_(idx, non_blocking=True)
            rec.record_stream(self.s_out)
            idx.record_stream(self.s_out)
            self.ev_out = torch.cuda.Event()
            self.ev_out.record(self.s_out)
        self.n += 1

    def drain(self, block_host: bool = True):
        """Wait for every outstanding result copy.  The current stream always waits (so a CUDA-event timer recorded after drain()
        covers the copies); with `block_host` (default) the HOST blocks too, after which the pinned result buffers are safe to read.
        `block_host=False` is the stream-only variant for timing loops that synchronise later."""
        if self.ev_out is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.ev_out)
            if block_host:
                self.ev_out.synchronize()
