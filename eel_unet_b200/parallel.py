"""Batch-sharded data parallelism for the EEL-Unet step.

One process per GPU.  Every parameter is re-homed into ONE flat fp32 buffer and its gradient into a
matching flat buffer, cut into size-balanced buckets in reverse-forward order (head / dec1 first, the
order backward produces them).  A post-accumulate hook per parameter copies the fresh gradient into its
bucket slot; when the last gradient of a bucket lands, the bucket is all-reduced (mean) on a side stream
while backward keeps running on the compute stream.  ``finish()`` joins the streams.  BatchNorm statistics
stay per replica, exactly like N independent copies of the reference (it has no SyncBN).

With world_size == 1 no collective is issued; the flat buffers still let the optimizer run as one kernel.
The reference has no distributed code at all; this replaces what a user would otherwise get from wrapping
the reference model in torch DistributedDataParallel.
"""
import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("start", "end", "params", "pending", "work", "event")

    def __init__(self, start):
        self.start, self.end, self.params, self.pending, self.work, self.event = start, start, [], 0, None, None


class GradBuckets:
    def __init__(self, params, bucket_mb=25.0, process_group=None, average=True):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        self.params = params
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.average = average
        dev = params[0].device
        self.device = dev
        order = list(reversed(params))            # backward produces gradients roughly in this order
        offs, total = {}, 0
        for p in order:
            offs[id(p)] = total
            total += (p.numel() + 3) // 4 * 4     # keep every slot 16-byte aligned
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets, self._bucket_of, self._slot = [], {}, {}
        cap = max(1, int(bucket_mb * (1 << 20) / 4))
        cur = _Bucket(0)
        for p in order:
            o, n = offs[id(p)], p.numel()
            with torch.no_grad():
                self.flat_param[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[o:o + n].view(p.shape)
            self._slot[id(p)] = (o, n)
            if cur.params and (o + n - cur.start) > cap:
                cur.end = o
                self.buckets.append(cur)
                cur = _Bucket(o)
            cur.params.append(p)
            self._bucket_of[id(p)] = cur
            cur.end = (o + n + 3) // 4 * 4
        self.buckets.append(cur)
        for b in self.buckets:
            b.pending = len(b.params)
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        if dev.type == "cuda":
            from . import ops
            ops.set_grad_flat(self.flat_grad, self._slot)

    # ------------------------------------------------------------------------------------------
    def _on_grad(self, p):
        o, n = self._slot[id(p)]
        slot = self.flat_grad[o:o + n].view(p.shape)
        if p.grad.data_ptr() != slot.data_ptr():
            # (the backward kernels of eel_unet_b200.ops write most gradients straight into their slot: ops._grad_out)
            slot.copy_(p.grad)
            p.grad = slot                  # the user-visible gradient now lives in the flat buffer
        b = self._bucket_of[id(p)]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b):
        if self.world == 1:
            return
        g = self.flat_grad[b.start:b.end]
        if self.comm_stream is not None:
            # NCCL averages in the collective itself (ReduceOp.AVG): no extra scaling pass over the bucket
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
                b.work = dist.all_reduce(g, op=op, group=self.group, async_op=True)
        else:
            # gloo (CPU tests of the host logic) has no AVG
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                g.div_(self.world)

    def finish(self):
        """Call after backward: waits for every bucket; gradients are then the mean over ranks."""
        for b in self.buckets:
            if b.pending != 0 and self.world > 1:
                # parameters that received no gradient this step: reduce the bucket anyway so ranks stay in step
                self._launch(b)
            if b.work is not None:
                b.work.wait()
                b.work = None
            b.pending = len(b.params)
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)

    def zero_grad(self):
        """set_to_none semantics: the next backward writes fresh gradients (no accumulate kernel per tensor)."""
        for p in self.params:
            p.grad = None
        if self.device.type == "cuda":
            from . import ops
            ops.grads_cleared()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        if self.device.type == "cuda":
            from . import ops
            ops.set_grad_flat(None, None)


class DataParallel(torch.nn.Module):
    """``DataParallel(model)``: same call signature and state_dict as the wrapped model."""

    def __init__(self, module, bucket_mb=25.0, process_group=None):
        super().__init__()
        self.module = module
        self.name = getattr(module, "name", None)
        self.buckets = GradBuckets(list(module.parameters()), bucket_mb, process_group)

    def forward(self, *a, **k):
        return self.module(*a, **k)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self.module.load_state_dict(*a, **k)

    def finish_backward(self):
        self.buckets.finish()

    def zero_grad(self, set_to_none=True):
        self.buckets.zero_grad()


class FusedAdam:
    """optim.Adam(lr, weight_decay) with L2-coupled decay (reference train.py:312) as one kernel over the flat buffers."""

    def __init__(self, buckets, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        from . import _lib

        self._lib = _lib
        self.b = buckets
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(buckets.flat_param)
        self.v = torch.zeros_like(buckets.flat_param)
        self.t = 0

    def step(self):
        self.t += 1
        b = self.b
        self._lib.call("eel_adam_step", b.flat_param.data_ptr(), b.flat_grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                       b.flat_param.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                       float(self.wd), int(self.t), self._lib.stream())
        from . import ops
        ops.weights_changed()      # the kernel rewrote the parameters through raw pointers: packed copies are stale

    def zero_grad(self, set_to_none=True):
        self.b.zero_grad()
