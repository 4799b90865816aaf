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
    def __init__(self, params, bucket_mb=25.0, process_group=None, average=True, tail_mb=1.0):
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
        # The LAST bucket's all-reduce cannot overlap anything: its final gradient (the first layer's) is the end of backward.
        # Keep that exposed collective small: the parameters that finish last (the first ~tail_mb of the model) get a bucket of
        # their own, whatever is in front of them goes out earlier.
        last = self.buckets[-1]
        tail_cap = max(1, int(tail_mb * (1 << 20) / 4))
        if len(last.params) > 1 and last.end - last.start > 2 * tail_cap:
            k = len(last.params)
            while k > 1 and last.end - self._slot[id(last.params[k - 1])][0] <= tail_cap:
                k -= 1
            if 0 < k < len(last.params):
                cut = self._slot[id(last.params[k])][0]
                tail = _Bucket(cut)
                tail.end, tail.params = last.end, last.params[k:]
                last.end, last.params = cut, last.params[:k]
                for p in tail.params:
                    self._bucket_of[id(p)] = tail
                self.buckets.append(tail)
        for b in self.buckets:
            b.pending = len(b.params)
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._got = set()              # ids of the parameters whose gradient landed this step
        self.missing = []              # parameters that received no gradient in the step `finish()` closed last
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
        self._got.add(id(p))
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
                if self.device.type == "cuda":
                    from . import ops
                    ops.join_wgrad_stream(self.comm_stream)      # weight gradients are produced on a second stream
                op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
                b.work = dist.all_reduce(g, op=op, group=self.group, async_op=True)
        else:
            # gloo (CPU tests of the host logic) has no AVG
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                g.div_(self.world)

    def finish(self):
        """Call after backward: waits for every bucket; gradients are then the mean over ranks."""
        # Parameters that received no gradient this step (frozen sub-graph, unused branch): their slot still holds the
        # PREVIOUS step's (possibly already averaged) gradient.  Zero it before anything reads the flat buffer; the optimizer
        # skips these parameters altogether, like torch.optim.Adam skips ``p.grad is None``.
        self.missing = [p for p in self.params if id(p) not in self._got] if len(self._got) != len(self.params) else []
        for p in self.missing:
            o, n = self._slot[id(p)]
            self.flat_grad[o:o + n].zero_()
        self._got = set()
        for b in self.buckets:
            if b.pending != 0 and self.world > 1:
                # reduce the bucket anyway so ranks stay in step
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
            ops.grads_cleared(self._slot)

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        if self.device.type == "cuda":
            from . import ops
            ops.set_grad_flat(None, self._slot)


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


class FusedAdam(torch.optim.Optimizer):
    """``optim.Adam(model.parameters(), lr, weight_decay=1e-5)`` (reference train.py:312: L2-coupled decay) as ONE kernel over
    the flat parameter / gradient buffers of a ``GradBuckets``.

    A genuine ``torch.optim.Optimizer``: one ``param_groups`` entry over every parameter, whose ``lr`` / ``betas`` / ``eps`` /
    ``weight_decay`` are read at every ``step()`` -- so ``StepLR(optimizer, 30, 0.5)`` (train.py:315,118) schedules it -- and
    ``state_dict()`` / ``load_state_dict()`` in ``torch.optim.Adam``'s own format (``step``, ``exp_avg``, ``exp_avg_sq`` per
    parameter index), so an optimizer checkpoint moves between the two.  Parameters that received no gradient in a step are
    skipped (no decay, no moment update, their step count does not advance), like ``torch.optim.Adam`` does for
    ``p.grad is None``."""

    def __init__(self, buckets, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        from . import _lib

        self._lib = _lib
        self.b = buckets
        super().__init__([{"params": list(buckets.params)}], dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.m = torch.zeros_like(buckets.flat_param)
        self.v = torch.zeros_like(buckets.flat_param)
        self.t = 0
        self._lag = {}                 # id(param) -> number of steps it was skipped in (no gradient)

    # hyper-parameters live in param_groups[0] (what LR schedulers rewrite); these names are kept for older callers
    lr = property(lambda self: self.param_groups[0]["lr"], lambda self, v: self.param_groups[0].__setitem__("lr", v))
    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    wd = property(lambda self: self.param_groups[0]["weight_decay"])

    def _ranges(self):
        """[(start, end, step)] element ranges of the flat buffers to update: everything in ONE range normally; when some
        parameters have missed gradients (now or earlier) the ranges leave out the ones missing now and carry a per-range
        step count, exactly like torch.optim.Adam's per-parameter ``state['step']``"""
        b = self.b
        total = b.flat_param.numel()
        if not b.missing and not self._lag:
            return [(0, total, self.t)]
        missing = {id(p) for p in b.missing}
        out = []
        for p in sorted(b.params, key=lambda q: b._slot[id(q)][0]):
            if id(p) in missing:
                continue
            o, n = b._slot[id(p)]
            end = (o + n + 3) // 4 * 4
            t = self.t - self._lag.get(id(p), 0)
            if out and out[-1][1] == o and out[-1][2] == t:
                out[-1] = (out[-1][0], end, t)
            else:
                out.append((o, end, t))
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.t += 1
        g = self.param_groups[0]
        b = self.b
        for p in b.missing:
            self._lag[id(p)] = self._lag.get(id(p), 0) + 1
        if b.flat_param.is_cuda:
            st = torch.cuda.current_stream(b.flat_param.device).cuda_stream
            with torch.cuda.device(b.flat_param.device):
                for lo, hi, t in self._ranges():
                    if t < 1:
                        continue
                    self._lib.call("eel_adam_step", b.flat_param.data_ptr() + 4 * lo, b.flat_grad.data_ptr() + 4 * lo,
                                   self.m.data_ptr() + 4 * lo, self.v.data_ptr() + 4 * lo, hi - lo, float(g["lr"]),
                                   float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                                   int(t), st)
        else:
            raise self._lib.EelError("FusedAdam runs on CUDA only (no CPU fallback)")
        from . import ops
        ops.weights_changed()      # the kernel rewrote the parameters through raw pointers: packed copies are stale
        return loss

    def zero_grad(self, set_to_none=True):
        self.b.zero_grad()

    # ---- checkpointing in torch.optim.Adam's format (the reference saves no optimizer state, train.py:157-197; a user who
    # adds it gets files that torch.optim.Adam reads, and vice versa)
    def state_dict(self):
        state = {}
        if self.t > 0:
            for i, p in enumerate(self.b.params):
                o, n = self.b._slot[id(p)]
                t = self.t - self._lag.get(id(p), 0)
                if t < 1:
                    continue
                state[i] = {"step": torch.tensor(float(t)), "exp_avg": self.m[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.v[o:o + n].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.b.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.b.params):
            raise ValueError("FusedAdam.load_state_dict: expected one param group over %d parameters" % len(self.b.params))
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = tuple(v) if k == "betas" else v
        self.m.zero_()
        self.v.zero_()
        steps = {}
        for i, p in enumerate(self.b.params):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            steps[id(p)] = 0 if st is None else int(float(st["step"]))
            if st is None:
                continue
            o, n = self.b._slot[id(p)]
            self.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
        self.t = max(steps.values()) if steps else 0
        self._lag = {k: self.t - v for k, v in steps.items() if v != self.t}
