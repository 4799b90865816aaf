"""torch.autograd bindings of the libeel.so kernels (host-side mirror of the ATen ops the reference's
hot path dispatches).  Activations are NHWC tensors [N, H, W, C] in fp32 or bf16; parameters are the
module's own fp32 nn.Parameters in the reference's layout.  Every op here ends in a C-ABI call --
there is no PyTorch / CPU fallback.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import call, dtype_code, ptr, stream, workspace

F32 = torch.float32


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# Flat gradient buffers of parallel.GradBuckets: id(param) -> [flat fp32 tensor, offset, numel, handed out this step].
# Backward kernels then write parameter gradients straight into their slots (a FRESH view per call, which autograd's
# AccumulateGrad adopts without a copy) instead of into temporaries that a copy per parameter moves there afterwards.
# Several GradBuckets (several models) can be registered at once; each removes its own entries.
_GRAD_SLOTS = {}


def set_grad_flat(flat, slots):
    """register (flat is a tensor) or unregister (flat is None) the slots {id(param): (offset, numel)} of one GradBuckets"""
    for pid, (o, n) in (slots or {}).items():
        if flat is None:
            _GRAD_SLOTS.pop(pid, None)
        else:
            _GRAD_SLOTS[pid] = [flat, o, n, False]


def grads_cleared(slots=None):
    """parallel.GradBuckets.zero_grad: its slots may be handed out again"""
    for pid in (slots if slots is not None else list(_GRAD_SLOTS)):
        e = _GRAD_SLOTS.get(pid)
        if e is not None:
            e[3] = False


def _grad_out(param, shape=None):
    """fp32 tensor a backward kernel writes ``param``'s gradient into"""
    shape = tuple(param.shape) if shape is None else tuple(shape)
    e = _GRAD_SLOTS.get(id(param))
    if e is not None and param.grad is None and not e[3] and e[0].device == param.device:
        e[3] = True
        return e[0][e[1]:e[1] + e[2]].view(shape)
    return torch.empty(shape, dtype=F32, device=param.device)


# ---- side channels between neighbouring autograd Functions ------------------------------------------------------------
# A producer sometimes computes, for free, something its neighbour would otherwise need a full pass for (BatchNorm sums out of
# a GEMM epilogue, a bias gradient out of a BatchNorm backward, ...).  The hand-over travels ON THE TENSOR that connects the
# two (a Python attribute on the very tensor object autograd passes along) together with the tensor's version counter: a
# consumer only sees it on that same object and only while nobody wrote to the tensor in between (autograd accumulating a
# second gradient in place, a user hook returning another tensor, an in-place op: the attribute is absent or its version is
# stale, and the consumer falls back to computing the quantity itself).  Nothing is keyed by raw addresses, nothing is global.
def _attach(t, key, value):
    setattr(t, key, (value, t._version))


def _take(t, key):
    hit = t.__dict__.pop(key, None) if hasattr(t, "__dict__") else None
    if hit is None or hit[1] != t._version:
        return None
    return hit[0]


# ---- weight-gradient stream -------------------------------------------------------------------------------------------
# Backward is a serial chain of data gradients and BatchNorm backward passes; the WEIGHT gradients hang off it as leaves
# (nothing in the backward reads them).  They are launched on a second stream per device: the tensor-core weight-gradient
# kernel of layer i then shares the GPU with the HBM-bound BatchNorm backward of layer i-1 instead of queueing in front of it.
# Ordering: the side stream waits for everything the launching stream has enqueued so far; the tensors it reads are kept
# referenced until the launching stream has been ordered behind their use (_Wgrad); the launching (and the caller's) stream
# join the side stream in a final callback of the autograd engine, and parallel.GradBuckets makes its communication stream
# wait for it too.
_WGRAD_ASYNC = __import__("os").environ.get("EEL_WGRAD_STREAM", "1") != "0"
_SIDE_STREAMS = {}        # device index -> torch.cuda.Stream
_JOIN_PENDING = {}        # (graph task id, launching stream handle, side stream handle) -> True while a join callback is queued
_JOIN_LOCK = __import__("threading").Lock()


# BatchNorm backward sums out of the producer of the gradient (A/B switches of the two newest producers)
_BRIDGE_BNSUMS = __import__("os").environ.get("EEL_BRIDGE_BNSUMS", "1") != "0"      # eel_add_interleave_bwd_bnsums
_BNSUMS_WIDE = __import__("os").environ.get("EEL_BNSUMS_WIDE", "1") != "0"          # conv data-gradient epilogue for every width
_BNSUMS_64 = __import__("os").environ.get("EEL_BNSUMS_64", "1") != "0"              # ... for the 64-channel full-resolution layers


def set_wgrad_stream(flag):
    """weight gradients on a second stream (default on; EEL_WGRAD_STREAM=0 turns it off at import)"""
    global _WGRAD_ASYNC
    _WGRAD_ASYNC = bool(flag)


def wgrad_stream(device_index=None):
    """the device's weight-gradient stream, or None when it was never used"""
    return _SIDE_STREAMS.get(torch.cuda.current_device() if device_index is None else device_index)


def join_wgrad_stream(waiter):
    """make ``waiter`` (a torch.cuda.Stream) wait for every weight gradient launched so far on its device"""
    side = _SIDE_STREAMS.get(waiter.device_index)
    if side is not None:
        waiter.wait_stream(side)


def _async_ok(*params):
    """the gradients may be produced on the side stream only if nothing reads them before the end of backward:
    AccumulateGrad must ADOPT them (p.grad is None, no tensor hooks), and a post-accumulate hook is only tolerated when it is
    GradBuckets' (the gradient was written into its flat-buffer slot, so the hook launches nothing on it)"""
    if not _WGRAD_ASYNC:
        return False
    for p in params:
        if p is None:
            continue
        if p.grad is not None or p._backward_hooks:
            return False
        if getattr(p, "_post_accumulate_grad_hooks", None):
            e = _GRAD_SLOTS.get(id(p))
            if e is None or e[3] or e[0].device != p.device:
                return False
    return True


class _Wgrad:
    """``with _Wgrad(on, t1, t2, ...):`` -- when ``on``, the launches inside go to the weight-gradient stream, ordered after
    everything enqueued so far on the current stream; t1, t2, ... are the tensors those launches READ.  Never pass the
    gradients they write: autograd's AccumulateGrad adopts a returned gradient only while nobody else references it and
    otherwise CLONES it on the spot -- on the launching stream, before the side stream has produced it.  (A small input that
    is also returned to autograd, like a bias gradient the side stream reads, may be passed: the clone is then of finished
    data.)  The outputs need no protection: they are flat-buffer slots or become ``p.grad``.

    Lifetimes are handled HERE, not with ``Tensor.record_stream``: the tensors stay referenced until the launching stream has
    been made to wait (a few weight gradients later, ``_WGRAD_LAG``) for the event recorded behind their last use, and only
    then are the references dropped.  Frees therefore happen at fixed points of the program and are ordered like ordinary
    same-stream frees; with record_stream the caching allocator saw blocks come back at times that depend on how far the side
    stream had got, kept growing (25 -> 48 GB reserved at batch 64) and cudaMalloc calls inside timed steps stalled them."""

    def __init__(self, on, *tensors):
        self.on, self.tensors, self.ctx, self.late_join = on, tensors, None, None

    def __enter__(self):
        if not self.on:
            return self
        main = torch.cuda.current_stream()
        idx = main.device_index
        side = _SIDE_STREAMS.get(idx)
        if side is None:
            side = _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
        side.wait_stream(main)
        self.main, self.side, self.idx = main, side, idx
        key = (torch._C._current_graph_task_id(), main.cuda_stream, side.cuda_stream)      # one join per backward pass
        with _JOIN_LOCK:
            if len(_JOIN_PENDING) > 256:       # keys of backward passes that died with an exception (a duplicate join is harmless)
                _JOIN_PENDING.clear()
            queued = _JOIN_PENDING.get(key)
            _JOIN_PENDING[key] = True
        if not queued:
            def join(main=main, side=side, key=key, idx=idx):
                with _JOIN_LOCK:
                    _JOIN_PENDING.pop(key, None)
                main.wait_stream(side)
                cur = torch.cuda.current_stream(idx)        # the stream backward() was called on
                if cur.cuda_stream != main.cuda_stream:
                    cur.wait_stream(side)
                _release_held(idx, main, 0)
            try:
                torch.autograd.Variable._execution_engine.queue_callback(join)
            except RuntimeError:
                # not inside a backward pass (a Function's backward called by hand): no leaf to protect, join right after
                with _JOIN_LOCK:
                    _JOIN_PENDING.pop(key, None)
                self.late_join = True
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is None:
            return False
        self.ctx.__exit__(*exc)
        ev = torch.cuda.Event()
        ev.record(self.side)
        with _JOIN_LOCK:
            _HELD.setdefault(self.idx, []).append((ev, self.tensors))
        if self.late_join:
            self.main.wait_stream(self.side)
            _release_held(self.idx, self.main, 0)
        else:
            _release_held(self.idx, self.main, _WGRAD_LAG)
        self.tensors = None
        return False


_HELD = {}            # device index -> [(event behind the last side-stream use, tensors kept alive for it)]
_WGRAD_LAG = 3        # weight gradients the side stream may fall behind before the launching stream waits for it


def _release_held(idx, main, keep):
    """drop the references of all but the newest ``keep`` side-stream launches, after ordering ``main`` behind them"""
    with _JOIN_LOCK:
        q = _HELD.get(idx)
        if not q or len(q) <= keep:
            return
        n = len(q) - keep
        done, _HELD[idx] = q[:n], q[n:]
    main.wait_event(done[-1][0])        # events of one stream complete in order: the newest of them covers all
    del done


# (the event behind the newest weight-packing launches lives in the forward's thread-local state: _TLS.pack_events[device index])


def _on_side_stream(fn):
    """The per-step weight packing (eel_pack_batch, eel_compose_batch: strided fp32 -> bf16 transposes, latency-bound) runs on
    the side stream, which idles during the forward: its first consumer is several kernels into the step, so it leaves the
    critical path.  Consumers order themselves behind it through ``_await_packed()``."""
    if not _WGRAD_ASYNC:
        fn()
        return
    main = torch.cuda.current_stream()
    idx = main.device_index
    side = _SIDE_STREAMS.get(idx)
    if side is None:
        side = _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    side.wait_stream(main)          # the optimizer (and the previous step's readers of the packed buffers) ran on `main`
    with torch.cuda.stream(side):
        fn()
    ev = torch.cuda.Event()
    ev.record(side)
    pe = getattr(_TLS, "pack_events", None)
    if pe is None:
        pe = _TLS.pack_events = {}
    pe[idx] = ev


def _await_packed():
    """before the first read of a packed / composed operand: order the current stream behind the packing launches (of this
    thread's forward: the packing and the first consumers run on the thread that calls the model)"""
    pe = getattr(_TLS, "pack_events", None)
    if pe:
        ev = pe.pop(torch.cuda.current_device(), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)


def _pack(w4, perm, dtype, out=None):
    """permute + cast a 4-D fp32 parameter into the operand layout a kernel wants."""
    w4 = _c(w4.detach())
    d = list(w4.shape)
    if out is None:
        out = torch.empty([d[p] for p in perm], dtype=dtype, device=w4.device)
    call("eel_permute4", ptr(w4), dtype_code(w4), ptr(out), dtype_code(out), d[0], d[1], d[2], d[3],
         perm[0], perm[1], perm[2], perm[3], stream())
    return out


def _as_dtype2d(w2, dtype):
    """[A, B] fp32 parameter -> same layout in the activation dtype (no copy for fp32)."""
    w2 = _c(w2.detach())
    if w2.dtype == dtype:
        return w2
    return _pack(w2.view(1, 1, w2.shape[0], w2.shape[1]), (0, 1, 2, 3), dtype).view(w2.shape)


def _reduce_ws(device, channels, quantities, extra=0):
    n = _lib.lib.eel_reduce_workspace_bytes(int(channels), int(quantities)) + extra
    return workspace(n, device), n


import threading

_TLS = threading.local()      # forward-scoped announcements of the calling thread (a model's forward runs on one thread)


def _stats_cols_ok(ncols):
    """fused BatchNorm statistics: the kernel's N tiles (256 wide when ncols allows) must divide its grid of 148 CTAs"""
    return ncols <= 256 or ncols in (512, 1024)


def _want_bn_sums(x, ncols):
    """a [2][ncols] fp32 buffer when the next op is a training-mode BatchNorm (see EELUnet._bn) and the producer's
    epilogue can accumulate its statistics"""
    if _bn_next() and _stats_cols_ok(ncols):
        return torch.empty((2, ncols), dtype=F32, device=x.device)
    return None


def expect_bn(flag):
    """the model announces that the tensor the next GEMM-class op produces goes straight into a training-mode BatchNorm"""
    _TLS.bn_next = bool(flag)


def _bn_next():
    return getattr(_TLS, "bn_next", False)


def _colsum(x2d, C):
    hit = _take(x2d, "_eel_colsum")
    if hit is not None and hit.numel() == C:
        return hit
    vec = 8 if x2d.dtype == BF16 else 4
    if C % vec != 0:
        # thin outputs (e.g. the 1-channel logits of Unet.final_conv): fold rows so a row is lcm(C, vec) wide
        import math
        L = C * vec // math.gcd(C, vec)
        if x2d.numel() % L != 0:
            raise _lib.EelError("colsum: %d elements do not fold into rows of %d" % (x2d.numel(), L))
        wide = torch.empty(L, dtype=F32, device=x2d.device)
        ws, n = _reduce_ws(x2d.device, L, 1)
        call("eel_colsum", ptr(x2d), ptr(wide), x2d.numel() // L, L, ptr(ws), n, dtype_code(x2d), stream())
        return wide.view(L // C, C).sum(0)
    out = torch.empty(C, dtype=F32, device=x2d.device)
    ws, n = _reduce_ws(x2d.device, C, 1)
    call("eel_colsum", ptr(x2d), ptr(out), x2d.numel() // C, C, ptr(ws), n, dtype_code(x2d), stream())
    return out


BF16 = torch.bfloat16



# --------------------------------------------------------------------------------------- batched weight packing
class WeightPacker:
    """Packs every tensor-core layer's fp32 weight into its two bf16 operand layouts (forward, data gradient)
    with ONE kernel launch per step (eel_pack_batch) instead of two small permute launches per layer.

    ``get(weight)`` -> (fwd_operand, dgrad_operand) or None when the parameter is not in the table (SIMT / fp32 layers
    pack on their own).  The table is rebuilt when a parameter moves (e.g. re-homed into the flat optimizer buffer)."""

    def __init__(self):
        self.entries = []      # (param, dims, perm_fwd, perm_dgrad)
        self.owners = []
        self.by_param = {}
        self.table = None
        self.ptrs = None
        self.versions = None
        self.split_ids = set()
        self.splits = []       # (param, dgrad operand, its de-interleaved copy)

    def add(self, param, dims, perm_fwd, perm_dgrad, owner=None):
        self.entries.append((param, tuple(int(d) for d in dims), perm_fwd, perm_dgrad))
        self.owners.append(owner)

    def want_split(self, param):
        """``param`` is the weight of a conv3x3 that reads a skip bridge: also keep its two operands with the input channels
        de-interleaved (BridgeConv3x3: eel_tc_conv3x3_2src, eel_tc_conv3x3_dgrad_split)"""
        if id(param) not in self.split_ids:
            self.split_ids.add(id(param))
            self.table = None               # rebuild: the split buffers are allocated with the table

    def stale(self):
        """a module's ``weight`` was replaced by another Parameter object (torch.nn.utils.prune.remove, manual surgery):
        the table must be rebuilt from the module tree"""
        return any(o is not None and getattr(o, "weight", None) is not e[0] for o, e in zip(self.owners, self.entries))

    def _build(self, device):
        import numpy as np
        rec = np.zeros((2 * len(self.entries), 8), dtype=np.int64)   # 64-byte records (eel_pack_job)
        self.by_param = {}
        keep = []
        for k, (w, dims, pf, pd) in enumerate(self.entries):
            outs = []
            for u, perm in enumerate((pf, pd)):
                dst = torch.empty([dims[i] for i in perm], dtype=BF16, device=device)
                r = rec[2 * k + u]
                r[0], r[1] = w.data_ptr(), dst.data_ptr()
                d32 = np.array(list(dims) + list(perm), dtype=np.int32)
                r[2:6] = d32.view(np.int64)
                outs.append(dst)
            self.by_param[id(w)] = tuple(outs)
            keep.append(w)
        self.splits = [(w, self.by_param[id(w)][0], torch.empty_like(self.by_param[id(w)][0]),
                        self.by_param[id(w)][1], torch.empty_like(self.by_param[id(w)][1]))
                       for (w, dims, pf, pd) in self.entries if id(w) in self.split_ids]
        self.table = torch.from_numpy(rec.reshape(-1).view(np.uint8).copy()).to(device)
        self.ptrs = [w.data_ptr() for w in keep]
        self.versions = None

    def refresh(self, device):
        """(re)pack if any weight changed since the last call; call at the start of every forward."""
        if not self.entries:
            return
        if self.table is None or self.table.device != device or self.ptrs != [e[0].data_ptr() for e in self.entries]:
            self._build(device)
        # torch-side in-place updates bump ``_version``; kernels that write parameters through raw pointers
        # (FusedAdam) announce themselves with ``weights_changed()``
        ver = [_WEIGHT_EPOCH] + [e[0]._version for e in self.entries]
        if ver == self.versions:
            return
        def launch():
            call("eel_pack_batch", ptr(self.table), 2 * len(self.entries), 128, stream())
            for w, fw, fsp, dg, dsp in self.splits:
                _deinterleave_operands(fw, fsp, dg, dsp)
        _on_side_stream(launch)
        self.versions = ver
        # publish on the parameters themselves: ops find the operands through the weight they are handed (forward and
        # backward, any thread, any number of models), and only while the weight is what was packed
        for e in self.entries:
            w = e[0]
            w._eel_packed = self.by_param[id(w)] + (_pack_key(w),)
        for w, fw, fsp, dg, dsp in self.splits:
            w._eel_packed_split = (fsp, dsp, _pack_key(w))

    def get(self, w):
        return self.by_param.get(id(w))


class FoldedPacker:
    """Inference only (eval-mode BatchNorm, autograd off): every tensor-core conv / ConvTranspose / to_space Linear that
    feeds a BatchNorm gets the BatchNorm folded into its packed bf16 weight (x gamma / sqrt(running_var + eps) per output
    channel) and an fp32 bias ((b - running_mean) * scale + beta); the ReLU that follows is the kernel's fused ReLU.
    Removes all 33 BatchNorm apply passes of a forward.  Two launches when anything changed (eel_bn_fold_batch,
    eel_pack_batch), none otherwise."""

    def __init__(self):
        self.entries = []      # (weight, bias, bn, dims, perm, scale_pos)
        self.owners = []
        self.by_weight = {}
        self.fold_table = self.pack_table = None
        self.ptrs = self.versions = None
        self.split_ids = set()     # weights of the convs that read a skip bridge: also kept with de-interleaved input channels
        self.split_by_weight = {}
        self.composed = ComposedPacker(BF16)       # (mlp[2], to_space, BatchNorm) triples: composed, then folded

    def bns(self):
        return [e[2] for e in self.entries] + self.composed.bns()

    def add(self, weight, bias, bn, dims, perm, scale_src_dim, owner=None):
        perm = tuple(perm)
        self.entries.append((weight, bias, bn, tuple(int(d) for d in dims), perm, perm.index(scale_src_dim)))
        self.owners.append(owner)

    def stale(self):
        return self.composed.stale() or any(
            o is not None and (getattr(o, "weight", None) is not e[0] or getattr(o, "bias", None) is not e[1])
            for o, e in zip(self.owners, self.entries))

    def _tensors(self, e):
        w, b, bn = e[0], e[1], e[2]
        return [w, b, bn.weight, bn.bias, bn.running_mean, bn.running_var]

    def _build(self, device):
        import numpy as np
        n = len(self.entries)
        fold = np.zeros((n, 8), dtype=np.int64)
        pack = np.zeros((n, 8), dtype=np.int64)
        self.by_weight, self._keep, self.split_by_weight = {}, [], {}
        for k, e in enumerate(self.entries):
            w, b, bn, dims, perm, spos = e
            C = bn.num_features
            scale = torch.empty(C, dtype=F32, device=device)
            fbias = torch.empty(C, dtype=F32, device=device)
            dst = torch.empty([dims[i] for i in perm], dtype=BF16, device=device)
            f = fold[k]
            f[0], f[1], f[2], f[3] = bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr()
            f[4] = b.data_ptr() if b is not None else 0
            f[5], f[6] = scale.data_ptr(), fbias.data_ptr()
            f[7] = int(np.array([C], dtype=np.int32).view(np.uint32)[0]) | (int(np.array([bn.eps], dtype=np.float32).view(np.uint32)[0]) << 32)
            r = pack[k]
            r[0], r[1] = w.data_ptr(), dst.data_ptr()
            r[2:6] = np.array(list(dims) + list(perm), dtype=np.int32).view(np.int64)
            r[6] = scale.data_ptr()
            r[7] = spos
            self.by_weight[id(w)] = (dst, fbias)
            self._keep.append((scale, fbias, dst))
            if id(w) in self.split_ids:
                self.split_by_weight[id(w)] = (dst, torch.empty_like(dst))
        to_dev = lambda a: torch.from_numpy(a.reshape(-1).view(np.uint8).copy()).to(device)
        self.fold_table, self.pack_table = to_dev(fold), to_dev(pack)
        self.ptrs = [t.data_ptr() for e in self.entries for t in self._tensors(e) if t is not None]
        self.versions = None

    def refresh(self, device):
        self.composed.refresh(device)
        if not self.entries:
            return
        ptrs = [t.data_ptr() for e in self.entries for t in self._tensors(e) if t is not None]
        if self.fold_table is None or self.fold_table.device != device or ptrs != self.ptrs:
            self._build(device)
        ver = [_WEIGHT_EPOCH, _STATS_EPOCH] + [t._version for e in self.entries for t in self._tensors(e) if t is not None]
        if ver == self.versions:
            return
        call("eel_bn_fold_batch", ptr(self.fold_table), len(self.entries), stream())
        call("eel_pack_batch", ptr(self.pack_table), len(self.entries), 128, stream())
        for fw, fsp in self.split_by_weight.values():          # [ky][kx][co][ci] -> ci columns de-interleaved
            call("eel_cols_deinterleave", ptr(fw), ptr(fsp), fw.shape[0] * fw.shape[1] * fw.shape[2], fw.shape[3], stream())
        self.versions = ver

    def want_split(self, weight):
        """conv3x3 that reads a skip bridge: keep its folded operand with de-interleaved input channels too (conv3x3_folded_2src)"""
        if id(weight) not in self.split_ids:
            self.split_ids.add(id(weight))
            self.fold_table = None

    def get_split(self, w):
        hit = self.split_by_weight.get(id(w))
        return None if hit is None else hit[1]

    def get(self, w):
        hit = self.by_weight.get(id(w))
        if hit is None:
            c = self.composed.get(w)
            if c is not None:
                hit = (c[0], c[2])
        return hit


class ComposedPacker:
    """ChannelAwarePatchedMLP ends in ``mlp[2]`` (Linear 256 -> Cout) followed directly by ``to_space`` (1x1 conv
    Cout -> Cout) (reference models/EELUnet.py:109-111,121-122): the hot path runs them as ONE GEMM with
    Wc = W_to_space W_mlp2 and bc = W_to_space b_mlp2 + b_to_space.  This table composes all of a model's pairs with one
    launch per step (eel_compose_batch, fp32 FFMA) into the forward [Cout][K] and data-gradient [K][Cout] operand layouts
    of the storage dtype; with ``bn`` given (inference) the eval-mode BatchNorm that follows is folded in as well.

    ``get(to_space.weight)`` -> (wc_fwd, wc_dgrad or None, bc)."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.entries = []      # (lin1, lin2, bn)
        self.by_weight = {}
        self.table = None
        self.ptrs = self.versions = self.params = None
        self.blocks = 1

    def add(self, lin1, lin2, bn=None):
        self.entries.append((lin1, lin2, bn))

    @staticmethod
    def _tensors(e):
        l1, l2, bn = e
        t = [l2.weight, l2.bias, l1.weight, l1.bias]
        if bn is not None:
            t += [bn.running_mean, bn.running_var, bn.weight, bn.bias]
        return t

    def stale(self):
        """a parameter object was replaced (prune.remove, manual surgery): rebuild from the module tree"""
        return self.params is not None and any(a is not b for e, p in zip(self.entries, self.params)
                                                for a, b in zip(self._tensors(e), p))

    def bns(self):
        return [e[2] for e in self.entries if e[2] is not None]

    def _build(self, device):
        import numpy as np
        n = len(self.entries)
        rec = np.zeros((n, 16), dtype=np.int64)       # 128-byte records (eel_compose_job)
        self.by_weight, self._keep = {}, []
        code = _lib.EEL_BF16 if self.dtype == BF16 else _lib.EEL_F32
        blocks = 1
        for k, e in enumerate(self.entries):
            l1, l2, bn = e
            cout, cmid, kin = l2.weight.shape[0], l1.weight.shape[0], l1.weight.shape[1]
            if l2.weight.shape[1] != cmid:
                raise _lib.EelError("composed linear pair: inner sizes differ (%d vs %d)" % (l2.weight.shape[1], cmid))
            fwd = torch.empty((cout, kin), dtype=self.dtype, device=device)
            dgr = torch.empty((kin, cout), dtype=self.dtype, device=device) if bn is None else None
            bc = torch.empty(cout, dtype=F32, device=device)
            r = rec[k]
            r[0], r[1], r[2], r[3] = l2.weight.data_ptr(), l2.bias.data_ptr(), l1.weight.data_ptr(), l1.bias.data_ptr()
            r[4], r[5], r[6] = fwd.data_ptr(), 0 if dgr is None else dgr.data_ptr(), bc.data_ptr()
            eps = 0.0
            if bn is not None:
                r[7], r[8], r[9], r[10] = bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr()
                eps = bn.eps
            r[11:13] = np.array([cout, cmid, kin, code], dtype=np.int32).view(np.int64)
            r[13] = int(np.array([eps], dtype=np.float32).view(np.uint32)[0])
            self.by_weight[id(l2.weight)] = (fwd, dgr, bc)
            self._keep.append((fwd, dgr, bc))
            blocks = max(blocks, ((cout + 63) // 64) * ((kin + 63) // 64 + 1))
        self.blocks = min(blocks, 128)
        self.table = torch.from_numpy(rec.reshape(-1).view(np.uint8).copy()).to(device)
        self.params = [self._tensors(e) for e in self.entries]
        self.ptrs = [t.data_ptr() for p in self.params for t in p]
        self.versions = None

    def refresh(self, device):
        if not self.entries:
            return
        ptrs = [t.data_ptr() for e in self.entries for t in self._tensors(e)]
        if self.table is None or self.table.device != device or ptrs != self.ptrs:
            self._build(device)
        ver = [_WEIGHT_EPOCH, _STATS_EPOCH] + [t._version for p in self.params for t in p]
        if ver == self.versions:
            return
        if any(e[2] is not None for e in self.entries):
            # (BatchNorm-folded inference tables: FoldedPacker reads them right away on the current stream)
            call("eel_compose_batch", ptr(self.table), len(self.entries), self.blocks, stream())
        else:
            _on_side_stream(lambda: call("eel_compose_batch", ptr(self.table), len(self.entries), self.blocks, stream()))
        self.versions = ver
        for e, p in zip(self.entries, self.params):
            if e[2] is None:                   # (BatchNorm-folded tables are inference-scoped: FoldedPacker hands them out)
                p[0]._eel_composed = self.by_weight[id(p[0])] + (self.dtype, tuple(_pack_key(t) for t in p))

    def get(self, w):
        return self.by_weight.get(id(w))


def build_composed(module, dtype):
    """table of every (mlp[2], to_space) pair of ``module`` (its ChannelAwarePatchedMLP blocks)"""
    import torch.nn as nn

    cp = ComposedPacker(dtype)
    for m in module.modules():
        mlp, ts = getattr(m, "mlp", None), getattr(m, "to_space", None)
        if isinstance(mlp, nn.Sequential) and len(mlp) == 3 and isinstance(mlp[2], nn.Linear) and isinstance(ts, nn.Conv2d):
            cp.add(mlp[2], ts)
    return cp


def set_folded(p):
    """inference forward of the calling thread: the FoldedPacker whose BatchNorm-folded weights the stages use (None: off)"""
    _TLS.folded = p


def folded(weight):
    """(packed bf16 weight with the BatchNorm folded in, fp32 folded bias) or None"""
    f = getattr(_TLS, "folded", None)
    return None if f is None else f.get(weight)


def folded_split(weight):
    """the BatchNorm-folded forward operand of ``weight`` with de-interleaved input channels, or None"""
    f = getattr(_TLS, "folded", None)
    return None if f is None else f.get_split(weight)


_IDENT = {}


def add2(a, b):
    """a + b on NHWC tensors (inference: the summed half of a skip bridge; eel_bn_add_fwd with identity constants)"""
    a, b = _c(a), _c(b)
    C = a.shape[-1]
    key = (a.device, C)
    cst = _IDENT.get(key)
    if cst is None:
        cst = _IDENT[key] = (torch.zeros(C, dtype=F32, device=a.device), torch.ones(C, dtype=F32, device=a.device))
    zero, one = cst
    out = torch.empty_like(a)
    call("eel_bn_add_fwd", ptr(a), ptr(b), ptr(out), a.numel() // C, C, ptr(zero), ptr(one), ptr(one), ptr(zero), dtype_code(a), stream())
    return out


def conv3x3_folded_2src(x1, x2, wk_split, fbias, relu):
    """inference: the conv that reads a skip bridge, its two halves as separate tensors (no interleaved tensor), + eval-mode
    BatchNorm (+ ReLU) as ONE tensor-core kernel"""
    x1, x2 = _c(x1), _c(x2)
    N, H, W, C1 = x1.shape
    C2 = x2.shape[-1]
    Cout = wk_split.shape[2]
    y = torch.empty((N, H, W, Cout), dtype=x1.dtype, device=x1.device)
    call("eel_tc_conv3x3_2src", ptr(x1), ptr(x2), ptr(wk_split), ptr(fbias), ptr(y), N, H, W, C1, C2, Cout, int(relu), None, stream())
    return y


def conv3x3_folded(x, wk, fbias, relu):
    """inference: conv3x3 + eval-mode BatchNorm (+ ReLU) as ONE tensor-core kernel"""
    x = _c(x)
    N, H, W, Cin = x.shape
    Cout = wk.shape[2]
    y = torch.empty((N, H, W, Cout), dtype=x.dtype, device=x.device)
    call("eel_tc_conv3x3", ptr(x), ptr(wk), ptr(fbias), ptr(y), N, H, W, Cin, Cout, int(relu), 0, None, stream())
    return y


def convt2x2_folded(x, wk, fbias):
    x = _c(x)
    N, h, w, Cin = x.shape
    Cout = wk.shape[2]
    y = torch.empty((N, 2 * h, 2 * w, Cout), dtype=x.dtype, device=x.device)
    call("eel_tc_convt2x2_fwd", ptr(x), ptr(wk), ptr(fbias), ptr(y), N, h, w, Cin, Cout, None, stream())
    return y


def linear_folded(x, w2, fbias, relu):
    x = _c(x)
    N, H, W, K = x.shape
    Nout = w2.shape[-2]
    y = torch.empty((N, H, W, Nout), dtype=x.dtype, device=x.device)
    call("eel_tc_linear", ptr(x), ptr(w2), ptr(fbias), ptr(y), N * H * W, K, Nout, int(relu), None, 0, 0, stream())
    return y


_WEIGHT_EPOCH = 0
_STATS_EPOCH = 0


def weights_changed():
    """To be called by anything that rewrites parameters behind torch's back: the fused Adam kernel does, and so must user
    code that updates weights through ``p.data`` (EMA swaps, clamping: ``p.data.mul_()`` bumps no version counter torch
    exposes).  Public as ``eel_unet_b200.invalidate_packed_weights()``.  Every packed / composed / folded copy is rebuilt at
    the next forward."""
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def running_stats_changed():
    """a training-mode BatchNorm kernel rewrote running_mean / running_var through raw pointers: the BatchNorm-folded
    inference weights (FoldedPacker, ComposedPacker with a BatchNorm) are stale.  Kept apart from the weight epoch, which a
    training forward must not bump halfway through (the packed operands published on the weights carry it)."""
    global _STATS_EPOCH
    _STATS_EPOCH += 1


def _pack_key(w):
    return (w._version, _WEIGHT_EPOCH, w.data_ptr())


def build_packer(module):
    """table of every tensor-core layer's weight of ``module`` -> (forward, data-gradient) bf16 operand layouts"""
    import torch.nn as nn

    pk = WeightPacker()
    composed = set()           # (mlp[2], to_space) pairs run as one composed matrix (ComposedPacker), never on their own
    for m in module.modules():
        mlp, ts = getattr(m, "mlp", None), getattr(m, "to_space", None)
        if isinstance(mlp, nn.Sequential) and len(mlp) == 3 and isinstance(mlp[2], nn.Linear) and isinstance(ts, nn.Conv2d):
            composed.update((id(mlp[2]), id(ts)))
    for m in module.modules():
        w = getattr(m, "weight", None)
        if not isinstance(w, nn.Parameter) or id(m) in composed:
            continue
        if isinstance(m, nn.ConvTranspose2d):
            ci, co = w.shape[0], w.shape[1]
            if ci % 64 == 0 and co % 64 == 0 and m.kernel_size == (2, 2):
                pk.add(w, w.shape, (2, 3, 1, 0), (0, 2, 3, 1), m)
        elif isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3):
            co, ci = w.shape[0], w.shape[1]
            if ci % 64 == 0 and co % 64 == 0:
                pk.add(w, w.shape, (2, 3, 0, 1), (2, 3, 1, 0), m)
        elif isinstance(m, nn.Linear) or (isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1)):
            no, k = w.shape[0], w.shape[1]
            if no % 64 == 0 and k % 64 == 0:
                pk.add(w, (no, k, 1, 1), (2, 3, 0, 1), (2, 3, 1, 0), m)      # [1][1][no][k] and [1][1][k][no], tiled transposes
    return pk


def _packed(weight, which):
    """operand ``which`` (0 forward, 1 data gradient) a WeightPacker published on ``weight``, or None when there is none or
    the weight changed since (the caller then packs on the fly)"""
    hit = getattr(weight, "_eel_packed", None)
    if hit is None or hit[2] != _pack_key(weight):
        return None
    _await_packed()
    return hit[which]


def _deinterleave_operands(fw, fsp, dg, dsp):
    """forward operand [ky][kx][co][ci] -> ci columns de-interleaved; data-gradient operand [ky][kx][ci][co] -> ci rows"""
    call("eel_cols_deinterleave", ptr(fw), ptr(fsp), fw.shape[0] * fw.shape[1] * fw.shape[2], fw.shape[3], stream())
    call("eel_rows_deinterleave", ptr(dg), ptr(dsp), dg.shape[0] * dg.shape[1], dg.shape[2], dg.shape[3] * dg.element_size(), stream())


def _packed_split(weight, dtype):
    """(forward, data-gradient) operands of ``weight`` with the input channels de-interleaved (even channels first): published by
    a WeightPacker, or made here"""
    hit = getattr(weight, "_eel_packed_split", None)
    if hit is not None and hit[2] == _pack_key(weight):
        _await_packed()
        return hit[0], hit[1]
    fw, dg = _packed(weight, 0), _packed(weight, 1)
    if fw is None:
        fw = _pack(weight, (2, 3, 0, 1), dtype)
    if dg is None:
        dg = _pack(weight, (2, 3, 1, 0), dtype)
    fsp, dsp = torch.empty_like(fw), torch.empty_like(dg)
    _deinterleave_operands(fw, fsp, dg, dsp)
    return fsp, dsp


def conv3x3(x, weight, bias, relu):
    """nn.Conv2d(k=3, pad=1): the 3-channel first layer of bf16 mode takes the im2col tensor-core path (StemConv)"""
    if not relu and StemConv.supported(x, weight) and bias is not None:
        return StemConv.apply(x, weight, bias)
    return Conv3x3.apply(x, weight, bias, relu)


def _tc_ok(x, *chans):
    """bf16 storage and every channel count a multiple of 64 -> tcgen05 tensor-core kernels (gemm_tc.cu)."""
    return x.dtype == BF16 and all(c % 64 == 0 for c in chans)


def _shift(x, inverse):
    N, H, W, C = x.shape
    y = torch.empty_like(x)
    call("eel_shift_channels", ptr(x), ptr(y), N, H, W, C, int(inverse), dtype_code(x), stream())
    return y


def nchw_to_nhwc(x, dtype):
    """fp32 NCHW model input -> NHWC activations (no gradient: the image is a leaf input)."""
    x = _c(x.detach().to(F32))
    N, C, H, W = x.shape
    y = torch.empty((N, H, W, C), dtype=dtype, device=x.device)
    call("eel_nchw_to_nhwc", ptr(x), ptr(y), N, C, H, W, dtype_code(y), stream())
    return y


# --------------------------------------------------------------------------------------- first conv (3 -> 64), bf16
class StemConv(Function):
    """The first nn.Conv2d(3, 64, 3, padding=1) (reference models/EELUnet.py:338) in bf16 mode: a compact im2col
    ([P][32], saved for the weight gradient) feeds the tensor-core GEMM / weight-gradient kernels (csrc/stem.cu).
    The input image is a leaf: there is no data gradient."""

    @staticmethod
    def supported(x, weight):
        N, H, W, Cin = x.shape
        return x.dtype == BF16 and Cin == 3 and tuple(weight.shape) == (64, 3, 3, 3) and (N * H * W) % 2 == 0

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        N, H, W, _ = x.shape
        P = N * H * W
        dev, st = x.device, stream()
        col = torch.empty((P, 32), dtype=BF16, device=dev)
        call("eel_stem_im2col", ptr(x), ptr(col), N, H, W, st)
        wblk = torch.empty((128, 64), dtype=BF16, device=dev)
        bias2 = torch.empty(128, dtype=F32, device=dev)
        call("eel_stem_pack", ptr(_c(weight.detach())), ptr(bias.detach()), ptr(wblk), ptr(bias2), st)
        y = torch.empty((N, H, W, 64), dtype=BF16, device=dev)
        sums2 = torch.empty((2, 128), dtype=F32, device=dev) if _bn_next() else None
        call("eel_tc_linear", ptr(col), ptr(wblk), None if sums2 is not None else ptr(bias2), ptr(y), P // 2, 64, 128, 0, ptr(sums2),
             0, 0, st)
        if sums2 is not None:
            sums = torch.empty((2, 64), dtype=F32, device=dev)
            call("eel_stem_fold_sums", ptr(sums2), ptr(sums), st)
            _attach(y, "_eel_bn_sums", (sums, bias.detach()))
        ctx.save_for_backward(col)
        ctx.weight = weight
        ctx.bn_in = None
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        dy = _c(dy)
        P = col.shape[0]
        st = stream()
        on = _async_ok(ctx.weight)
        dw = _grad_out(ctx.weight)
        with _Wgrad(on, dy, col):
            st = stream()
            dwblk = torch.empty((128, 64), dtype=F32, device=dy.device)
            call("eel_tc_wgrad", ptr(dy), ptr(col), ptr(dwblk), P // 2, 128, 64, 64, 1, dwblk.numel(), 0, st)
            call("eel_stem_unpack_dw", ptr(dwblk), ptr(dw), st)
        db = _colsum(dy, 64)
        return None, dw, db


# --------------------------------------------------------------------------------------- conv 3x3
class Conv3x3(Function):
    """nn.Conv2d(k=3, pad=1) (reference models/EELUnet.py:338,341,351,257), optional fused ReLU."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = _c(x)
        N, H, W, Cin = x.shape
        Cout = weight.shape[0]
        y = torch.empty((N, H, W, Cout), dtype=x.dtype, device=x.device)
        if _tc_ok(x, Cin, Cout):
            wk = _packed(weight, 0)
            if wk is None:
                wk = _pack(weight, (2, 3, 0, 1), x.dtype)  # [ky][kx][co][ci]  (K-major B operand)
            sums = _want_bn_sums(x, Cout) if not relu else None
            # a training-mode BatchNorm follows: the bias cancels in it, z is stored without (one FADD + a load less per
            # output in the epilogue that bounds these kernels); eel_bn_stats_from_sums adds it to the running mean
            b = bias.detach() if bias is not None else None
            call("eel_tc_conv3x3", ptr(x), ptr(wk), None if sums is not None else ptr(b), ptr(y), N, H, W, Cin, Cout, int(relu), 0,
                 ptr(sums), stream())
            if sums is not None:
                _attach(y, "_eel_bn_sums", (sums, b))
        else:
            wp = _pack(weight, (2, 3, 1, 0), x.dtype)  # [ky][kx][ci][co]
            call("eel_conv3x3_fwd", ptr(x), ptr(wp), ptr(bias.detach()), ptr(y), N, H, W, Cin, Cout, int(relu), 0,
                 dtype_code(x), stream())
        ctx.relu = relu
        bn_in = _take(x, "_eel_bn_out")
        ctx.bn_in = bn_in if _tc_ok(x, Cin, Cout) else None
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = _c(dy)
        N, H, W, Cin = x.shape
        Cout = weight.shape[0]
        st = stream()
        if ctx.relu:
            dz = torch.empty_like(dy)
            call("eel_relu_bwd", ptr(y), ptr(dy), ptr(dz), dy.numel(), dtype_code(dy), st)
            dy = dz
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if _tc_ok(x, Cin, Cout):
                wk = _packed(weight, 1)
                if wk is None:
                    wk = _pack(weight, (2, 3, 1, 0), x.dtype)  # [ky][kx][ci][co]: K-major B of the transposed problem
                # Measured on B200 (batch 64): the fused epilogue costs +0.12 ms on the 64 -> 64 full-resolution layer and saves
                # the 0.20 ms reduction pass; on the 128-channel half-resolution layers cost and saving cancel in step time
                # (+0.09 / -0.10) but the two reads of (dy, z) leave the HBM-bound side of the ledger (EEL_BNSUMS_WIDE=0
                # restores the plain launch there); small maps (bottleneck) are free.
                bn_in, ctx.bn_in = ctx.bn_in, None      # (not a saved tensor: drop it here, or it lives as long as the graph does)
                if bn_in is not None and _stats_cols_ok(Cin) and ((Cin == 64 and _BNSUMS_64) or N * H * W <= 32768 or
                                                                  (Cin != 64 and _BNSUMS_WIDE)):
                    # x = relu(bn(z)) feeds only this conv: dx is that BatchNorm's whole upstream gradient, and its
                    # backward sums come out of this launch's epilogue
                    z, mean, rstd, gamma, beta, bn_relu = bn_in
                    sums = torch.empty((2, Cin), dtype=F32, device=x.device)
                    cws = workspace(16 * Cin, x.device, slot=1)
                    call("eel_tc_conv3x3_dgrad_bnsums", ptr(dy), ptr(wk), ptr(dx), N, H, W, Cout, Cin, ptr(z), ptr(mean), ptr(rstd),
                         ptr(gamma.detach()), ptr(beta.detach()), int(bn_relu), ptr(sums), ptr(cws), st)
                    _attach(dx, "_eel_bn_bwd_sums", (sums, z))
                else:
                    call("eel_tc_conv3x3", ptr(dy), ptr(wk), None, ptr(dx), N, H, W, Cout, Cin, 0, 1, None, st)
            else:
                wd = _pack(weight, (2, 3, 0, 1), x.dtype)  # [ky][kx][co][ci]
                call("eel_conv3x3_fwd", ptr(dy), ptr(wd), None, ptr(dx), N, H, W, Cout, Cin, 0, 1, dtype_code(x), st)
        on = _async_ok(weight)
        dw = _grad_out(weight)
        with _Wgrad(on, x, dy):
            st = stream()
            dwp = torch.empty((3, 3, Cin, Cout), dtype=F32, device=x.device)
            if _tc_ok(x, Cin, Cout) and (Cin == 64 or Cin % 128 == 0):
                call("eel_tc_conv3x3_wgrad", ptr(x), ptr(dy), ptr(dwp), N, H, W, Cin, Cout, st)
            else:
                call("eel_conv3x3_wgrad", ptr(x), ptr(dy), ptr(dwp), N, H, W, Cin, Cout, dtype_code(x), st)
            _pack(dwp, (3, 2, 0, 1), F32, out=dw)
        db = _colsum(dy, Cout)
        return dx, dw, db, None


# --------------------------------------------------------------------------------------- conv transpose
class ConvT2x2(Function):
    """nn.ConvTranspose2d(k=2, s=2) (reference models/EELUnet.py:364,371)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        N, h, w, Cin = x.shape
        Cout = weight.shape[1]
        y = torch.empty((N, 2 * h, 2 * w, Cout), dtype=x.dtype, device=x.device)
        if _tc_ok(x, Cin, Cout):
            wk = _packed(weight, 0)
            if wk is None:
                wk = _pack(weight, (2, 3, 1, 0), x.dtype)  # [ky][kx][co][ci]
            sums = torch.empty((2, Cout), dtype=F32, device=x.device) if (_bn_next() and _stats_cols_ok(4 * Cout)) else None
            b = bias.detach()
            call("eel_tc_convt2x2_fwd", ptr(x), ptr(wk), None if sums is not None else ptr(b), ptr(y), N, h, w, Cin, Cout, ptr(sums),
                 stream())
            if sums is not None:
                _attach(y, "_eel_bn_sums", (sums, b))
        else:
            wp = _pack(weight, (0, 2, 3, 1), x.dtype)  # [ci][ky][kx][co]
            call("eel_convt2x2_fwd", ptr(x), ptr(wp), ptr(bias.detach()), ptr(y), N, h, w, Cin, Cout, dtype_code(x), stream())
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        N, h, w, Cin = x.shape
        Cout = weight.shape[1]
        st = stream()
        dx = None
        if ctx.needs_input_grad[0]:
            wp = _packed(weight, 1)
            if wp is None:
                wp = _pack(weight, (0, 2, 3, 1), x.dtype)
            dx = torch.empty_like(x)
            gw = min(w, 128)
            if _tc_ok(x, Cin, Cout) and 128 % gw == 0 and w % gw == 0:
                call("eel_tc_convt2x2_dgrad", ptr(dy), ptr(wp), ptr(dx), N, h, w, Cin, Cout, st)
            else:
                call("eel_convt2x2_dgrad", ptr(dy), ptr(wp), ptr(dx), N, h, w, Cin, Cout, dtype_code(x), st)
        on = _async_ok(weight)
        dw = _grad_out(weight)
        with _Wgrad(on, x, dy):
            st = stream()
            dwp = torch.empty((Cin, 2, 2, Cout), dtype=F32, device=x.device)
            gw = min(w, 64)
            if _tc_ok(x, Cin, Cout) and Cin % 128 == 0 and 64 % gw == 0 and w % gw == 0:
                call("eel_tc_wgrad", ptr(x), ptr(dy), ptr(dwp), N * h * w, Cin, 4 * Cout, 4 * Cout, 1, dwp.numel(), w, st)
            else:
                call("eel_convt2x2_wgrad", ptr(x), ptr(dy), ptr(dwp), N, h, w, Cin, Cout, dtype_code(x), st)
            _pack(dwp, (0, 3, 1, 2), F32, out=dw)
        db = _colsum(dy, Cout)
        return dx, dw, db


# --------------------------------------------------------------------------------------- 1x1 conv / Linear
class Linear(Function):
    """nn.Linear / nn.Conv2d(k=1) on the channel axis (reference models/EELUnet.py:105-112).

    ``shift=True`` folds ShiftedChannel (reference :88-97) into the operand addressing; ``shift="pre"``: the input was already
    stored through the shift by its producer (BNAct shift_out) -- only the backward's adjoint scatter remains.
    weight is [Nout, K] or [Nout, K, 1, 1].
    """

    @staticmethod
    def forward(ctx, x, weight, bias, shift):
        x = _c(x)
        N, H, W, K = x.shape
        Nout = weight.shape[0]
        ctx.tc = _tc_ok(x, K, Nout)
        if shift == "pre" and not (ctx.tc and K % 128 == 0):
            raise _lib.EelError("Linear(shift='pre'): a pre-shifted input needs the tensor-core path with K a multiple of 128")
        w2 = _packed(weight, 0) if ctx.tc else None
        if w2 is None:
            w2 = _as_dtype2d(weight.view(Nout, K), x.dtype)
        w2 = w2.view(Nout, K)
        y = torch.empty((N, H, W, Nout), dtype=x.dtype, device=x.device)
        if ctx.tc:
            if shift and shift != "pre":       # "pre": the producer already stored x through the shift (BNAct shift_out)
                x = _shift(x, False)           # saved shifted: wgrad then needs no gather
            sums = _want_bn_sums(x, Nout)
            b = bias.detach()
            call("eel_tc_linear", ptr(x), ptr(w2), None if sums is not None else ptr(b), ptr(y), N * H * W, K, Nout, 0, ptr(sums), 0, 0,
                 stream())
            if sums is not None:
                _attach(y, "_eel_bn_sums", (sums, b))
        else:
            sh, sw = (H, W) if shift else (0, 0)
            call("eel_linear_fwd", ptr(x), ptr(w2), ptr(bias.detach()), ptr(y), N * H * W, K, Nout, sh, sw, dtype_code(x), stream())
        ctx.shift = bool(shift)
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        N, H, W, K = x.shape
        Nout = weight.shape[0]
        P = N * H * W
        sh, sw = (H, W) if (ctx.shift and not ctx.tc) else (0, 0)
        st = stream()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if ctx.tc:
                wt = _packed(weight, 1)
                if wt is None:
                    wt = _pack(weight.detach().view(1, 1, Nout, K), (0, 1, 3, 2), x.dtype)   # [K][Nout]
                fold = ctx.shift and K % 128 == 0          # the epilogue stores through the adjoint shift
                call("eel_tc_linear", ptr(dy), ptr(wt), None, ptr(dx), P, Nout, K, 0, None, H if fold else 0, W if fold else 0, st)
                if ctx.shift and not fold:
                    dx = _shift(dx, True)
            else:
                w2 = _as_dtype2d(weight.view(Nout, K), x.dtype)
                call("eel_linear_dgrad", ptr(dy), ptr(w2), ptr(dx), P, K, Nout, sh, sw, dtype_code(x), st)
        on = _async_ok(weight)
        dw = _grad_out(weight, (Nout, K))
        with _Wgrad(on, x, dy):
            st = stream()
            if ctx.tc and Nout % 128 == 0:
                call("eel_tc_wgrad", ptr(dy), ptr(x), ptr(dw), P, Nout, K, K, 1, dw.numel(), 0, st)      # D[m=nout][n=k]
            elif ctx.tc and K % 128 == 0:
                call("eel_tc_wgrad", ptr(x), ptr(dy), ptr(dw), P, K, Nout, 1, K, dw.numel(), 0, st)      # D[m=k][n=nout] -> dw[n][m]
            else:
                call("eel_linear_wgrad", ptr(x), ptr(dy), ptr(dw), P, K, Nout, sh, sw, dtype_code(x), st)
        db = _colsum(dy, Nout)
        return dx, dw.view(weight.shape), db, None


class _Pair:
    """the two layers of a composed pair as bare tensors (ComposedLinear outside a model's table)"""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class ComposedLinear(Function):
    """``to_space(mlp[2](x))`` of ChannelAwarePatchedMLP (reference models/EELUnet.py:109-111,121-122) as ONE GEMM:
    y = (W2 W1) x + (W2 b1 + b2), w1/b1 = mlp[2] ([Cmid, K]), w2/b2 = to_space ([Cout, Cmid(,1,1)]).  The backward runs one
    data-gradient and one weight-gradient GEMM for the composed matrix; the four parameter gradients of the reference
    layers follow exactly in weight space (eel_compose_linear_bwd)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x = _c(x)
        N, H, W, K = x.shape
        Cout, Cmid = w2.shape[0], w1.shape[0]
        hit = getattr(w2, "_eel_composed", None)
        _await_packed()
        if hit is not None and hit[3] == x.dtype and hit[4] == tuple(_pack_key(t) for t in (w2, b2, w1, b1)):
            hit = hit[:3]
        else:
            hit = None
        if hit is None:
            one = ComposedPacker(x.dtype)
            one.add(_Pair(w1.detach(), b1.detach()), _Pair(w2.detach().view(Cout, Cmid), b2.detach()))
            one.refresh(x.device)
            _await_packed()
            hit = one._keep[0]
        wc, wct, bc = hit
        ctx.tc = _tc_ok(x, K, Cout)
        y = torch.empty((N, H, W, Cout), dtype=x.dtype, device=x.device)
        if ctx.tc:
            sums = _want_bn_sums(x, Cout)
            call("eel_tc_linear", ptr(x), ptr(wc), None if sums is not None else ptr(bc), ptr(y), N * H * W, K, Cout, 0, ptr(sums), 0, 0,
                 stream())
            if sums is not None:
                _attach(y, "_eel_bn_sums", (sums, bc))
        else:
            call("eel_linear_fwd", ptr(x), ptr(wc), ptr(bc), ptr(y), N * H * W, K, Cout, 0, 0, dtype_code(x), stream())
        ctx.save_for_backward(x, w1, b1, w2, wc, wct)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, b1, w2, wc, wct = ctx.saved_tensors
        dy = _c(dy)
        N, H, W, K = x.shape
        Cout, Cmid = w2.shape[0], w1.shape[0]
        P = N * H * W
        st = stream()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if ctx.tc:
                call("eel_tc_linear", ptr(dy), ptr(wct), None, ptr(dx), P, Cout, K, 0, None, 0, 0, st)
            else:
                call("eel_linear_dgrad", ptr(dy), ptr(wc), ptr(dx), P, K, Cout, 0, 0, dtype_code(x), st)
        s = _colsum(dy, Cout)
        on = _async_ok(w1, b1, w2)
        dw2 = _grad_out(w2, (Cout, Cmid))
        dw1 = _grad_out(w1, (Cmid, K))
        db1 = _grad_out(b1)
        with _Wgrad(on, x, dy, s):
            st = stream()
            dwc = torch.empty((Cout, K), dtype=F32, device=x.device)
            if ctx.tc and Cout % 128 == 0:
                call("eel_tc_wgrad", ptr(dy), ptr(x), ptr(dwc), P, Cout, K, K, 1, dwc.numel(), 0, st)
            elif ctx.tc and K % 128 == 0:
                call("eel_tc_wgrad", ptr(x), ptr(dy), ptr(dwc), P, K, Cout, 1, K, dwc.numel(), 0, st)
            else:
                call("eel_linear_wgrad", ptr(x), ptr(dy), ptr(dwc), P, K, Cout, 0, 0, dtype_code(x), st)
            call("eel_compose_linear_bwd", ptr(dwc), ptr(s), ptr(_c(w2.detach())), ptr(_c(w1.detach())), ptr(b1.detach()),
                 ptr(dw2), ptr(dw1), ptr(db1), Cout, Cmid, K, st)
        return dx, dw1, db1, dw2.view(w2.shape), s


def mlp_chain_supported(t, w0, cout):
    """the fused token-MLP kernel (csrc/capmlp_tc.cu): bf16, Linear 64 -> 256 first, 256 / 512 / 1024 output channels, whole
    128-pixel tiles"""
    N, H, W, K = t.shape
    return (t.dtype == BF16 and K == 64 and tuple(w0.shape) == (256, 64) and cout in (256, 512, 1024) and (N * H * W) % 128 == 0
            and t.is_cuda)


def mlp_chain_folded(t, w0, b0, wc_folded, fbias, relu):
    """inference: mlp[0] -> GELU -> composed (mlp[2], to_space, eval-mode BatchNorm) [+ ReLU] as one kernel; no intermediate is kept"""
    t = _c(t)
    N, H, W, K = t.shape
    C = wc_folded.shape[-2]
    w0p = _packed(w0, 0)
    if w0p is None:
        w0p = _as_dtype2d(w0.view(256, 64), BF16)
    z = torch.empty((N, H, W, C), dtype=BF16, device=t.device)
    call("eel_tc_capmlp_fwd", ptr(t), ptr(w0p), ptr(b0.detach()), ptr(wc_folded), ptr(fbias), None, None, ptr(z), N * H * W, C, int(relu),
         None, stream())
    return z


class MlpChain(Function):
    """``to_space(mlp[2](GELU(mlp[0](u))))`` of ChannelAwarePatchedMLP (reference models/EELUnet.py:107-111,121-122) with the
    forward as ONE tensor-core kernel (eel_tc_capmlp_fwd): the 256-channel h and a = GELU(h) are written once for the backward
    and never read back in the forward.  Its outputs are bit-identical to Linear -> Gelu -> ComposedLinear, and the backward IS
    those three backwards in sequence (same kernels), with the same free bias gradients and BatchNorm hand-overs."""

    @staticmethod
    def forward(ctx, u, w0, b0, w1, b1, w2, b2):
        u = _c(u)
        N, H, W, K = u.shape
        Cout, Cmid = w2.shape[0], w1.shape[0]
        P = N * H * W
        hit = getattr(w2, "_eel_composed", None)
        _await_packed()
        if hit is not None and hit[3] == u.dtype and hit[4] == tuple(_pack_key(t) for t in (w2, b2, w1, b1)):
            wc, wct, bc = hit[:3]
        else:
            one = ComposedPacker(u.dtype)
            one.add(_Pair(w1.detach(), b1.detach()), _Pair(w2.detach().view(Cout, Cmid), b2.detach()))
            one.refresh(u.device)
            _await_packed()
            wc, wct, bc = one._keep[0]
        w0p = _packed(w0, 0)
        if w0p is None:
            w0p = _as_dtype2d(w0.view(256, 64), u.dtype)
        dev = u.device
        h = torch.empty((N, H, W, 256), dtype=u.dtype, device=dev)
        a = torch.empty((N, H, W, 256), dtype=u.dtype, device=dev)
        z = torch.empty((N, H, W, Cout), dtype=u.dtype, device=dev)
        sums = _want_bn_sums(u, Cout)
        call("eel_tc_capmlp_fwd", ptr(u), ptr(w0p), ptr(b0.detach()), ptr(wc), None if sums is not None else ptr(bc), ptr(h), ptr(a), ptr(z),
             P, Cout, 0, ptr(sums), stream())
        if sums is not None:
            _attach(z, "_eel_bn_sums", (sums, bc))
        ctx.save_for_backward(u, h, a, w0, w1, b1, w2, wc, wct)
        return z

    @staticmethod
    def backward(ctx, dz):
        u, h, a, w0, w1, b1, w2, wc, wct = ctx.saved_tensors
        dz = _c(dz)
        N, H, W, K = u.shape
        Cout, Cmid = w2.shape[0], w1.shape[0]
        P = N * H * W
        st = stream()
        dev = u.device
        # ---- composed layer (ComposedLinear.backward)
        da = torch.empty_like(a)
        call("eel_tc_linear", ptr(dz), ptr(wct), None, ptr(da), P, Cout, 256, 0, None, 0, 0, st)
        s = _colsum(dz, Cout)
        on = _async_ok(w0, w1, b1, w2)
        dw2 = _grad_out(w2, (Cout, Cmid))
        dw1 = _grad_out(w1, (Cmid, 256))
        db1 = _grad_out(b1)
        with _Wgrad(on, dz, a, s):
            sst = stream()
            dwc = torch.empty((Cout, 256), dtype=F32, device=dev)
            call("eel_tc_wgrad", ptr(dz), ptr(a), ptr(dwc), P, Cout, 256, 256, 1, dwc.numel(), 0, sst)
            call("eel_compose_linear_bwd", ptr(dwc), ptr(s), ptr(_c(w2.detach())), ptr(_c(w1.detach())), ptr(b1.detach()),
                 ptr(dw2), ptr(dw1), ptr(db1), Cout, Cmid, 256, sst)
        # ---- GELU (Gelu.backward); the column sums of dh are mlp[0]'s bias gradient
        dh = torch.empty_like(h)
        db0 = torch.empty(256, dtype=F32, device=dev)
        call("eel_gelu_bwd_colsum", ptr(h), ptr(da), ptr(dh), ptr(db0), h.numel(), 256, dtype_code(h), st)
        # ---- mlp[0] (Linear.backward)
        du = None
        if ctx.needs_input_grad[0]:
            w0t = _packed(w0, 1)
            if w0t is None:
                w0t = _pack(w0.detach().view(1, 1, 256, 64), (0, 1, 3, 2), u.dtype)
            du = torch.empty_like(u)
            call("eel_tc_linear", ptr(dh), ptr(w0t), None, ptr(du), P, 256, 64, 0, None, 0, 0, st)
        dw0 = _grad_out(w0, (256, 64))
        with _Wgrad(on, dh, u):
            call("eel_tc_wgrad", ptr(dh), ptr(u), ptr(dw0), P, 256, 64, 64, 1, dw0.numel(), 0, stream())
        return du, dw0.view(w0.shape), db0, dw1, db1, dw2.view(w2.shape), s


def _bn_statistics(z, running_mean, running_var, training, momentum, eps):
    """(mean, rstd) of a BatchNorm over z [.., C]: batch statistics (+ running-stat update) in training -- from the sums the
    tensor-core producer left behind when there are any -- else the running statistics"""
    C = z.shape[-1]
    P = z.numel() // C
    dev = z.device
    mean = torch.empty(C, dtype=F32, device=dev)
    rstd = torch.empty(C, dtype=F32, device=dev)
    st = stream()
    hit = _take(z, "_eel_bn_sums")
    sums, skipped_bias = hit if hit is not None else (None, None)
    if sums is not None and (not training or sums.shape[1] != C):
        raise _lib.EelError("BatchNorm sums were produced for a tensor that is not consumed by a matching training-mode BatchNorm")
    if training:
        running_stats_changed()
    if training and sums is not None:
        call("eel_bn_stats_from_sums", ptr(sums), P, C, ptr(mean), ptr(rstd), ptr(running_mean), ptr(running_var),
             float(momentum), float(eps), ptr(skipped_bias), st)
    elif training:
        ws, n = _reduce_ws(dev, C, 2)
        call("eel_bn_stats", ptr(z), P, C, ptr(mean), ptr(rstd), ptr(running_mean), ptr(running_var),
             float(momentum), float(eps), ptr(ws), n, dtype_code(z), st)
    else:
        call("eel_bn_eval_stats", ptr(running_mean), ptr(running_var), float(eps), ptr(mean), ptr(rstd), C, st)
    return mean, rstd


# --------------------------------------------------------------------------------------- BatchNorm (+ReLU)
class BNAct(Function):
    """nn.BatchNorm2d [+ nn.ReLU] (reference models/EELUnet.py:339-344,352-357,365,373,256)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, relu, momentum, eps, producer_bias=False,
                single_conv_consumer=False, shift_out=False):
        z = _c(z)
        C = z.shape[-1]
        P = z.numel() // C
        st = stream()
        mean, rstd = _bn_statistics(z, running_mean, running_var, training, momentum, eps)
        y = torch.empty_like(z)
        g, b = gamma.detach(), beta.detach()
        if shift_out:
            # the result is stored through ShiftedChannel for the to_patch Linear that consumes it (Linear(shift="pre")): that
            # Linear's data gradient comes back through the adjoint shift, so the backward below is the plain one
            N, H, W, _ = z.shape
            call("eel_bn_act_shift_fwd", ptr(z), ptr(y), ptr(mean), ptr(rstd), ptr(g), ptr(b), N, H, W, C, int(relu), dtype_code(z), st)
        else:
            call("eel_bn_act_fwd", ptr(z), ptr(y), ptr(mean), ptr(rstd), ptr(g), ptr(b), P, C, int(relu), dtype_code(z), st)
        ctx.relu, ctx.training, ctx.producer_bias = relu, training, producer_bias
        ctx.save_for_backward(z, mean, rstd, gamma, beta)
        # single_conv_consumer: y feeds exactly one conv3x3 or (as the edge feature) exactly one skip bridge
        if single_conv_consumer and z.dtype == BF16 and ctx.needs_input_grad[0]:
            _attach(y, "_eel_bn_out", (z, mean, rstd, gamma, beta, relu))
        return y

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, gamma, beta = ctx.saved_tensors
        dy = _c(dy)
        C = z.shape[-1]
        P = z.numel() // C
        dz = torch.empty_like(z)
        dzsum = torch.empty(C, dtype=F32, device=z.device) if ctx.producer_bias else None
        hit = _take(dy, "_eel_bn_bwd_sums")
        sums = hit[0] if hit is not None and hit[1] is z else None
        if sums is not None and sums.shape[1] == C:
            # the conv that consumed this BatchNorm's output accumulated {dbeta, dgamma} in its data-gradient epilogue
            call("eel_bn_act_bwd_apply", ptr(dy), ptr(z), ptr(mean), ptr(rstd), ptr(gamma.detach()), ptr(beta.detach()), ptr(sums),
                 ptr(dz), ptr(dzsum), P, C, int(ctx.relu), int(ctx.training), dtype_code(z), stream())
            dgamma, dbeta = sums[1], sums[0]
        else:
            dgamma, dbeta = _grad_out(gamma), _grad_out(beta)
            ws, n = _reduce_ws(z.device, C, 2, extra=8 * C)
            call("eel_bn_act_bwd", ptr(dy), ptr(z), ptr(mean), ptr(rstd), ptr(gamma.detach()), ptr(beta.detach()), ptr(dz),
                 ptr(dgamma), ptr(dbeta), ptr(dzsum), P, C, int(ctx.relu), int(ctx.training), ptr(ws), n, dtype_code(z), stream())
        if dzsum is not None:
            _attach(dz, "_eel_colsum", dzsum)
        return dz, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class BNReluPool(Function):
    """nn.BatchNorm2d -> nn.ReLU -> {skip tensor, nn.MaxPool2d(2)} at the end of an encoder stage (reference
    models/EELUnet.py:387-406): returns (a, pooled).  The backward combines the gradient from the skip bridge (da) and from
    the pool (dp) inside the BatchNorm backward passes -- no max-pool backward tensor, no gradient-accumulation pass."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, momentum, eps, producer_bias=False):
        z = _c(z)
        N, H, W, C = z.shape
        mean, rstd = _bn_statistics(z, running_mean, running_var, training, momentum, eps)
        a = torch.empty_like(z)
        pooled = torch.empty((N, H // 2, W // 2, C), dtype=z.dtype, device=z.device)
        vec = 8 if z.dtype == BF16 else 4
        amax = torch.empty((N * (H // 2) * (W // 2), C // vec), dtype=torch.int16, device=z.device)   # 2 bits per channel
        call("eel_bn_relu_pool_fwd", ptr(z), ptr(a), ptr(pooled), ptr(amax), ptr(mean), ptr(rstd), ptr(gamma.detach()),
             ptr(beta.detach()), N, H, W, C, dtype_code(z), stream())
        ctx.training, ctx.producer_bias = training, producer_bias
        ctx.save_for_backward(z, mean, rstd, gamma, beta, amax)
        return a, pooled

    @staticmethod
    def backward(ctx, da, dp):
        z, mean, rstd, gamma, beta, amax = ctx.saved_tensors
        N, H, W, C = z.shape
        da = torch.zeros_like(z) if da is None else _c(da)
        dp = torch.zeros((N, H // 2, W // 2, C), dtype=z.dtype, device=z.device) if dp is None else _c(dp)
        dz = torch.empty_like(z)
        dgamma, dbeta = _grad_out(gamma), _grad_out(beta)
        ws, n = _reduce_ws(z.device, C, 2, extra=8 * C)
        dzsum = torch.empty(C, dtype=F32, device=z.device) if ctx.producer_bias else None
        call("eel_bn_relu_pool_bwd", ptr(da), ptr(dp), ptr(z), ptr(amax), ptr(mean), ptr(rstd), ptr(gamma.detach()), ptr(beta.detach()),
             ptr(dz), ptr(dgamma), ptr(dbeta), ptr(dzsum), N, H, W, C, int(ctx.training), ptr(ws), n, dtype_code(z), stream())
        if dzsum is not None:
            _attach(dz, "_eel_colsum", dzsum)
        return dz, dgamma, dbeta, None, None, None, None, None, None


class Relu(Function):
    """standalone nn.ReLU (reference models/EELUnet.py:260)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        call("eel_relu_fwd", ptr(x), ptr(y), x.numel(), dtype_code(x), stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        call("eel_relu_bwd", ptr(y), ptr(dy), ptr(dx), dy.numel(), dtype_code(dy), stream())
        return dx


class Gelu(Function):
    """nn.GELU() (erf form; reference models/EELUnet.py:109).  producer_bias: x comes straight from a biased Linear,
    whose bias gradient (the column sums of dx) the backward kernel then delivers for free."""

    @staticmethod
    def forward(ctx, x, producer_bias=False):
        x = _c(x)
        y = torch.empty_like(x)
        call("eel_gelu_fwd", ptr(x), ptr(y), x.numel(), dtype_code(x), stream())
        ctx.save_for_backward(x)
        ctx.producer_bias = producer_bias
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        C = x.shape[-1]
        vec = 8 if x.dtype == BF16 else 4
        if ctx.producer_bias and C % vec == 0 and 256 % (C // vec) == 0:
            colsum = torch.empty(C, dtype=F32, device=x.device)
            call("eel_gelu_bwd_colsum", ptr(x), ptr(dy), ptr(dx), ptr(colsum), x.numel(), C, dtype_code(x), stream())
            _attach(dx, "_eel_colsum", colsum)
        else:
            call("eel_gelu_bwd", ptr(x), ptr(dy), ptr(dx), x.numel(), dtype_code(x), stream())
        return dx, None


class MaxPool2(Function):
    """nn.MaxPool2d(2) (reference models/EELUnet.py:391,396,401,406)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
        call("eel_maxpool2_fwd", ptr(x), ptr(y), N, H, W, C, dtype_code(x), stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        N, H, W, C = x.shape
        dx = torch.empty_like(x)
        call("eel_maxpool2_bwd", ptr(x), ptr(dy), ptr(dx), N, H, W, C, dtype_code(x), stream())
        return dx


class AddInterleave(Function):
    """torch.add + FeatureInterleaveBridge (reference models/EELUnet.py:422-426,132-141)."""

    @staticmethod
    def forward(ctx, a, b, e):
        a, b, e = _c(a), _c(b), _c(e)
        N, H, W, C = a.shape
        out = torch.empty((N, H, W, 2 * C), dtype=a.dtype, device=a.device)
        call("eel_add_interleave_fwd", ptr(a), ptr(b), ptr(e), ptr(out), N * H * W, C, None, None, None, None, dtype_code(a), stream())
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        N, H, W, C2 = dout.shape
        C = C2 // 2
        dab = torch.empty((N, H, W, C), dtype=dout.dtype, device=dout.device)
        de = torch.empty_like(dab)
        call("eel_add_interleave_bwd", ptr(dout), ptr(dab), ptr(de), N * H * W, C, dtype_code(dout), stream())
        return dab, dab, de


class BNAddInterleave(Function):
    """nn.BatchNorm2d (no ReLU: the end of an upconv block, reference models/EELUnet.py:365,373) + torch.add +
    FeatureInterleaveBridge (:422-426) in one pass: the normalised tensor is never written.  Inputs: the pre-BatchNorm
    tensor z, the edge feature b and the encoder skip e."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, momentum, eps, b, e, producer_bias):
        z, b, e = _c(z), _c(b), _c(e)
        N, H, W, C = z.shape
        mean, rstd = _bn_statistics(z, running_mean, running_var, training, momentum, eps)
        out = torch.empty((N, H, W, 2 * C), dtype=z.dtype, device=z.device)
        call("eel_add_interleave_fwd", ptr(z), ptr(b), ptr(e), ptr(out), N * H * W, C, ptr(mean), ptr(rstd), ptr(gamma.detach()),
             ptr(beta.detach()), dtype_code(z), stream())
        ctx.training, ctx.producer_bias = training, producer_bias
        # b = relu(bn(z_b)) with this bridge as its only consumer (the model says so through BNAct's single-consumer flag):
        # the backward pass below then also delivers THAT BatchNorm's backward sums
        ctx.b_bn = _take(b, "_eel_bn_out") if ctx.needs_input_grad[8] else None
        ctx.save_for_backward(z, mean, rstd, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, mean, rstd, gamma, beta = ctx.saved_tensors
        dout = _c(dout)
        N, H, W, C2 = dout.shape
        C = C2 // 2
        P = N * H * W
        dab = torch.empty((N, H, W, C), dtype=dout.dtype, device=dout.device)
        de = torch.empty_like(dab)
        st = stream()
        dz = torch.empty_like(z)
        dzsum = torch.empty(C, dtype=F32, device=z.device) if ctx.producer_bias else None
        g, bt = gamma.detach(), beta.detach()
        if not _BRIDGE_BNSUMS:
            call("eel_add_interleave_bwd", ptr(dout), ptr(dab), ptr(de), P, C, dtype_code(dout), st)
            dgamma, dbeta = _grad_out(gamma), _grad_out(beta)
            ws, n = _reduce_ws(z.device, C, 2, extra=8 * C)
            call("eel_bn_act_bwd", ptr(dab), ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(bt), ptr(dz),
                 ptr(dgamma), ptr(dbeta), ptr(dzsum), P, C, 0, int(ctx.training), ptr(ws), n, dtype_code(z), st)
        else:
            # the de-interleaving pass also accumulates this BatchNorm's backward sums (and those of the BatchNorm + ReLU that
            # produced b, when the bridge is its only consumer): the BatchNorm backward is then a single apply pass
            sums = torch.empty((2, C), dtype=F32, device=z.device)
            other, ctx.b_bn = ctx.b_bn, None            # (not a saved tensor: drop it here)
            sums1 = torch.empty((2, C), dtype=F32, device=z.device) if other is not None else None
            z1, m1, r1, g1, b1, relu1 = other if other is not None else (None, None, None, None, None, 0)
            ws, n = _reduce_ws(z.device, C, 4)
            call("eel_add_interleave_bwd_bnsums", ptr(dout), ptr(dab), ptr(de), P, C, ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(bt), 0,
                 ptr(sums), ptr(z1), ptr(m1), ptr(r1), ptr(g1.detach()) if g1 is not None else None,
                 ptr(b1.detach()) if b1 is not None else None, int(relu1), ptr(sums1), ptr(ws), n, dtype_code(dout), st)
            call("eel_bn_act_bwd_apply", ptr(dab), ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(bt), ptr(sums), ptr(dz), ptr(dzsum), P, C, 0,
                 int(ctx.training), dtype_code(z), st)
            dgamma, dbeta = sums[1], sums[0]
            if other is not None:
                _attach(dab, "_eel_bn_bwd_sums", (sums1, z1))
        if dzsum is not None:
            _attach(dz, "_eel_colsum", dzsum)
        return dz, dgamma, dbeta, None, None, None, None, None, dab, de, None


class BridgeConv3x3(Function):
    """The skip bridge and the decoder block's first conv3x3 as ONE autograd node (bf16 tensor-core path):

        x = interleave(BatchNorm(z) + b, e)      (models/EELUnet.py:365/373, :422-426, :132-141)
        y = conv3x3(x)                           (:338 of the decoder block; pre-BatchNorm, statistics from the epilogue)

    The interleaved 2C-channel tensor x is never built.  Forward: s = BatchNorm(z) + b (eel_bn_add_fwd), then the conv reads
    its K chunks from the TWO tensors s and e through two tensor maps, on a forward operand whose input-channel columns were
    de-interleaved (eel_tc_conv3x3_2src).  Backward, 128 channels and up: the data gradient runs on the operand with de-interleaved
    rows, so its epilogue stores d(s) and d(e) as two tensors and accumulates the BatchNorm's backward sums over the first
    (eel_tc_conv3x3_dgrad_split) -- no interleaved gradient tensor, no de-interleaving pass, no BatchNorm reduction pass.  At 64
    channels (the full-resolution bridge) the data gradient is all epilogue, so there the plain launch writes the interleaved
    gradient and the de-interleaving pass delivers the sums (eel_add_interleave_bwd_bnsums), also those of the BatchNorm + ReLU
    that produced b when this bridge is its only consumer.  The weight gradient is two half launches (inputs s and e) whose
    results eel_dw_interleave writes into the reference layout."""

    # Measured on B200 (batch 64): against the plain data gradient + eel_add_interleave_bwd_bnsums the split launch saves
    # 0.13 / 0.07 / 0.01 ms at 128 / 256 / 512 channels, but LOSES 0.3 ms at the 64-channel full-resolution bridge, whose
    # data gradient (K = 64 per tap) is all epilogue: 0.44 -> 1.08 ms with the BatchNorm sums on half of its columns.
    MIN_SPLIT_CHANNELS = 128

    @staticmethod
    def supported(z, weight):
        C = z.shape[-1]
        return (z.dtype == BF16 and z.is_cuda and C % 64 == 0 and (C == 64 or C % 128 == 0) and 256 % (C // 8) == 0
                and tuple(weight.shape[1:]) == (2 * C, 3, 3) and weight.shape[0] % 64 == 0 and _stats_cols_ok(C))

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, momentum, eps, b, e, producer_bias, weight, bias):
        z, b, e = _c(z), _c(b), _c(e)
        N, H, W, C = z.shape
        Cout = weight.shape[0]
        st = stream()
        mean, rstd = _bn_statistics(z, running_mean, running_var, training, momentum, eps)
        s_t = torch.empty_like(z)
        call("eel_bn_add_fwd", ptr(z), ptr(b), ptr(s_t), N * H * W, C, ptr(mean), ptr(rstd), ptr(gamma.detach()), ptr(beta.detach()),
             dtype_code(z), st)
        wf, _ = _packed_split(weight, z.dtype)
        y = torch.empty((N, H, W, Cout), dtype=z.dtype, device=z.device)
        sums = _want_bn_sums(z, Cout)
        bd = bias.detach() if bias is not None else None
        call("eel_tc_conv3x3_2src", ptr(s_t), ptr(e), ptr(wf), None if sums is not None else ptr(bd), ptr(y), N, H, W, C, C, Cout, 0,
             ptr(sums), st)
        if sums is not None:
            _attach(y, "_eel_bn_sums", (sums, bd))
        ctx.training, ctx.producer_bias = training, producer_bias
        # 64 channels: b = relu(bn(z_b)) with this bridge as its only consumer -> the de-interleaving pass of the backward also
        # delivers THAT BatchNorm's sums (as BNAddInterleave)
        ctx.b_bn = _take(b, "_eel_bn_out") if (ctx.needs_input_grad[8] and C < BridgeConv3x3.MIN_SPLIT_CHANNELS) else None
        ctx.save_for_backward(z, mean, rstd, gamma, beta, s_t, e, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, gamma, beta, s_t, e, weight = ctx.saved_tensors
        dy = _c(dy)
        N, H, W, C = z.shape
        Cout = weight.shape[0]
        P = N * H * W
        st = stream()
        dev = z.device
        g, bt = gamma.detach(), beta.detach()
        dab = torch.empty((N, H, W, C), dtype=z.dtype, device=dev)
        de = torch.empty_like(dab)
        sums = torch.empty((2, C), dtype=F32, device=dev)
        if C >= BridgeConv3x3.MIN_SPLIT_CHANNELS:
            # data gradient in two halves + the BatchNorm's backward sums over the first
            _, wsp = _packed_split(weight, z.dtype)
            cws = workspace(16 * C, dev, slot=1)
            call("eel_tc_conv3x3_dgrad_split", ptr(dy), ptr(wsp), ptr(dab), ptr(de), N, H, W, Cout, 2 * C, ptr(z), ptr(mean), ptr(rstd),
                 ptr(g), ptr(bt), 0, ptr(sums), ptr(cws), st)
        else:
            wk = _packed(weight, 1)
            if wk is None:
                wk = _pack(weight, (2, 3, 1, 0), z.dtype)
            dx = torch.empty((N, H, W, 2 * C), dtype=z.dtype, device=dev)
            call("eel_tc_conv3x3", ptr(dy), ptr(wk), None, ptr(dx), N, H, W, Cout, 2 * C, 0, 1, None, st)
            other, ctx.b_bn = ctx.b_bn, None
            sums1 = torch.empty((2, C), dtype=F32, device=dev) if other is not None else None
            z1, m1, r1, g1, b1, relu1 = other if other is not None else (None, None, None, None, None, 0)
            ws, n = _reduce_ws(dev, C, 4)
            call("eel_add_interleave_bwd_bnsums", ptr(dx), ptr(dab), ptr(de), P, C, ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(bt), 0,
                 ptr(sums), ptr(z1), ptr(m1), ptr(r1), ptr(g1.detach()) if g1 is not None else None,
                 ptr(b1.detach()) if b1 is not None else None, int(relu1), ptr(sums1), ptr(ws), n, dtype_code(dx), st)
            if other is not None:
                _attach(dab, "_eel_bn_bwd_sums", (sums1, z1))
        # weight gradient (side stream): one half launch per input tensor, interleaved into the reference layout
        on = _async_ok(weight)
        dw = _grad_out(weight)
        with _Wgrad(on, s_t, e, dy):
            sst = stream()
            dwp0 = torch.empty((3, 3, C, Cout), dtype=F32, device=dev)
            dwp1 = torch.empty((3, 3, C, Cout), dtype=F32, device=dev)
            call("eel_tc_conv3x3_wgrad", ptr(s_t), ptr(dy), ptr(dwp0), N, H, W, C, Cout, sst)
            call("eel_tc_conv3x3_wgrad", ptr(e), ptr(dy), ptr(dwp1), N, H, W, C, Cout, sst)
            call("eel_dw_interleave", ptr(dwp0), ptr(dwp1), ptr(dw), 9, C, Cout, sst)
        db = _colsum(dy, Cout)
        # the BatchNorm's backward: one apply pass
        dz = torch.empty_like(z)
        dzsum = torch.empty(C, dtype=F32, device=dev) if ctx.producer_bias else None
        call("eel_bn_act_bwd_apply", ptr(dab), ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(bt), ptr(sums), ptr(dz), ptr(dzsum), P, C, 0,
             int(ctx.training), dtype_code(z), st)
        if dzsum is not None:
            _attach(dz, "_eel_colsum", dzsum)
        return dz, sums[1], sums[0], None, None, None, None, None, dab, de, None, dw, db


class Concat(Function):
    """torch.concat((a, b), dim=1) of the reference (models/Unet.py:78,83,88,93) on NHWC tensors."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        N, H, W, Ca = a.shape
        Cb = b.shape[-1]
        out = torch.empty((N, H, W, Ca + Cb), dtype=a.dtype, device=a.device)
        P, st = N * H * W, stream()
        call("eel_copy_cols", ptr(a), Ca, 0, ptr(out), Ca + Cb, 0, P, Ca, dtype_code(a), st)
        call("eel_copy_cols", ptr(b), Cb, 0, ptr(out), Ca + Cb, Ca, P, Cb, dtype_code(a), st)
        ctx.split = (Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        Ca, Cb = ctx.split
        N, H, W, _ = dout.shape
        da = torch.empty((N, H, W, Ca), dtype=dout.dtype, device=dout.device)
        db = torch.empty((N, H, W, Cb), dtype=dout.dtype, device=dout.device)
        P, st = N * H * W, stream()
        call("eel_copy_cols", ptr(dout), Ca + Cb, 0, ptr(da), Ca, 0, P, Ca, dtype_code(dout), st)
        call("eel_copy_cols", ptr(dout), Ca + Cb, Ca, ptr(db), Cb, 0, P, Cb, dtype_code(dout), st)
        return da, db


class ToNCHW(Function):
    """NHWC activations -> fp32 NCHW (the layout / dtype the reference's callers receive)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty((N, C, H, W), dtype=F32, device=x.device)
        call("eel_permute4", ptr(x), dtype_code(x), ptr(y), dtype_code(y), N, H, W, C, 0, 3, 1, 2, stream())
        ctx.dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy.to(F32))
        N, C, H, W = dy.shape
        dx = torch.empty((N, H, W, C), dtype=ctx.dtype, device=dy.device)
        call("eel_permute4", ptr(dy), dtype_code(dy), ptr(dx), dtype_code(dx), N, C, H, W, 0, 2, 3, 1, stream())
        return dx


class PGR(Function):
    """PredictionGuidedRefinement (reference models/EELUnet.py:200-203).  Returns (x*(1+s), s[N,1,H,W] fp32)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty_like(x)
        s = torch.empty((N, 1, H, W), dtype=F32, device=x.device)
        call("eel_pgr_fwd", ptr(x), ptr(_c(weight.detach())), ptr(bias.detach()), ptr(y), ptr(s), N * H * W, C,
             dtype_code(x), stream())
        ctx.save_for_backward(x, s, weight)
        return y, s

    @staticmethod
    def backward(ctx, dy, ds):
        x, s, weight = ctx.saved_tensors
        N, H, W, C = x.shape
        dy = _c(dy)
        ds = _c(ds.to(F32)) if ds is not None else None
        dx = torch.empty_like(x)
        dw = torch.empty(C, dtype=F32, device=x.device)
        db = torch.empty(1, dtype=F32, device=x.device)
        n = 4 * (2 * _lib_sms() + 1) * (C + 1)
        ws = workspace(n, x.device)
        call("eel_pgr_bwd", ptr(x), ptr(s), ptr(_c(weight.detach())), ptr(dy), ptr(ds), ptr(dx), ptr(dw), ptr(db),
             N * H * W, C, ptr(ws), n, dtype_code(x), stream())
        return dx, dw.view(weight.shape), db


class BNReluPGR(Function):
    """nn.BatchNorm2d + nn.ReLU (the end of a decoder block, reference models/EELUnet.py:343-344 / 356-357) followed by
    PredictionGuidedRefinement (:200-203) in one pass over the pre-BatchNorm tensor z.  The backward recomputes relu(bn(z)),
    and its PGR pass also produces the BatchNorm's two reduction sums, so BatchNorm backward is a single apply pass.
    Returns (x * (1 + s), s[N,1,H,W] fp32)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, momentum, eps, weight, bias, producer_bias):
        z = _c(z)
        N, H, W, C = z.shape
        mean, rstd = _bn_statistics(z, running_mean, running_var, training, momentum, eps)
        y = torch.empty_like(z)
        s = torch.empty((N, 1, H, W), dtype=F32, device=z.device)
        call("eel_bn_pgr_fwd", ptr(z), ptr(mean), ptr(rstd), ptr(gamma.detach()), ptr(beta.detach()), ptr(_c(weight.detach())),
             ptr(bias.detach()), ptr(y), ptr(s), N * H * W, C, dtype_code(z), stream())
        ctx.training, ctx.producer_bias = training, producer_bias
        ctx.save_for_backward(z, mean, rstd, gamma, beta, s, weight)
        return y, s

    @staticmethod
    def backward(ctx, dy, ds):
        z, mean, rstd, gamma, beta, s, weight = ctx.saved_tensors
        N, H, W, C = z.shape
        P = N * H * W
        dy = _c(dy)
        ds = _c(ds.to(F32)) if ds is not None else None
        dev = z.device
        dx = torch.empty_like(z)
        dw = torch.empty(C, dtype=F32, device=dev)
        db = torch.empty(1, dtype=F32, device=dev)
        sums = torch.empty((2, C), dtype=F32, device=dev)          # {dbeta, dgamma} of the BatchNorm
        n = 4 * (2 * _lib_sms()) * (3 * C + 1)
        ws = workspace(n, dev)
        st = stream()
        g, b = gamma.detach(), beta.detach()
        call("eel_bn_pgr_bwd", ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(s), ptr(_c(weight.detach())), ptr(dy), ptr(ds),
             ptr(dx), ptr(dw), ptr(db), ptr(sums), P, C, ptr(ws), n, dtype_code(z), st)
        dz = torch.empty_like(z)
        dzsum = torch.empty(C, dtype=F32, device=dev) if ctx.producer_bias else None
        call("eel_bn_act_bwd_apply", ptr(dx), ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(sums), ptr(dz), ptr(dzsum), P, C, 1,
             int(ctx.training), dtype_code(z), st)
        if dzsum is not None:
            _attach(dz, "_eel_colsum", dzsum)
        return dz, sums[1], sums[0], None, None, None, None, None, dw.view(weight.shape), db, None


def bn_pgr_supported(z):
    """the fused kernel keeps at most two channel vectors per lane (C <= 512 in bf16, 256 in fp32)"""
    vec = 8 if z.dtype == BF16 else 4
    C = z.shape[-1]
    return C % vec == 0 and C // vec <= 64 and (C // vec) & (C // vec - 1) == 0


def _lib_sms():
    """the SM count the library sizes its persistent grids for (workspace formulas below follow it)"""
    return int(_lib.lib.eel_num_sms())


class Head(Function):
    """final = LayerNorm(channels_first) + conv1x1, then sigmoid (reference models/EELUnet.py:217-225,330-333,467-469)."""

    @staticmethod
    def forward(ctx, x, lnw, lnb, weight, bias):
        x = _c(x)
        N, H, W, C = x.shape
        if C != 64:
            raise _lib.EelError("head kernel is specialised for the reference's 64-channel final stage")
        O = weight.shape[0]
        prob = torch.empty((N, O, H, W), dtype=F32, device=x.device)
        call("eel_head_fwd", ptr(x), ptr(lnw.detach()), ptr(lnb.detach()), ptr(_c(weight.detach())), ptr(bias.detach()),
             ptr(prob), N, H * W, O, dtype_code(x), stream())
        ctx.save_for_backward(x, lnw, lnb, weight, bias, prob)
        return prob

    @staticmethod
    def backward(ctx, dprob):
        x, lnw, lnb, weight, bias, prob = ctx.saved_tensors
        N, H, W, C = x.shape
        O = weight.shape[0]
        dprob = _c(dprob.to(F32))
        dev = x.device
        dx = torch.empty_like(x)
        dlnw, dlnb = _grad_out(lnw), _grad_out(lnb)
        dw, db = _grad_out(weight, (O, 64)), _grad_out(bias)
        n = 4 * (4 * _lib_sms() + 1) * ((2 + O) * 64 + O)
        ws = workspace(n, dev)
        call("eel_head_bwd", ptr(x), ptr(lnw.detach()), ptr(lnb.detach()), ptr(_c(weight.detach())), ptr(bias.detach()),
             ptr(prob), ptr(dprob), ptr(dx), ptr(dlnw), ptr(dlnb), ptr(dw), ptr(db), N, H * W, O, ptr(ws), n,
             dtype_code(x), stream())
        return dx, dlnw, dlnb, dw.view(weight.shape), db


class SE(Function):
    """ChannelAttention (reference models/EELUnet.py:57-80) on NHWC tokens t:[N,H,W,C].  producer_bias: t comes straight
    from a biased conv / linear (to_patch), whose bias gradient (column sums of dt) the backward then delivers for free."""

    @staticmethod
    def forward(ctx, t, w1, b1, w2, b2, producer_bias=False):
        t = _c(t)
        N, H, W, C = t.shape
        R = w1.shape[0]
        dev = t.device
        out = torch.empty_like(t)
        mean = torch.empty((N, C), dtype=F32, device=dev)
        att = torch.empty((N, C), dtype=F32, device=dev)
        hid = torch.empty((N, R), dtype=F32, device=dev)
        ws, n = _reduce_ws(dev, C, 1)
        call("eel_se_fwd", ptr(t), ptr(_c(w1.detach())), ptr(b1.detach()), ptr(_c(w2.detach())), ptr(b2.detach()), ptr(out),
             ptr(mean), ptr(att), ptr(hid), N, H * W, C, R, ptr(ws), n, dtype_code(t), stream())
        ctx.save_for_backward(t, mean, att, hid, w1, w2)
        ctx.producer_bias = producer_bias
        return out

    @staticmethod
    def backward(ctx, dout):
        t, mean, att, hid, w1, w2 = ctx.saved_tensors
        dout = _c(dout)
        N, H, W, C = t.shape
        R = w1.shape[0]
        dev = t.device
        dt = torch.empty_like(t)
        dw1, dw2 = _grad_out(w1, (R, C)), _grad_out(w2, (C, R))
        db1 = torch.empty(R, dtype=F32, device=dev)
        db2 = torch.empty(C, dtype=F32, device=dev)
        dtsum = torch.empty(C, dtype=F32, device=dev) if ctx.producer_bias else None
        ws, n = _reduce_ws(dev, C, 2, extra=20 * N * C)
        call("eel_se_bwd", ptr(t), ptr(dout), ptr(att), ptr(hid), ptr(mean), ptr(_c(w1.detach())), ptr(_c(w2.detach())),
             ptr(dt), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(dtsum), N, H * W, C, R, ptr(ws), n, dtype_code(t), stream())
        if dtsum is not None:
            _attach(dt, "_eel_colsum", dtsum)
        return dt, dw1.view(w1.shape), db1, dw2.view(w2.shape), db2, None


class HFT(Function):
    """HighFourierTransform (reference models/EELUnet.py:153-191) as an exact low-rank projection."""

    @staticmethod
    def forward(ctx, x, mask_range):
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty_like(x)
        # the unit phase z/|z| is only kept when a backward can follow (inference: up to two thirds of the last step's writes)
        phase = None
        if ctx.needs_input_grad[0]:
            phase = torch.empty(_lib.lib.eel_hft_phase_elems(N, H, W, C, int(mask_range), dtype_code(x)), dtype=x.dtype, device=x.device)
        n = _lib.lib.eel_hft_workspace_bytes(N, H, W, C, int(mask_range))
        ws = workspace(n, x.device, slot=1)
        call("eel_hft_fwd", ptr(x), ptr(y), ptr(phase), N, H, W, C, int(mask_range), ptr(ws), n, dtype_code(x), stream())
        ctx.mask_range = int(mask_range)
        ctx.save_for_backward(phase)
        return y

    @staticmethod
    def backward(ctx, dy):
        (phase,) = ctx.saved_tensors
        dy = _c(dy)
        N, H, W, C = dy.shape
        dx = torch.empty_like(dy)
        n = _lib.lib.eel_hft_workspace_bytes(N, H, W, C, ctx.mask_range)
        ws = workspace(n, dy.device, slot=1)
        call("eel_hft_bwd", ptr(dy), ptr(phase), ptr(dx), N, H, W, C, ctx.mask_range, ptr(ws), n, dtype_code(dy), stream())
        return dx, None


# --------------------------------------------------------------------------------------- loss
import ctypes as _ct


def _ptr_array(tensors):
    arr = (_ct.c_void_p * 6)()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t) if t is not None else None
    return arr


class EdgeBceDice(Function):
    """edge_BceDiceLoss.forward (reference utils/Loss.py:97-113) over (seg, edge_5..edge_1, target)."""

    @staticmethod
    def forward(ctx, seg, e5, e4, e3, e2, e1, target, wb, wd):
        preds = [_c(t.to(F32)) for t in (seg, e5, e4, e3, e2, e1)]
        target = _c(target.to(F32))
        N, H, W = target.shape[0], target.shape[-2], target.shape[-1]
        if seg.numel() != target.numel():
            raise _lib.EelError("edge_BceDiceLoss: seg and target must have the same number of elements (1 class)")
        dev = target.device
        loss = torch.empty((), dtype=F32, device=dev)
        sums = torch.empty((6, N, 4), dtype=torch.float64, device=dev)
        call("eel_edge_loss_fwd", _ptr_array(preds), ptr(target), N, H, W, float(wb), float(wd), ptr(loss), ptr(sums), stream())
        ctx.wb, ctx.wd = float(wb), float(wd)
        ctx.save_for_backward(target, sums, *preds)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        target, sums, *preds = ctx.saved_tensors
        N, H, W = target.shape[0], target.shape[-2], target.shape[-1]
        dloss = _c(dloss.to(F32))
        grads = [torch.empty_like(p) if ctx.needs_input_grad[i] else None for i, p in enumerate(preds)]
        call("eel_edge_loss_bwd", _ptr_array(preds), ptr(target), ptr(sums), ptr(dloss), _ptr_array(grads), N, H, W,
             ctx.wb, ctx.wd, stream())
        return (*grads, None, None, None)
