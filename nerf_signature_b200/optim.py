"""Optimizer for watermark training: torch.optim.Adam semantics over `NeRFNetwork.get_params(lr)`
(nerf/network_wtmk_tcnn.py:179-188, nerf/utils_wtmk_disen.py:1175-1181), with the message tables updated
by one fused kernel (csrc/optim.cu) from the single gradient G = dL/dS they all share.

With plain torch.optim.Adam (which stays fully supported — the encoder then fans G out to the selected
tables' .grad exactly like the reference's autograd does) a step costs message_dim 4 MiB gradient copies,
an unscale pass and a 7-stream multi-tensor Adam, and the set of parameters that have a gradient changes
with every message, which rules out CUDA-graph capture.  WatermarkAdam keeps G in one persistent buffer,
lets the kernel pick the tables from the device-side message, and is capture-safe; it implements the
`_step_supports_amp_scaling` contract of torch.amp.GradScaler (grad_scale / found_inf tensors), so
`scaler.step(optimizer)` / `scaler.update()` work unchanged, inf checks included.

Everything that is not a message table (the HiDDeN decoder) is delegated to a fused torch.optim.Adam.
"""
import torch

from . import _lib

_P = _lib.ptr


class WatermarkAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, capturable=False, grad_buffer=None, flat_bucket=None,
                 shard=None):
        """grad_buffer: optional pre-allocated [T*2] fp32 storage for G (the data-parallel harness passes a slice
        of its flat all-reduce bucket).
        flat_bucket: the WHOLE flat gradient bucket [dL/dS | gradients of every other trainable parameter, in get_params
        order] (parallel.GradSync.make_flat_buffer).  When given, the other parameters (the HiDDeN decoder) are re-pointed
        at ONE flat parameter buffer and updated by one Adam kernel (nsig_flat_adam_step) instead of torch's multi-tensor
        Adam, and FusedGradScaler can check the whole bucket for non-finite values in one pass.
        shard: (rank, world_size[, group]) - ZeRO-style sharding of the message tables' optimizer over a data-parallel job.
        The exchanged gradient G is identical on every rank, so rank r updates only the r-th contiguous slice of every
        selected table (and keeps moments for that slice only): the HBM-bound Adam pass costs 1/world.  The forward needs
        nothing but the summed table S, which each rank builds for its slice and all-gathers (4 MiB per step instead of
        broadcasting 2*message_dim tables).  A rank's copy of the OTHER slices goes stale: call gather_tables() before
        anything reads the tables themselves (checkpoints, single-rank evaluation)."""
        enc = model.msg_encoder
        tables = enc.tables()
        dev = tables[0].device
        if dev.type != "cuda":
            raise _lib.NsigError("WatermarkAdam needs the model on a CUDA device (no CPU path)")
        table_ids = {id(t) for t in tables}
        others = []
        self._train_tables = not getattr(model, "finetune_decoder", False)
        for group in model.get_params(lr):
            ps = [p for p in group["params"] if id(p) not in table_ids]
            if ps:
                others.append({"params": ps, "lr": group["lr"]})
        # G = dL/dS lives here; the proxy parameter only exists so that GradScaler's inf check
        # (_check_inf_per_device walks param_groups) sees G like any other gradient.
        self.G = torch.zeros_like(tables[0]) if grad_buffer is None else grad_buffer.view_as(tables[0])
        self._proxy = torch.nn.Parameter(torch.empty_like(tables[0]), requires_grad=True)
        self._proxy.grad = self.G
        groups = [{"params": g["params"], "lr": g["lr"]} for g in others]   # outer groups: what lr schedulers act on
        if self._train_tables:
            groups.append({"params": [self._proxy], "lr": lr, "msg_tables": True})
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps))
        # capturable: learning rates live in device scalars (tensor lr for the inner fused Adam, `_lr_dev` for the table
        # kernel) so that a per-step scheduler acting on param_groups[...]["lr"] still takes effect when the step is
        # replayed from a CUDA graph: call sync_lr() before each replay (harness.Scene does)
        self.capturable = capturable
        self.inner = None
        self.flat_bucket = flat_bucket
        self._flat = None
        if others and flat_bucket is not None and len(others) == 1 and self._train_tables:
            ps = others[0]["params"]
            n_other = sum(p.numel() for p in ps)
            ok = all(p.dtype == torch.float32 and p.is_contiguous() and p.requires_grad for p in ps)
            if ok and flat_bucket.numel() >= tables[0].numel() + n_other:
                flat_p = torch.empty(n_other, dtype=torch.float32, device=dev)
                o = 0
                for p in ps:   # the module keeps its parameters (names, shapes, state dict); their storage becomes one buffer
                    flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
                    p.data = flat_p[o:o + p.numel()].view(p.shape)
                    o += p.numel()
                self._flat = {"p": flat_p, "g": flat_bucket[tables[0].numel():tables[0].numel() + n_other],
                              "m": torch.zeros_like(flat_p), "v": torch.zeros_like(flat_p),
                              "step": torch.zeros(1, dtype=torch.float32, device=dev), "params": ps,
                              "lr_dev": torch.tensor([float(others[0]["lr"])], dtype=torch.float32, device=dev)}
                self.scaler_bumps_step = False   # FusedGradScaler increments _flat["step"] inside its check kernel
        if others and self._flat is None:
            inner_groups = [{"params": g["params"],
                             "lr": torch.tensor(float(g["lr"]), dtype=torch.float32, device=dev) if capturable else g["lr"]}
                            for g in others]
            self.inner = torch.optim.Adam(inner_groups, lr=lr, betas=betas, eps=eps, fused=True, capturable=capturable)
        self._lr_dev = torch.tensor([float(lr)], dtype=torch.float32, device=dev) if capturable else None
        self._lr_host = [float(g["lr"]) for g in self.param_groups]
        self.enc = enc
        self._model = model
        self.tables = tables
        self.shard = None
        if shard is not None and self._train_tables and int(shard[1]) > 1:
            rank, world = int(shard[0]), int(shard[1])
            total = tables[0].numel()
            if total % (8 * world):
                raise ValueError("table size must be a multiple of 8 * world_size to shard the optimizer")
            self.shard = (rank * (total // world), total // world, shard[2] if len(shard) > 2 else None)
            enc.shard = self.shard
        self.message = None  # device float [md]; set by the training step before backward
        if self._train_tables:
            enc.grad_sink = self.G
            n = len(tables)
            self.exp_avg = [torch.zeros_like(t) for t in tables]
            self.exp_avg_sq = [torch.zeros_like(t) for t in tables]
            self.steps = torch.zeros(n, dtype=torch.float32, device=dev)
            self._coef = torch.zeros(n, 2, dtype=torch.float32, device=dev)
            ptrs = [[t.data_ptr() for t in tables], [t.data_ptr() for t in self.exp_avg],
                    [t.data_ptr() for t in self.exp_avg_sq]]
            self._ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
            self._ptr_key = tuple(ptrs[0])

    def set_message(self, message_dev):
        self.message = message_dev

    def sync_lr(self):
        """Propagate the learning rates of the outer param_groups (what lr schedulers modify) to the inner Adam and, in
        capturable mode, to the device scalars the captured kernels read.  Host-side no-op while nothing changed."""
        n_inner = len(self.inner.param_groups) if self.inner is not None else 0
        for i, g in enumerate(self.param_groups):
            lr = float(g["lr"])
            changed = lr != self._lr_host[i]
            self._lr_host[i] = lr
            if self._flat is not None and not g.get("msg_tables"):
                if changed:
                    self._flat["lr_dev"].fill_(lr)
            elif i < n_inner:
                g_in = self.inner.param_groups[i]
                if isinstance(g_in["lr"], torch.Tensor):
                    if changed:
                        g_in["lr"].fill_(lr)
                else:
                    g_in["lr"] = lr
            elif g.get("msg_tables") and self._lr_dev is not None and changed:
                self._lr_dev.fill_(lr)

    @torch.no_grad()
    def gather_tables(self, moments=True):
        """Sharded mode: make every rank's message tables (and Adam moments) complete again by all-gathering the slices
        their owners updated.  Collective; 2*message_dim x 4 MiB (x3 with moments): for checkpoints and evaluation only."""
        if self.shard is None:
            return
        import torch.distributed as dist
        lo, n, group = self.shard
        sets = [self.tables] + ([self.exp_avg, self.exp_avg_sq] if moments else [])
        for tensors in sets:
            for t in tensors:
                flat = t.data.view(-1)
                dist.all_gather_into_tensor(flat, flat[lo:lo + n].clone(), group=group)
        self._model._S_cache = None

    # ---- checkpointing: the layout torch.optim.Adam(model.get_params(lr)).state_dict() has ---------------------------
    def state_dict(self):
        """Same structure as the reference's optimizer checkpoint (nerf/utils_wtmk_disen.py:1410, an Adam over
        get_params): parameter indices follow get_params order - the 2*message_dim message tables first, then the
        decoder - and state[i] = {step, exp_avg, exp_avg_sq} exists for every parameter that has been updated."""
        n_t = len(self.tables) if self._train_tables else 0
        groups, state = [], {}
        self.gather_tables()   # sharded mode: collective - every rank must call state_dict()
        if self._train_tables:
            g = {k: v for k, v in self.param_groups[-1].items() if k not in ("params", "msg_tables")}
            groups.append({**g, "params": list(range(n_t))})
            steps = self.steps.detach().cpu()
            for t in range(n_t):
                if float(steps[t]) > 0:
                    state[t] = {"step": steps[t].clone(), "exp_avg": self.exp_avg[t].detach().clone(),
                                "exp_avg_sq": self.exp_avg_sq[t].detach().clone()}
        if self._flat is not None:
            f, g0 = self._flat, self.param_groups[0]
            g = {k: v for k, v in g0.items() if k != "params"}
            groups.append({**g, "params": list(range(n_t, n_t + len(f["params"])))})
            if float(f["step"]) > 0:
                o = 0
                for j, p in enumerate(f["params"]):
                    n = p.numel()
                    state[n_t + j] = {"step": f["step"].detach().reshape(()).clone(),
                                      "exp_avg": f["m"][o:o + n].view(p.shape).clone(),
                                      "exp_avg_sq": f["v"][o:o + n].view(p.shape).clone()}
                    o += n
        if self.inner is not None:
            sd = self.inner.state_dict()
            for g in sd["param_groups"]:
                g = dict(g)
                g["lr"] = float(g["lr"])
                g["params"] = [i + n_t for i in g["params"]]
                groups.append(g)
            for k, v in sd["state"].items():
                state[k + n_t] = v
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Accepts state_dict() of this class and of a torch.optim.Adam built over the same get_params groups."""
        n_t = len(self.tables) if self._train_tables else 0
        groups, state = state_dict["param_groups"], state_dict["state"]
        if self._train_tables:
            tg = groups[0]
            if list(tg["params"]) != list(range(n_t)):
                raise ValueError(f"first param group must hold the {n_t} message tables (get_params order)")
            self.steps.zero_()
            for t in range(n_t):
                st = state.get(t)
                self.exp_avg[t].zero_(); self.exp_avg_sq[t].zero_()
                if st is not None:
                    self.exp_avg[t].copy_(st["exp_avg"]); self.exp_avg_sq[t].copy_(st["exp_avg_sq"])
                    self.steps[t] = float(st["step"])
            for k in ("lr", "betas", "eps"):
                if k in tg:
                    self.param_groups[-1][k] = float(tg[k]) if k != "betas" else tuple(tg[k])
            groups = groups[1:]
        if self._flat is not None:
            f = self._flat
            f["m"].zero_(); f["v"].zero_(); f["step"].zero_()
            o = 0
            for j, p in enumerate(f["params"]):
                st = state.get(n_t + j)
                n = p.numel()
                if st is not None:
                    f["m"][o:o + n].copy_(st["exp_avg"].reshape(-1)); f["v"][o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    f["step"].fill_(float(st["step"]))
                o += n
            if groups:
                self.param_groups[0]["lr"] = float(groups[0]["lr"])
        if self.inner is not None:
            # torch's Optimizer.load_state_dict keeps a tensor that already has the right dtype/device BY REFERENCE; a state dict
            # taken from a live optimizer (not from torch.load) would then share its moments with this one: copy them
            def _copy(v):
                return {kk: (vv.detach().clone() if isinstance(vv, torch.Tensor) else vv) for kk, vv in v.items()}
            sd = {"param_groups": [], "state": {k - n_t: _copy(v) for k, v in state.items() if k >= n_t}}
            for g, g_in, g_out in zip(groups, self.inner.param_groups, self.param_groups):
                g2 = dict(g)
                g2["params"] = [i - n_t for i in g["params"]]
                g_out["lr"] = float(g["lr"])
                if isinstance(g_in["lr"], torch.Tensor):   # keep the device scalar the captured step reads
                    g_in["lr"].fill_(float(g["lr"]))
                    g2["lr"] = g_in["lr"]
                sd["param_groups"].append(g2)
            self.inner.load_state_dict(sd)
        self._lr_host = [float("nan")] * len(self.param_groups)   # force a refresh of the device scalars
        self.sync_lr()

    def _refresh_decoder(self):
        fn = getattr(self._model, "refresh_decoder_weights", None)
        if fn is not None:
            fn()

    def zero_grad(self, set_to_none=True):
        if self.inner is not None:
            self.inner.zero_grad(set_to_none=set_to_none)
        # the field backward ACCUMULATES dL/dS into G (FieldConfig.S_sink); the proxy keeps pointing at it
        if self.flat_bucket is not None:
            self.flat_bucket.zero_()   # one fill: G and the flat-mode decoder gradients live in the same bucket
        else:
            self.G.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closures are not supported")
        self.step_decoder()
        self.step_tables()
        return None

    @torch.no_grad()
    def step_decoder(self):
        """The decoder's share of step(): Adam over the flat parameter vector (or torch's fused Adam over the decoder's
        tensors), on a side stream that is joined before returning, then the refresh of the fused decoder's fp16 weights."""
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        side = None
        self.sync_lr()
        if self._flat is not None:
            f, group = self._flat, self.param_groups[0]
            if not self.scaler_bumps_step:   # plain GradScaler / no scaler: count the step here (skipped steps do not count)
                f["step"].add_(1.0 if found_inf is None else (1.0 - found_inf.reshape(1)))
            side = _lib.side_stream(self.G.device, 1)
            b1, b2 = group["betas"]
            cur = torch.cuda.current_stream()
            if side is not None:
                side.wait_stream(cur)
            with torch.cuda.stream(side if side is not None else cur):   # next to the HBM-bound message-table Adam
                _lib.call("nsig_flat_adam_step", _P(f["p"]), _P(f["g"]), _P(f["m"]), _P(f["v"]), f["p"].numel(),
                          _P(f["step"]), _P(grad_scale), _P(found_inf), float(group["lr"]), _P(f["lr_dev"]), float(b1),
                          float(b2), float(group["eps"]))
                self._refresh_decoder()   # fp16 weight copies of the fused decoder, behind the update
        if self.inner is not None:
            self.inner.grad_scale, self.inner.found_inf = grad_scale, found_inf
            # the decoder's (tiny, latency-bound) Adam runs next to the HBM-bound message-table Adam
            side = _lib.side_stream(self.G.device, 1) if self._train_tables else None
            try:
                if side is not None:
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        self.inner.step()
                        self._refresh_decoder()
                else:
                    self.inner.step()
                    self._refresh_decoder()
            finally:
                del self.inner.grad_scale, self.inner.found_inf
        self._decoder_side = side

    def _table_call_args(self):
        if self.message is None:
            raise RuntimeError("WatermarkAdam.set_message(message) must be called before step()")
        if tuple(t.data_ptr() for t in self.tables) != self._ptr_key:
            raise RuntimeError("message tables were re-allocated after the optimizer was built")
        group = self.param_groups[-1]
        beta1, beta2 = group["betas"]
        return group, float(beta1), float(beta2)

    @torch.no_grad()
    def lookahead_sum(self, next_message):
        """S = sum_i table[2i + next_message_i] with the values the tables WILL hold after the pending step_tables() (for the
        message given to set_message), computed without touching the tables (nsig_msg_adam_lookahead_sum), and handed to the
        encoder as the pre-summed table of `next_message`.  Must be followed by exactly one step_tables() before anything
        else writes G; grad_scale / found_inf are read from the attributes a scaler's check() left on this optimizer."""
        if not self._train_tables:
            return None
        group, beta1, beta2 = self._table_call_args()
        self.sync_lr()
        md = self.enc.message_dim
        if getattr(self, "_S_ahead", None) is None:
            self._S_ahead = torch.empty_like(self.tables[0])
        nxt = next_message.to(device=self.G.device, dtype=torch.float32)
        _lib.call("nsig_msg_adam_lookahead_sum", _P(self._ptrs), len(self.tables), md, _P(self.message), _P(nxt), _P(self.G),
                  _P(self.steps), _P(self._coef), _P(getattr(self, "grad_scale", None)), _P(getattr(self, "found_inf", None)),
                  float(group["lr"]), beta1, beta2, float(group["eps"]), self.enc.log2_hashmap_size, _P(self._lr_dev),
                  self.shard[0] if self.shard else 0, self.shard[1] if self.shard else 0, _P(self._S_ahead))
        self._steps_prepared = True
        self.enc.presummed = (next_message, self._S_ahead)
        self._model._S_cache = None
        return self._S_ahead

    def sum_with_next_step(self, next_message):
        """Fold the NEXT forward's table sum into the next step()/step_tables(): the update kernel then also accumulates
        S = sum_i table[2i + next_message_i] over the updated tables (nsig_msg_adam_step_sum, bit-identical to update-then-
        sum) and hands it to the encoder as the pre-summed table of `next_message` - the separate pass over message_dim
        tables and its kernel disappear from the head of the next forward.  One use; None cancels."""
        self._sum_next = next_message

    @torch.no_grad()
    def step_tables(self):
        """The message tables' share of step(): one Adam kernel over the tables the message selects; joins the decoder's
        side stream (step_decoder) into the current stream."""
        nxt, self._sum_next = getattr(self, "_sum_next", None), None
        if self._train_tables and nxt is not None and not getattr(self, "_steps_prepared", False):
            group, beta1, beta2 = self._table_call_args()
            md = self.enc.message_dim
            if getattr(self, "_S_ahead", None) is None:
                self._S_ahead = torch.empty_like(self.tables[0])
            nx = nxt.to(device=self.G.device, dtype=torch.float32)
            _lib.call("nsig_msg_adam_step_sum", _P(self._ptrs), len(self.tables), md, _P(self.message), _P(nx), _P(self.G),
                      _P(self.steps), _P(self._coef), _P(getattr(self, "grad_scale", None)),
                      _P(getattr(self, "found_inf", None)), float(group["lr"]), beta1, beta2, float(group["eps"]),
                      self.enc.log2_hashmap_size, _P(self._lr_dev), self.shard[0] if self.shard else 0,
                      self.shard[1] if self.shard else 0, _P(self._S_ahead))
            self.enc.presummed = (nxt, self._S_ahead)
            self._model._S_cache = None
        elif self._train_tables:
            group, beta1, beta2 = self._table_call_args()
            md = self.enc.message_dim
            prepared = 1 if getattr(self, "_steps_prepared", False) else 0
            self._steps_prepared = False
            _lib.call("nsig_msg_adam_step", _P(self._ptrs), len(self.tables), md, _P(self.message), _P(self.G),
                      _P(self.steps), _P(self._coef), _P(getattr(self, "grad_scale", None)),
                      _P(getattr(self, "found_inf", None)), float(group["lr"]), beta1, beta2, float(group["eps"]),
                      self.enc.log2_hashmap_size, _P(self._lr_dev), self.shard[0] if self.shard else 0,
                      self.shard[1] if self.shard else 0, prepared)
            # the kernel writes the tables through raw pointers (no autograd version bump): drop the model's
            # cached S so the next forward re-sums the updated tables
            if not prepared:
                self._model._S_cache = None
        side = getattr(self, "_decoder_side", None)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            self._decoder_side = None


class _ScaledLoss:
    """What FusedGradScaler.scale(loss) returns: `.backward()` seeds autograd with the device-resident scale instead of
    multiplying the loss (no extra kernels; the loss-head backward kernel consumes the seed as d(loss))."""

    def __init__(self, loss, scale):
        self.loss, self.scale = loss, scale

    def backward(self, **kw):
        self.loss.backward(gradient=self.scale.reshape(self.loss.shape).to(self.loss.dtype), **kw)


class FusedGradScaler:
    """torch.amp.GradScaler semantics (the reference trains with --fp16: utils_wtmk_disen.py:1170-1181) for the flat-bucket
    optimizer, as ONE kernel per step: non-finite check of every gradient, found_inf, the scale the optimizer kernels
    divide by, and the scale update (backoff x0.5 on overflow, growth x2 after 2000 clean steps) - see
    nsig_grad_check_update_scale.  `scale(loss).backward(); step(optimizer); update()` as usual; update() is a no-op
    because the update already happened on the device.  state_dict() uses GradScaler's keys."""

    def __init__(self, device, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, int(growth_interval)
        self._scale = torch.tensor([float(init_scale)], dtype=torch.float32, device=device)
        self._growth_tracker = torch.zeros(1, dtype=torch.int32, device=device)
        self._found_inf = torch.zeros(1, dtype=torch.float32, device=device)
        self._step_scale = torch.tensor([float(init_scale)], dtype=torch.float32, device=device)
        self._scratch = torch.zeros(2, dtype=torch.int32, device=device)

    def is_enabled(self):
        return True

    def scale(self, loss):
        return _ScaledLoss(loss, self._scale)

    def check(self, optimizer, enabled=None):
        """The per-step kernel alone (non-finite check, found_inf, the scale this step's gradients carry, scale update) and
        the hand-over of `grad_scale` / `found_inf` to the optimizer as attributes; the caller then runs the optimizer's
        pieces itself (harness look-ahead schedule: step_decoder, lookahead_sum, step_tables) and calls release().
        enabled: optional device word (any 4-byte dtype); 0 = nothing pending, skip the step and leave the scaler alone."""
        if getattr(optimizer, "flat_bucket", None) is None:
            raise _lib.NsigError("FusedGradScaler needs an optimizer built over a flat gradient bucket "
                                 "(WatermarkAdam(flat_bucket=...)); use torch.amp.GradScaler otherwise")
        flat = optimizer.flat_bucket
        f = getattr(optimizer, "_flat", None)
        if f is not None:
            optimizer.scaler_bumps_step = True
        _lib.call("nsig_grad_check_update_scale", _P(flat), flat.numel(), _P(self._scale), _P(self._growth_tracker),
                  float(self.growth_factor), float(self.backoff_factor), self.growth_interval, _P(self._found_inf),
                  _P(self._step_scale), _P(f["step"]) if f is not None else None, _P(self._scratch), _P(enabled))
        # 0-dim views: torch's fused Adam (the non-flat decoder path) broadcasts found_inf against its 0-dim step tensors
        optimizer.grad_scale, optimizer.found_inf = self._step_scale.reshape(()), self._found_inf.reshape(())

    @staticmethod
    def release(optimizer):
        for a in ("grad_scale", "found_inf"):
            if hasattr(optimizer, a):
                delattr(optimizer, a)

    def step(self, optimizer, enabled=None):
        """check() + optimizer.step().  enabled: see check()."""
        self.check(optimizer, enabled)
        try:
            return optimizer.step()
        finally:
            self.release(optimizer)

    def update(self, new_scale=None):
        if new_scale is not None:
            self._scale.fill_(float(new_scale))

    def get_scale(self):
        return float(self._scale)

    def state_dict(self):
        return {"scale": self.get_scale(), "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": int(self._growth_tracker)}

    def load_state_dict(self, sd):
        self._scale.fill_(float(sd["scale"]))
        self.growth_factor, self.backoff_factor = float(sd["growth_factor"]), float(sd["backoff_factor"])
        self.growth_interval = int(sd["growth_interval"])
        self._growth_tracker.fill_(int(sd["_growth_tracker"]))
