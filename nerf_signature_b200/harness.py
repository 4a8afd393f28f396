"""Synthetic watermark-training harness: the caller side of the hot path.

The reference's Trainer and datasets are out of scope (host orchestration); this module
restates only what they feed the hot path with, on synthetic data (no datasets or checkpoints offline):
  * scene: random-init watermark NeRFNetwork + a deterministic occupancy grid (SURVEY 8d fixtures);
  * batch: `rays_o_block/rays_d_block [md, pH, pW, 3]` watermark blocks (provider_wtmk.py:481-496 shapes:
    pH = H // num_rows, pW = W // num_cols of a downscaled frame) and `rays_o/rays_d [1, num_rays, 3]`
    content rays with ground-truth colours (provider_wtmk.py:527-572);
  * step: utils_wtmk_disen.py:1164-1181 + 579-646 — fresh random message, two render passes
    (force_all_rays=True, perturb=False, bg_color=1), HiDDeN decoder, BCE(temp 10) + MSE, backward, Adam.
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import synthetic as syn
from . import parallel
from . import hash_encoding_wtmk_bit as _hmsg
from . import _lib

CONFIGS = {
    # BASELINE.json configs[1]: Blender shape, bound 1.0, scale 0.8, dt_gamma 0, message_dim 32, 32x32 codebook
    "blender_wtmk": dict(bound=1.0, scale=0.8, dt_gamma=0.0, message_dim=32, num_rows=32, num_cols=32,
                         H=400, W=400, num_rays=4096, camera="blender", occupancy="sphere"),
    # configs[2]: Mip-NeRF-360 shape, bound 2 (CLI default), scale 0.33, dt_gamma 0
    "360_wtmk": dict(bound=2.0, scale=0.33, dt_gamma=0.0, message_dim=32, num_rows=32, num_cols=32,
                     H=756, W=1008, num_rays=4096, camera="360", occupancy="sphere", grid_update_every=16),
    # configs[4] per-rank shape: message_dim 48, rays sharded across ranks
    "shard262144_wtmk": dict(bound=1.0, scale=0.8, dt_gamma=0.0, message_dim=48, num_rows=32, num_cols=32,
                             H=400, W=400, num_rays=262144, camera="blender", occupancy="sphere"),
}


def _pose(cfg, rs):
    radius = 4.0311 * cfg["scale"] if cfg["camera"] == "blender" else 4.0 * cfg["scale"]
    return syn.orbit_pose(rs.uniform(math.pi / 3, 2 * math.pi / 3), rs.uniform(0, 2 * math.pi), radius)


def _focal(cfg):
    fov = 0.6911112 if cfg["camera"] == "blender" else 0.9
    return 0.5 * cfg["W"] / math.tan(0.5 * fov)


def make_batch(cfg, seed, num_rays=None, n_blocks=None):
    """Host (numpy) batch: block rays of `n_blocks` (default message_dim) randomly chosen blocks and
    `num_rays` random content pixels of one synthetic view, plus random ground-truth colours."""
    rs = np.random.RandomState(seed)
    H, W = cfg["H"], cfg["W"]
    md = cfg["message_dim"] if n_blocks is None else n_blocks
    num_rays = cfg["num_rays"] if num_rays is None else num_rays
    pH, pW = H // cfg["num_rows"], W // cfg["num_cols"]
    focal = _focal(cfg)
    pose = _pose(cfg, rs)
    blocks = rs.permutation(cfg["num_rows"] * cfg["num_cols"])[:md]  # provider_wtmk.py:199-204 (seeded randperm)
    ii, jj = np.meshgrid(np.arange(pH), np.arange(pW), indexing="ij")
    ids = []
    for b in blocks:
        r, c = divmod(int(b), cfg["num_cols"])
        ids.append(((r * pH + ii) * W + (c * pW + jj)).reshape(-1))
    ids = np.concatenate(ids)
    bo, bd = syn.camera_rays(pose, H, W, focal, ids)
    pose2 = _pose(cfg, rs)
    co, cd = syn.camera_rays(pose2, H, W, focal, rs.randint(0, H * W, size=num_rays))
    return {
        "rays_o_block": bo.reshape(md, pH, pW, 3), "rays_d_block": bd.reshape(md, pH, pW, 3),
        "rays_o": co.reshape(1, num_rays, 3), "rays_d": cd.reshape(1, num_rays, 3),
        "gt": rs.uniform(size=(1, num_rays, 3)).astype(np.float32),
    }


def shard_batch(batch_np, rank, world_size):
    """SURVEY 8(e) partitioning of ONE global batch: rank r takes the contiguous ray range [r*N/k, (r+1)*N/k) of the
    content rays (with their ground truth) and of the flattened watermark-block rays.  Returns (local batch, block_shape,
    per-rank block-ray counts); the block rays come back flat ([n_local, 3]) - the rendered pixels are all-gathered into
    the full [md, pH, pW, 3] image before the decoder (parallel.all_gather_pixels)."""
    bo, bd = batch_np["rays_o_block"], batch_np["rays_d_block"]
    block_shape = tuple(bo.shape[:-1])
    nb = int(np.prod(block_shape))
    n = batch_np["rays_o"].shape[1]
    blo, bhi = parallel.shard_range(nb, rank, world_size)
    clo, chi = parallel.shard_range(n, rank, world_size)
    counts = [parallel.shard_range(nb, r, world_size)[1] - parallel.shard_range(nb, r, world_size)[0]
              for r in range(world_size)]
    local = {"rays_o_block": np.ascontiguousarray(bo.reshape(-1, 3)[blo:bhi]),
             "rays_d_block": np.ascontiguousarray(bd.reshape(-1, 3)[blo:bhi]),
             "rays_o": np.ascontiguousarray(batch_np["rays_o"][:, clo:chi]),
             "rays_d": np.ascontiguousarray(batch_np["rays_d"][:, clo:chi]),
             "gt": np.ascontiguousarray(batch_np["gt"][:, clo:chi])}
    return local, block_shape, counts


def occupancy(cfg, cascade, seed=0):
    if cfg["occupancy"] == "sphere":
        return syn.sphere_grid(cascade)
    return syn.bernoulli_grid(cascade, p=0.5, seed=seed)


class _join_on_backward(torch.autograd.Function):
    """Identity whose backward makes the current stream wait for `stream` first: placed on the rendered block pixels, the
    wait lands between the decoder's backward and the composite / field backward."""

    @staticmethod
    def forward(ctx, x, stream):
        ctx.stream = stream
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        torch.cuda.current_stream().wait_stream(ctx.stream)
        return g, None


class Scene:
    """Model + optimizer + GradScaler as main_nerf_wtmk.py:92-117 sets them up."""

    def __init__(self, cfg, device, seed=0, lr=1e-2, fp16=True, table_scale=1.0, optimizer="fused", graph=False,
                 merged_render=False, fused_decoder=False, fused_losses=False, overlap_decoder=False,
                 shard_blocks=None, distributed=True, fused_scaler=None, defer_optimizer=False, shard_optimizer=None,
                 lookahead=None, march_ahead=False, distortion="none"):
        """optimizer: "fused" = optim.WatermarkAdam (one kernel for the message tables, capture-safe);
        "torch" = torch.optim.Adam over get_params, exactly as main_nerf_wtmk.py:107 builds it.
        graph: capture the whole step (both render passes, decoder, losses, backward, optimizer, scaler)
        in one CUDA graph on first use; needs optimizer="fused".
        fused_losses: clamp / MSE / BCE / weighting through the loss-head kernels (nerf/loss_ops.py) instead of the
        plain torch expressions of utils_wtmk_disen.py:593,636-644.
        overlap_decoder (two render calls, like the reference step): the decoder runs on a side stream.  Its forward
        and backward are chains of ~10 small latency-bound kernels each that leave most SMs idle; with the watermark
        blocks rendered first, the content pass (march, field forward, composite - and in the backward pass its field
        backward) executes next to them.  In the captured step these are parallel branches of one graph.
        shard_blocks: (block_shape, per-rank block-ray counts) from shard_batch - the batch holds this rank's slice of
        ONE global batch: block pixels are all-gathered before the decoder so that every rank decodes the
        full batch (global BatchNorm statistics) and N GPUs compute what one GPU would.
        distributed=False: ignore the process group (single-GPU semantics inside a multi-rank job; gradient checks).
        defer_optimizer: software-pipeline the optimizer: the Adam update of step t (HBM-bound, ~0.14 ms) is issued at the START
        of step t+1 on the side stream that also builds the summed message table, next to the latency-bound march of step t+1
        (which needs neither the tables nor G).  Every kernel sees exactly the data it would see in the sequential order -
        Adam(t) still completes before the table sum of step t+1 - so results are unchanged; reads of the message tables
        from outside (evaluation, checkpoints, update_extra_state) must call flush_optimizer() first (train_step does).
        march_ahead (captured, merged step only): software-pipeline the sample generation.  near/far + march of batch t+1 read
        only its rays and the occupancy bitfield, so they are issued inside step t on a branch next to the decoder's
        chain of small latency-bound kernels (which leaves most of every SM idle), instead of at the head of step t+1 where
        they compete with the HBM-saturating table Adam and finish 150 us late.  train_step(batch, message, next_batch=...,
        next_message=...) announces the following batch; an unannounced batch is staged and marched on the spot (correct,
        not overlapped).  Two input buffers / sample-buffer sets / graphs alternate.  Kernels and data are those of the
        plain schedule; only when the march runs changes.  Not for configs that rebuild the bitfield (grid_update_every)."""
        from .nerf.network_wtmk_tcnn import NeRFNetwork
        from .optim import WatermarkAdam
        torch.manual_seed(seed)
        self.cfg = cfg
        self.device = device
        model = NeRFNetwork(bound=cfg["bound"], cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10,
                            bg_radius=-1, message_dim=cfg["message_dim"], n_views=1)
        if table_scale != 1.0:
            with torch.no_grad():
                for e in list(model.encoder.embeddings) + list(model.msg_encoder.embeddings):
                    e.weight.mul_(table_scale)
        grid = occupancy(cfg, model.cascade, seed)
        model.density_grid.copy_(torch.from_numpy(grid))
        model.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, 0.5)))
        if not (fused_decoder and fp16):
            # plain-module decoder: NHWC weights.  Its input `pred.permute(0, 3, 1, 2)` of an [md, pH, pW, 3] image is
            # already a channels_last view and cuDNN's tensor-core kernels are NHWC; with NCHW weights every conv is
            # wrapped in layout-conversion kernels (-4 % step time).  Shapes and state-dict keys are unaffected.
            model.msg_decoder = model.msg_decoder.to(memory_format=torch.channels_last)
        self.model = model.to(device).train()
        if graph and optimizer != "fused":
            raise ValueError("graph capture needs the fused optimizer (the set of tables with a gradient changes "
                             "with every message under torch.optim.Adam)")
        self.fused = optimizer == "fused"
        self.sync = parallel.GradSync()
        if not distributed:
            self.sync.enabled, self.sync.exchange = False, "none"
        self.shard_blocks = shard_blocks
        self._decoder_params = [p for p in self.model.msg_decoder.parameters()]
        # one flat [dL/dS | decoder gradients] buffer whenever the fused optimizer is used: a single fill per step
        # instead of one zeros/add pair per decoder parameter, and the bucket of the multi-GPU exchange
        self.flat_sync = self.fused
        if self.fused:
            gbuf = None
            if self.flat_sync:  # one flat bucket [dL/dS | decoder grads] -> one all-reduce per step
                gbuf = self.sync.make_flat_buffer(self.model.msg_encoder.tables()[0].numel(), self._decoder_params, device)
            # multi-GPU: the message tables' optimizer is sharded over the ranks (every rank sees the same exchanged G)
            do_shard = (self.sync.enabled and os.environ.get("NSIG_NO_SHARD_OPT") != "1") if shard_optimizer is None \
                else (shard_optimizer and self.sync.enabled)
            self.optimizer = WatermarkAdam(self.model, lr=lr, betas=(0.9, 0.99), eps=1e-15, capturable=graph,
                                           grad_buffer=gbuf, flat_bucket=self.sync.flat if self.flat_sync else None,
                                           shard=parallel.world() if do_shard else None)
        else:
            self.optimizer = torch.optim.Adam(self.model.get_params(lr), betas=(0.9, 0.99), eps=1e-15, fused=True)
        self.fp16 = fp16
        # fused_scaler (default with the fused optimizer): GradScaler's check / found_inf / scale update as one kernel over
        # the flat bucket (optim.FusedGradScaler); False keeps torch.amp.GradScaler (what the reference uses)
        use_fs = (self.fused and fp16 and os.environ.get("NSIG_TORCH_SCALER") != "1") if fused_scaler is None \
            else (fused_scaler and self.fused and fp16)
        if use_fs:
            from .optim import FusedGradScaler
            self.scaler = FusedGradScaler(device)
        else:
            self.scaler = torch.amp.GradScaler("cuda", enabled=fp16)
        self.lambda_w, self.lambda_i = 0.005, 1.0  # README.md:40,45
        self.opt = dict(dt_gamma=cfg["dt_gamma"], max_steps=1024, T_thresh=1e-4)
        _hmsg.grad_reducer = self.sync.reduce_table_grad if (self.sync.enabled and not self.flat_sync) else None
        self.use_graph = graph
        self.merged_render = merged_render  # one render call over [block rays | content rays] instead of two
        self.fused_decoder = fused_decoder and fp16  # the kernels implement the float16-autocast arithmetic
        self.fused_losses = fused_losses
        if self.fused_decoder and self.fused and os.environ.get("NSIG_NO_DEC_PREP") != "1":
            # fp16 weight copies of the decoder: converted behind every optimizer update (on its side stream) instead of at the
            # head of every decoder forward, i.e. off the step's critical path
            self.model.prepare_decoder_weights()
        # the decoder backward hands back the block-pixel gradient first and finishes its weight gradients next to the composite /
        # field backward (decoder_ops.finish_backward after loss.backward()); NSIG_NO_DEC_DEFER=1 joins them inside the call
        self.model.defer_decoder_weight_grads = self.fused_decoder and os.environ.get("NSIG_NO_DEC_DEFER") != "1"
        self.overlap_decoder = overlap_decoder and not merged_render
        self.defer_optimizer = bool(defer_optimizer) and self.fused and use_fs   # needs the one-kernel scaler's skip flag
        # look-ahead schedule of the deferred optimizer (see _step_impl).  Opt-in (lookahead=True or NSIG_LOOKAHEAD=1): measured
        # SLOWER than the plain deferred order (1.02 vs 0.97 ms) - the field forward does start 80 us earlier, but the table
        # Adam saturates HBM and every latency-bound kernel that runs beside it crawls (the decoder forward took 238-285 us
        # instead of 113, finishing ~90 us after the Adam whatever its grid) - profiles/r02_experiments.txt
        if lookahead is None:
            lookahead = os.environ.get("NSIG_LOOKAHEAD") == "1"
        self.lookahead = self.defer_optimizer and bool(lookahead) and merged_render
        # deferred order: the pending update's kernel also builds S for the step that issues it (nsig_msg_adam_step_sum)
        self.fuse_table_sum = self.defer_optimizer and not self.lookahead and os.environ.get("NSIG_NO_FUSED_SUM") != "1"
        if self.defer_optimizer:   # [message of the step whose update is pending (md) | pending flag (1)]
            self._opt_state = torch.zeros(cfg["message_dim"] + 1, dtype=torch.float32, device=device)
        self.march_ahead = bool(march_ahead)
        if self.march_ahead and not (graph and merged_render and shard_blocks is None
                                     and not cfg.get("grid_update_every", 0)):
            raise ValueError("march_ahead needs graph=True, merged_render=True, an unsharded batch and a config without "
                             "grid_update_every (marched-ahead samples go stale when the occupancy bitfield changes)")
        # grid limit of the marched-ahead kernels (CTAs of 8 warps): they must leave every SM room for a decoder CTA
        self.march_ahead_blocks = int(os.environ.get("NSIG_MARCH_AHEAD_CTAS_PER_SM", "2")) * 148
        self._ahead_p = 0           # parity: which input buffer / sample-buffer set holds the batch to train on next
        self._ahead_token = None    # id() of the batch announced for the next call
        self._ahead_last = 0
        self.iteration = 0
        self.keep_outputs = False   # parity tests: keep the step's rendered pixels and decoder logits in self.last
        # training-time attack on the rendered blocks before the decoder (Trainer.distortion_layer, CLI --distortion);
        # 'none' (default) adds nothing to the step
        from .nerf.distortion import CAPTURABLE, KINDS
        if distortion not in KINDS:
            raise ValueError(f"distortion must be one of {KINDS}, got {distortion!r}")
        if graph and distortion not in CAPTURABLE:
            raise ValueError(f"distortion {distortion!r} runs in the eager step only (four of the attacks draw a parameter on the "
                             "host, as the reference does; 'noise' was never measured inside the captured step): use graph=False")
        self.distortion = distortion
        self.last = None
        self._graph = None
        self._static = None
        self.launches_per_step = None

    def new_message(self, generator=None):
        """utils_wtmk_disen.py:1165 — a fresh random message every step.  Generated on the host so the
        bit pattern is known without a device round trip; the device copy is asynchronous."""
        return torch.randint(0, 2, (self.cfg["message_dim"],), generator=generator).float()

    def to_device(self, batch_np, pinned=None):
        """numpy batch -> device tensors (through pinned host buffers, asynchronously)."""
        out = {}
        for k, v in batch_np.items():
            t = torch.from_numpy(v)
            if pinned is not None:
                pinned[k].copy_(t)
                t = pinned[k]
            out[k] = t.to(self.device, non_blocking=True)
        return out

    def _step_impl(self, batch, message, ahead=None):
        """utils_wtmk_disen.py:1164-1181 + 579-646; `message` is a host tensor (torch-Adam path: the bits
        select which tables receive a gradient) or a device tensor (fused path: nothing on the host depends
        on the bits)."""
        model = self.model
        if self.defer_optimizer:
            pass                      # cleared after the deferred optimizer step below
        elif self.flat_sync:
            self.sync.zero_flat()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        msg_dev = message.to(self.device, non_blocking=True) if not message.is_cuda else message
        if self.defer_optimizer:
            # the PREVIOUS step's optimizer update, on the stream the summed-table prefetch uses (so the table sum is ordered
            # after it) while this step's near/far + march run on the main stream; then the bucket is cleared for this
            # step's backward.  The render call joins the side stream before the first field kernel.
            md = self.cfg["message_dim"]
            side = _lib.side_stream(self.device, 0)
            main = torch.cuda.current_stream()
            if side is not None:
                side.wait_stream(main)
            with torch.cuda.stream(side if side is not None else main):
                self.optimizer.set_message(self._opt_state[:md])
                if self.lookahead:
                    # look-ahead schedule: only what the field forward needs happens here - the scaler's check, the
                    # decoder's Adam (+ its fp16 weights) and S as the tables WILL be after the pending update, read-only
                    # (~0.26 GB instead of the update's ~0.8 GB); the table update itself is issued after the composite
                    # forward, next to the latency-bound decoder chain, and joined before the field backward writes G
                    self.scaler.check(self.optimizer, enabled=self._opt_state[md:])
                    self.optimizer.step_decoder()
                    ds = self.optimizer._decoder_side
                    if ds is not None:
                        torch.cuda.current_stream().wait_stream(ds)
                    self.sync.zero_decoder_grads()
                    self.optimizer.lookahead_sum(msg_dev)
                else:
                    if self.fuse_table_sum:   # S of THIS step's message accumulated by the pending update's own pass
                        self.optimizer.sum_with_next_step(msg_dev)
                    self.scaler.step(self.optimizer, enabled=self._opt_state[md:])
                    self.sync.zero_flat()
            if side is not None:
                self._opt_side = side
            message = msg_dev
        elif self.fused:
            self.optimizer.set_message(msg_dev)
            message = msg_dev
        from .nerf.loss_ops import split_clamp, wtmk_loss
        ob = batch["rays_o_block"]
        nb = ob.numel() // 3
        block_shape = tuple(ob.shape)
        gather = None
        if self.shard_blocks is not None:   # this rank's slice of the block rays; pixels are gathered before the decoder
            block_shape = tuple(self.shard_blocks[0]) + (3,)
            if self.sync.enabled:
                ws = parallel.world()[1]
                gather = lambda px: parallel.all_gather_pixels(px.reshape(-1, 3), self.shard_blocks[1], grad_scale=ws)
        if self.merged_render:
            # the two render passes of the reference step (utils_wtmk_disen.py:592,641) see the same network and the
            # same message and rays are independent: one launch chain over [block rays | content rays]
            if "rays_o_all" in batch:   # static buffers of the captured step: already concatenated
                o_all, d_all = batch["rays_o_all"], batch["rays_d_all"]
            else:
                o_all = torch.cat([ob.reshape(1, nb, 3), batch["rays_o"]], dim=1)
                d_all = torch.cat([batch["rays_d_block"].reshape(1, nb, 3), batch["rays_d"]], dim=1)
            out = model.render(o_all, d_all, message, staged=False, bg_color=1, perturb=False, force_all_rays=True,
                               premarched=None if ahead is None else ahead[0], **self.opt)
            if ahead is not None:
                # samples of the NEXT batch (ahead = (this batch's sample buffers, next input buffer, its sample buffers)):
                # a branch that starts behind the composite forward and runs beside the decoder; joined at the end of the step
                mside = _lib.side_stream(self.device, 4)
                main = torch.cuda.current_stream()
                if mside is not None:
                    mside.wait_stream(main)
                with torch.cuda.stream(mside if mside is not None else main):
                    model.march_ahead(ahead[1]["rays_o_all"], ahead[1]["rays_d_all"], ahead[2],
                                      dt_gamma=self.opt["dt_gamma"], max_steps=self.opt["max_steps"],
                                      max_blocks=self.march_ahead_blocks)
            if self.fused_losses:
                pred, image_c = split_clamp(out["image"], nb)
                image_c = image_c.view(batch["rays_o"].shape)
            else:
                image_w, image_c = out["image"][0, :nb], out["image"][:, nb:]
        else:
            image_w = model.render(batch["rays_o_block"].reshape(1, nb, 3), batch["rays_d_block"].reshape(1, nb, 3), message,
                                   staged=False, bg_color=1, perturb=False, force_all_rays=True, **self.opt)["image"]
            if self.fused_losses:
                pred = split_clamp(image_w, nb)[0]
        if not self.fused_losses:
            pred = torch.clamp(image_w, min=0, max=1)
        if self.defer_optimizer and self.lookahead:
            # the pending table update (reads G and the tables, writes the tables) + the clear of G: a branch that starts
            # here, after the field forward (which holds every SM) and ends before the field backward writes G
            tail = _lib.side_stream(self.device, 3)
            main = torch.cuda.current_stream()
            if tail is not None:
                tail.wait_stream(main)
            with torch.cuda.stream(tail if tail is not None else main):
                self.optimizer.step_tables()
                self.sync.zero_table_grad()
                self.scaler.release(self.optimizer)
            if tail is not None:
                pred = _join_on_backward.apply(pred, tail)
        if gather is not None:
            pred = gather(pred)
        pred = pred.reshape(block_shape)
        if self.distortion != "none":   # utils_wtmk_disen.py:594: decoder sees the attacked blocks, the loss the clean ones
            from .nerf.distortion import distortion_layer
            pred = distortion_layer(pred, self.distortion)
        side = _lib.side_stream(self.device, 2) if self.overlap_decoder else None
        main = torch.cuda.current_stream()
        if side is not None:     # fork: the decoder chain runs next to the content pass below
            side.wait_stream(main)
            pred.record_stream(side)
        with torch.cuda.stream(side if side is not None else main):
            if self.fused_decoder:   # normalisation + HiDDeN decoder forward/backward as tensor-core kernels (csrc/decoder.cu)
                decoded = model.decode_blocks(pred)
            else:                    # utils_wtmk_disen.py:592-595 verbatim: the plain module under autocast
                with torch.autocast("cuda", dtype=torch.float16, enabled=self.fp16):
                    decoded = model.msg_decoder(model.normalization(pred.permute(0, 3, 1, 2)))
        if not self.merged_render:
            image_c = model.render(batch["rays_o"], batch["rays_d"], message, staged=False, bg_color=1,
                                   perturb=False, force_all_rays=True, **self.opt)["image"]
        if side is not None:     # join before the loss reads the logits
            main.wait_stream(side)
            decoded.record_stream(main)
        if self.fused_losses:
            loss, lossi, lossw = wtmk_loss(image_c, batch["gt"], decoded, msg_dev, self.lambda_w, self.lambda_i, 10.0)
        else:
            lossi = F.mse_loss(image_c, batch["gt"], reduction="none").mean()
            lossw = F.binary_cross_entropy_with_logits(decoded.float() * 10.0, msg_dev.unsqueeze(-1), reduction="mean")
            loss = self.lambda_w * lossw + self.lambda_i * lossi
        if self.keep_outputs:
            self.last = {"pred": pred.detach(), "image_c": image_c.detach(), "decoded": decoded.detach()}
        self.scaler.scale(loss).backward()
        if self.fused_decoder:   # deferred tail of the decoder backward (its weight gradients ran next to the renderer's backward)
            from .nerf.decoder_ops import finish_backward
            finish_backward()
        if self.flat_sync:
            self.sync.reduce_flat()
        else:
            self.sync.reduce_params(self._decoder_params)
        if self.defer_optimizer:   # remember whose update is pending; it runs at the start of the next step (or in flush)
            md = self.cfg["message_dim"]
            self._opt_state[:md].copy_(msg_dev)
            self._opt_state[md:].fill_(1.0)
        else:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        if ahead is not None:
            mside = _lib.side_stream(self.device, 4)
            if mside is not None:
                torch.cuda.current_stream().wait_stream(mside)
        return loss, lossi, lossw

    def flush_optimizer(self):
        """Apply the pending (deferred) optimizer update now; no-op otherwise.  After it the message tables and the decoder
        hold what a sequential step would have left, and the next step's leading update is skipped on the device."""
        if not self.defer_optimizer:
            return
        md = self.cfg["message_dim"]
        with torch.no_grad():
            self.optimizer.set_message(self._opt_state[:md])
            self.scaler.step(self.optimizer, enabled=self._opt_state[md:])
            self._opt_state[md:].zero_()
        self.model._S_cache = None
        self.model.msg_encoder.presummed = None

    def _capture(self, batch, message, next_batch=None, next_message=None):
        """Warm up on a side stream, then record one step into a CUDA graph with static input buffers (march_ahead: two
        buffers and two graphs - graph p trains on buffer p, whose samples exist, and marches buffer 1-p)."""
        from . import _lib
        # ONE flat static input buffer; the per-key entries are views of it.  Layout (floats):
        #   [rays_o_block | rays_o | rays_d_block | rays_d | gt | message]
        # so [block rays | content rays] are already concatenated for the merged render (no torch.cat in the captured
        # step) and a host batch packed the same way (`pinned_batch`) needs a single H2D copy per step.
        md = self.cfg["message_dim"]
        order = ["rays_o_block", "rays_o", "rays_d_block", "rays_d", "gt"]
        batch = {k: batch[k] for k in order}
        self._layout, o = {}, 0
        for k in order:
            self._layout[k] = (o, tuple(batch[k].shape))
            o += batch[k].numel()
        self._layout["message"] = (o, (md,))
        nb, nc = batch["rays_o_block"].numel() // 3, batch["rays_o"].numel() // 3

        def make_static():
            flat = torch.empty(o + md, dtype=torch.float32, device=self.device)
            views = {k: flat[a:a + math.prod(shp)].view(shp) for k, (a, shp) in self._layout.items()}
            if self.merged_render:
                for a in ("o", "d"):
                    lo = self._layout[f"rays_{a}_block"][0]
                    views[f"rays_{a}_all"] = flat[lo:lo + 3 * (nb + nc)].view(1, nb + nc, 3)
            return flat, views

        self._static_flat, self._static = make_static()
        self._copy_inputs(batch, message)
        side = torch.cuda.Stream()
        if self.march_ahead:
            self._statics = [(self._static_flat, self._static), make_static()]
            self._pm = [self.model.march_buffers(nb + nc, self.opt["max_steps"], self.device) for _ in range(2)]
            if next_batch is None:
                next_batch, next_message = batch, message
            self._copy_inputs(next_batch, next_message, which=1)
            self._march_now(0)

            def step(p):
                st, other = self._statics[p][1], self._statics[1 - p][1]
                return self._step_impl(st, st["message"], ahead=(self._pm[p], other, self._pm[1 - p]))

            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(4):      # an even count: buffer 0 is the marched one again afterwards
                    step(i % 2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graphs, self._static_outs = [], []
            _lib.timing_reset()
            for p in range(2):
                g = torch.cuda.CUDAGraph()
                _lib.timing_tag = p
                n0 = _lib.launch_count
                with torch.cuda.graph(g):
                    self._static_outs.append(step(p))
                self.launches_per_step = _lib.launch_count - n0
                self._graphs.append(g)
            _lib.timing_tag = None
            self._graph = self._graphs[0]
            self._ahead_p = 0
            return
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step_impl(self._static, self._static["message"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        _lib.timing_reset()  # keep only the (external) events recorded inside the graph
        n0 = _lib.launch_count
        with torch.cuda.graph(self._graph):
            self._static_out = self._step_impl(self._static, self._static["message"])
        self.launches_per_step = _lib.launch_count - n0
        ls = self.model.local_step
        n_calls = 1 if self.merged_render else 2
        self._graph_rows = [(ls - k) % 16 for k in range(n_calls, 0, -1)]  # the march counters baked into the graph

    def _march_now(self, p):
        """Samples of input buffer p, marched on the spot (priming, or a batch that was not announced)."""
        st = self._statics[p][1]
        self.model.march_ahead(st["rays_o_all"], st["rays_d_all"], self._pm[p], dt_gamma=self.opt["dt_gamma"],
                               max_steps=self.opt["max_steps"])

    def _copy_inputs(self, batch, message, which=None):
        sflat, static = (self._static_flat, self._static) if which is None else self._statics[which]
        flat = batch.get("_flat") if isinstance(batch, dict) else None
        if flat is not None and flat.numel() == sflat.numel():
            sflat.copy_(flat, non_blocking=True)   # packed host batch (pinned_batch): one H2D copy
            return
        for k, v in batch.items():
            if k not in ("_flat", "message"):
                static[k].copy_(v, non_blocking=True)
        static["message"].copy_(message, non_blocking=True)

    def pinned_batch(self, batch_np):
        """A host batch in pinned memory, packed in the layout of the captured step's static input buffer (what a data
        loader with pin_memory would hand over): dict of views + "_flat" (the whole buffer) + "message" (view to fill
        with the step's message).  train_step(pb, pb["message"]) then moves it to the device with one copy."""
        if self._graph is None:
            raise RuntimeError("pinned_batch needs the captured step (graph=True, after the first train_step)")
        flat = torch.empty(self._static_flat.numel(), dtype=torch.float32).pin_memory()
        out = {"_flat": flat}
        for k, (a, shp) in self._layout.items():
            out[k] = flat[a:a + math.prod(shp)].view(shp)
            if k != "message":
                out[k].copy_(torch.from_numpy(batch_np[k]))
        return out

    def train_step(self, batch, message, next_batch=None, next_message=None):
        """One optimisation step; returns (loss, lossi, lossw) as device scalars (no host sync here).  Configs with
        `grid_update_every` also run NeRFRenderer.update_extra_state every that many iterations (the clean trainer's
        cadence, nerf/utils.py:855-857; BASELINE configs[2] asks for it in watermark training too).
        march_ahead: `next_batch` / `next_message` announce the batch of the following call (its input copy and its
        march happen during this step); the message of an announced batch is the one it was announced with."""
        if self.march_ahead:
            out = self._train_step_ahead(batch, message, next_batch, next_message)
            self.iteration += 1
            return out
        out = self._train_step(batch, message)
        self.iteration += 1
        every = self.cfg.get("grid_update_every", 0)
        if every and self.iteration % every == 0:
            self.flush_optimizer()   # the occupancy sweep reads the message tables: they must be up to date
            self.model.update_extra_state(message.to(self.device) if not message.is_cuda else message)
        return out

    def _train_step(self, batch, message):
        if not self.use_graph:
            batch = {k: (v if v.is_cuda else v.to(self.device, non_blocking=True)) for k, v in batch.items()
                     if k not in ("_flat", "message")}
            return self._step_impl(batch, message)
        if self._graph is None:
            self._capture(batch, message)
        self._copy_inputs(batch, message)
        if self.fused:
            self.optimizer.sync_lr()        # lr schedulers act on host floats; the captured kernels read device scalars
        self._graph.replay()
        # the replayed Adam kernel rewrote the message tables through raw pointers: neither optimizer.step()'s cache
        # invalidation nor an autograd version bump happened, so drop the cached summed table S here
        self.model._S_cache = None
        return self._static_out

    def _train_step_ahead(self, batch, message, next_batch, next_message):
        if self._graph is None:
            self._capture(batch, message, next_batch, next_message)
            self._ahead_token = None
        p = self._ahead_p
        if self._ahead_token is None or self._ahead_token != id(batch):
            self._copy_inputs(batch, message, which=p)     # not announced: stage and march it now
            self._march_now(p)
        if next_batch is not None:
            self._copy_inputs(next_batch, message if next_message is None else next_message, which=1 - p)
            self._ahead_token = id(next_batch)
        else:   # nothing announced: the graph re-marches whatever buffer 1-p holds; the next call stages its own batch
            self._ahead_token = None
        self.optimizer.sync_lr()
        self._graphs[p].replay()
        self.model._S_cache = None
        self._ahead_last = p
        self._ahead_p = 1 - p
        return self._static_outs[p]

    def _counter_rows(self):
        if self.use_graph and getattr(self, "_graph_rows", None) is not None:
            return self._graph_rows
        ls = self.model.local_step
        return [(ls - k) % 16 for k in range(1 if self.merged_render else 2, 0, -1)]

    def samples_per_step(self):
        """Measured (samples, rays) of the two most recent render calls = one training step (reads the march
        counters the kernels left on the device)."""
        if self.march_ahead and self._graph is not None:
            r = self._pm[self._ahead_last]["counter"].tolist()
            return r[0], r[1]
        rows = self.model.step_counter[self._counter_rows()].tolist()
        return sum(r[0] for r in rows), sum(r[1] for r in rows)

    def samples_per_ray(self):
        s, r = self.samples_per_step()
        return s / max(r, 1)


# -------------------------------------------------------------------------------------------------------
# full-frame inference (BASELINE configs[3]): 800x800 Blender / 1008x756 LLFF-shaped test views
# -------------------------------------------------------------------------------------------------------
FRAME_CONFIGS = {
    "blender_800x800": dict(H=800, W=800, bound=1.0, fov=0.6911112, radius=4.0311 * 0.8),
    "llff_1008x756": dict(H=756, W=1008, bound=2.0, fov=0.9, radius=4.0 * 0.33),
}


def frame_rays(fc, view):
    """All H*W rays of synthetic test view `view` (host numpy)."""
    rs = np.random.RandomState(9000 + view)
    focal = 0.5 * fc["W"] / math.tan(0.5 * fc["fov"])
    pose = syn.orbit_pose(rs.uniform(math.pi / 3, 2 * math.pi / 3), rs.uniform(0, 2 * math.pi), fc["radius"])
    return syn.camera_rays(pose, fc["H"], fc["W"], focal, np.arange(fc["H"] * fc["W"]))


def frame_block_ids(fc, message_dim, num_rows=32, num_cols=32, seed=0):
    """Flat pixel ids of `message_dim` randomly chosen (seeded randperm, provider_wtmk.py:199-204) blocks of
    pH x pW = (H // num_rows) x (W // num_cols) pixels of a full frame: [md * pH * pW] int64 + the block shape."""
    H, W = fc["H"], fc["W"]
    pH, pW = H // num_rows, W // num_cols
    rs = np.random.RandomState(seed)
    blocks = rs.permutation(num_rows * num_cols)[:message_dim]
    ii, jj = np.meshgrid(np.arange(pH), np.arange(pW), indexing="ij")
    ids = np.concatenate([(((b // num_cols) * pH + ii) * W + ((b % num_cols) * pW + jj)).reshape(-1) for b in blocks])
    return ids.astype(np.int64), (message_dim, pH, pW)


def time_frames(name, device, views, message_dim=32, seed=0, fused=True, staged=False, extract_bits=True):
    """Render `views` (list of view ids) of frame config `name` through NeRFRenderer.render in eval mode
    (main_nerf_wtmk.py --test path: renderer_wtmk.py:541-574 -> run_cuda inference branch) with a random-init
    watermark network and the sphere occupancy fixture, and - BASELINE configs[3] "+ HiDDeN bit extraction" - decode the
    message from message_dim blocks of every rendered frame (Trainer.test_bitacc -> eval_step, utils_wtmk_disen.py:935,
    663-669: clamp, normalise, msg_decoder; bit = logit > 0) inside the timed region.
    Returns a dict: ms per frame, samples per frame, bit accuracy (random-init weights: ~0.5), launches per frame."""
    from .nerf.network_wtmk_tcnn import NeRFNetwork
    fc = FRAME_CONFIGS[name]
    torch.manual_seed(seed)
    net = NeRFNetwork(bound=fc["bound"], cuda_ray=True, message_dim=message_dim).to(device).eval()
    grid = syn.sphere_grid(net.cascade)
    net.density_grid.copy_(torch.from_numpy(grid))
    net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, 0.5)))
    net.fused_inference = fused
    msg = torch.randint(0, 2, (message_dim,)).float().to(device)
    ids, bshape = frame_block_ids(fc, message_dim)
    ids = torch.from_numpy(ids).to(device)
    frames = []
    for v in views:
        o, d = frame_rays(fc, v)
        frames.append((torch.from_numpy(o)[None].to(device), torch.from_numpy(d)[None].to(device)))
    kw = dict(staged=staged, bg_color=1, perturb=False, dt_gamma=0.0, max_steps=1024)
    counts = []
    hits = torch.zeros((), dtype=torch.float32, device=device)

    def one(f):
        out = net.render(*f, msg, **kw)
        if extract_bits:
            blocks = torch.clamp(out["image"].view(-1, 3)[ids].view(*bshape, 3), min=0, max=1)
            logits = net.decode_blocks(blocks)
            hits.add_(((logits.float().view(-1) > 0) == (msg > 0.5)).float().mean())
        return out

    with torch.no_grad():
        one(frames[0])  # warm-up
        torch.cuda.synchronize()
        hits.zero_()
        n0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in frames:
            one(f)
            if fused and not staged:
                counts.append(net.last_render_samples)
        e1.record()
        torch.cuda.synchronize()
    samples = float(sum(int(c) for c in counts)) / max(len(counts), 1) if counts else None
    return {"ms_per_frame": e0.elapsed_time(e1) / len(frames), "samples_per_frame": samples,
            "bit_accuracy": float(hits) / len(frames) if extract_bits else None,
            "launches_per_frame": (_lib.launch_count - n0) / len(frames)}
