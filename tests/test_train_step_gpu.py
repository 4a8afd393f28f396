"""-m gpu tests of the training-step plumbing around the kernels: the fused message-table optimizer
against torch.optim.Adam (+ GradScaler), and the CUDA-graph-captured step against the eager step."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL = dict(bound=1.0, scale=0.8, dt_gamma=0.0, message_dim=4, num_rows=32, num_cols=32, H=128, W=128, num_rays=256,
             camera="blender", occupancy="sphere")


def _scene(**kw):
    from nerf_signature_b200 import harness
    return harness.Scene(dict(SMALL), torch.device("cuda:0"), seed=0, table_scale=300.0, **kw)


def _batches(scene, n):
    from nerf_signature_b200 import harness
    return [scene.to_device(harness.make_batch(scene.cfg, seed=50 + i)) for i in range(n)]


def _close_frac(x, y, rtol, atol, max_bad=1e-4):
    """Adam's update is lr * sign-like for tiny gradients (eps = 1e-15), and the scatter-add order of G is not
    deterministic, so an element whose gradient cancels to rounding noise may legitimately differ by ~lr:
    require all but a 1e-4 fraction of the elements to agree."""
    bad = ((x - y).abs() > atol + rtol * y.abs()).float().mean().item()
    assert bad <= max_bad, bad


def _bad_frac(x, y, rtol=1e-3, atol=1e-5):
    return ((x - y).abs() > atol + rtol * y.abs()).float().mean().item()


def _msg_tables(scene):
    return [e.weight.detach().clone() for e in scene.model.msg_encoder.embeddings]


@pytest.mark.parametrize("merged", [True, False])
def test_fused_adam_matches_torch_adam_with_gradscaler(merged):
    """Same seeds, same messages: WatermarkAdam (one kernel from G) == torch.optim.Adam on the fanned-out
    per-table gradients, including the untouched unselected tables and the per-table step counts.
    With one render call per step both optimizers see the SAME scatter-added dL/dS, so the losses agree to 1e-5.
    With the reference's two render calls the torch path sums two separately accumulated gradients (autograd) while the
    fused path accumulates both passes into one buffer: fp32 rounding differs, and Adam (eps 1e-15) turns a gradient
    that cancels to rounding noise into a +-lr update of that element - the losses then agree to 1e-3 only."""
    a, b = _scene(optimizer="torch", merged_render=merged), _scene(optimizer="fused", merged_render=merged)
    batches = _batches(a, 3)
    gen = torch.Generator().manual_seed(3)
    msgs = [a.new_message(gen) for _ in range(6)]
    tol = 1e-5 if merged else 1e-3
    for i, m in enumerate(msgs):
        la = a.train_step(batches[i % 3], m)
        lb = b.train_step(batches[i % 3], m)
        assert abs(float(la[0]) - float(lb[0])) <= tol * abs(float(la[0])) + 1e-7, i
    ta, tb = _msg_tables(a), _msg_tables(b)
    for x, y in zip(ta, tb):
        _close_frac(x, y, rtol=2e-4, atol=2e-6)
    # per-table step counts == number of times the table was selected
    want = np.zeros(2 * SMALL["message_dim"])
    for m in msgs:
        for i, bit in enumerate(m.tolist()):
            want[2 * i + int(bit)] += 1
    assert b.optimizer.steps.cpu().numpy().tolist() == want.tolist()
    for pa, pb in zip(a.model.msg_decoder.parameters(), b.model.msg_decoder.parameters()):
        _close_frac(pa, pb, rtol=1e-3, atol=1e-5, max_bad=1e-3)
    assert float(a.scaler.get_scale()) == float(b.scaler.get_scale())


def test_fused_adam_skips_step_on_inf():
    s = _scene(optimizer="fused")
    batch = _batches(s, 1)[0]
    msg = s.new_message(torch.Generator().manual_seed(0))
    s.train_step(batch, msg)
    before = _msg_tables(s)
    steps = s.optimizer.steps.clone()
    scale0 = float(s.scaler.get_scale())
    bad = dict(batch)
    bad["gt"] = batch["gt"].clone()
    bad["gt"][0, :, 0] = float("inf")  # every ray: the ones that hit the scene carry the inf into G
    s.train_step(bad, msg)
    after = _msg_tables(s)
    assert all(torch.equal(x, y) for x, y in zip(before, after))     # no update at all
    assert torch.equal(steps, s.optimizer.steps)
    assert float(s.scaler.get_scale()) == scale0 * 0.5               # GradScaler backed off


def test_graph_step_matches_eager_step():
    e, g = _scene(optimizer="fused"), _scene(optimizer="fused", graph=True)
    batches = _batches(e, 2)
    gen = torch.Generator().manual_seed(5)
    msgs = [e.new_message(gen) for _ in range(5)]
    # the capture warm-up runs 3 eager steps + the capture itself does not execute: replay the same
    # sequence on the eager scene so both have seen identical updates
    g._capture(batches[0], msgs[0])
    for _ in range(3):
        e.train_step(batches[0], msgs[0])
    for i, m in enumerate(msgs):
        le = e.train_step(batches[i % 2], m)
        lg = g.train_step(batches[i % 2], m)
        le, lg = [float(x) for x in le], [float(x) for x in lg]
        np.testing.assert_allclose(lg, le, rtol=1e-4, atol=1e-6)
    for x, y in zip(_msg_tables(e), _msg_tables(g)):
        _close_frac(x, y, rtol=1e-3, atol=1e-5)
    assert g.launches_per_step and g.launches_per_step > 10


def test_merged_render_matches_two_render_calls():
    """One render call over [block rays | content rays] (harness merged_render) is the same step as the reference's
    two calls: rays are independent, only the scatter order of dL/dS changes."""
    a, b = _scene(optimizer="fused"), _scene(optimizer="fused", merged_render=True)
    batches = _batches(a, 2)
    gen = torch.Generator().manual_seed(6)
    for i in range(4):
        m = a.new_message(gen)
        la = [float(x) for x in a.train_step(batches[i % 2], m)]
        lb = [float(x) for x in b.train_step(batches[i % 2], m)]
        np.testing.assert_allclose(lb, la, rtol=1e-4, atol=1e-6)
    assert a.samples_per_step() == b.samples_per_step()
    for x, y in zip(_msg_tables(a), _msg_tables(b)):
        _close_frac(x, y, rtol=1e-3, atol=1e-5)


def test_fused_decoder_step_matches_module_step():
    """The training step with the fused decoder kernels against the step with the plain PyTorch decoder: same losses
    on the first steps (both use float16-autocast arithmetic; they differ by summation order only)."""
    a, b = _scene(optimizer="fused"), _scene(optimizer="fused", fused_decoder=True)
    batches = _batches(a, 2)
    gen = torch.Generator().manual_seed(8)
    for i in range(3):
        m = a.new_message(gen)
        la = [float(x) for x in a.train_step(batches[i % 2], m)]
        lb = [float(x) for x in b.train_step(batches[i % 2], m)]
        np.testing.assert_allclose(lb, la, rtol=5e-3 if i else 2e-3, atol=1e-5)


def test_fused_loss_head_step_matches_torch_losses():
    """clamp / MSE / BCE / weighting through the loss-head kernels (nerf/loss_ops.py) against the plain torch
    expressions of utils_wtmk_disen.py:593,636-644, merged and split render, eager and captured."""
    for merged in (True, False):
        a = _scene(optimizer="fused", merged_render=merged)
        b = _scene(optimizer="fused", merged_render=merged, fused_losses=True)
        batches = _batches(a, 2)
        gen = torch.Generator().manual_seed(11)
        for i in range(4):
            m = a.new_message(gen)
            la = [float(x) for x in a.train_step(batches[i % 2], m)]
            lb = [float(x) for x in b.train_step(batches[i % 2], m)]
            np.testing.assert_allclose(lb, la, rtol=1e-4, atol=1e-6)
        for x, y in zip(_msg_tables(a), _msg_tables(b)):
            _close_frac(x, y, rtol=1e-3, atol=1e-5)


def test_loss_ops_against_torch():
    import torch.nn.functional as F
    from nerf_signature_b200.nerf.loss_ops import split_clamp, wtmk_loss
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    img = (torch.rand(1, 700, 3, generator=g) * 1.6 - 0.3).to(dev)   # values below 0 and above 1
    img[0, 5, 1] = 0.0
    img[0, 6, 2] = 1.0                                                # boundaries pass the gradient (min <= x <= max)
    gt = torch.rand(1, 400, 3, generator=g).to(dev)
    logits = torch.randn(7, 1, generator=g).to(dev)
    msg = torch.randint(0, 2, (7,), generator=g).float().to(dev)
    wp = torch.randn(300, 3, generator=g).to(dev)
    outs = []
    for fused in (False, True):
        x = img.clone().requires_grad_(True)
        z = logits.clone().requires_grad_(True)
        if fused:
            pred, content = split_clamp(x, 300)
            loss, lossi, lossw = wtmk_loss(content.view(1, 400, 3), gt, z, msg, 0.005, 1.0, 10.0)
        else:
            pred, content = torch.clamp(x[0, :300], min=0, max=1), x[:, 300:]
            lossi = F.mse_loss(content, gt, reduction="none").mean()
            lossw = F.binary_cross_entropy_with_logits(z * 10.0, msg.unsqueeze(-1), reduction="mean")
            loss = 0.005 * lossw + 1.0 * lossi
        ((loss + (pred * wp).sum()) * 128.0).backward()
        outs.append([t.detach().cpu().numpy() for t in (loss, lossi, lossw, pred, x.grad, z.grad)])
    for a, b in zip(*outs):
        np.testing.assert_allclose(b, a, rtol=2e-6, atol=1e-7)


def test_decoder_on_side_stream_matches_sequential_step():
    """overlap_decoder only changes WHERE the decoder kernels run (a side stream next to the content pass), eagerly and
    as parallel branches of the captured graph: same losses, same table updates."""
    a = _scene(optimizer="fused", fused_decoder=True, fused_losses=True)
    b = _scene(optimizer="fused", fused_decoder=True, fused_losses=True, overlap_decoder=True)
    c = _scene(optimizer="fused", fused_decoder=True, fused_losses=True, overlap_decoder=True, graph=True)
    assert b.overlap_decoder and c.overlap_decoder
    batches = _batches(a, 2)
    gen = torch.Generator().manual_seed(12)
    msgs = [a.new_message(gen) for _ in range(5)]
    c._capture(batches[0], msgs[0])
    for _ in range(3):          # the capture warm-up ran 3 eager steps on (batch 0, message 0)
        a.train_step(batches[0], msgs[0])
        b.train_step(batches[0], msgs[0])
    for i, m in enumerate(msgs):
        la = [float(x) for x in a.train_step(batches[i % 2], m)]
        lb = [float(x) for x in b.train_step(batches[i % 2], m)]
        lc = [float(x) for x in c.train_step(batches[i % 2], m)]
        np.testing.assert_allclose(lb, la, rtol=5e-3 if i else 1e-4, atol=1e-5)
        np.testing.assert_allclose(lc, la, rtol=5e-3 if i else 1e-4, atol=1e-5)
    # the content pass's field backward now scatters into dL/dS BEFORE the block pass's (it no longer waits for the
    # decoder), and the fp16 decoder backward sums with atomics: after 8 Adam steps (eps 1e-15: sign-like updates for
    # gradients at rounding-noise level) a few percent of the entries may sit one lr apart
    for x, y in zip(_msg_tables(a), _msg_tables(b)):
        _close_frac(x, y, rtol=1e-3, atol=1e-5, max_bad=5e-2)
    for x, y in zip(_msg_tables(a), _msg_tables(c)):
        _close_frac(x, y, rtol=1e-3, atol=1e-5, max_bad=5e-2)


def test_packed_pinned_batch_single_copy_matches_per_tensor_copies():
    """Scene.pinned_batch: the host batch packed like the captured step's static input buffer and moved with ONE H2D copy
    (bench.py's e2e path) drives exactly the same step as handing over the tensors one by one."""
    from nerf_signature_b200 import harness
    a = _scene(optimizer="fused", graph=True, merged_render=True, fused_decoder=True, fused_losses=True)
    b = _scene(optimizer="fused", graph=True, merged_render=True, fused_decoder=True, fused_losses=True)
    host = [harness.make_batch(a.cfg, seed=70 + i) for i in range(2)]
    dev = [a.to_device(h) for h in host]
    gen = torch.Generator().manual_seed(13)
    m0 = a.new_message(gen)
    a.train_step(dev[0], m0)
    b.train_step(dev[0], m0)          # captures; pinned_batch needs the static layout
    pool = [b.pinned_batch(h) for h in host]
    for i in range(4):
        m = a.new_message(gen)
        la = [float(x) for x in a.train_step(dev[i % 2], m)]
        pb = pool[i % 2]
        pb["message"].copy_(m)
        lb = [float(x) for x in b.train_step(pb, pb["message"])]
        np.testing.assert_allclose(lb, la, rtol=5e-3 if i else 1e-4, atol=1e-5)


def test_optimizer_state_dict_round_trip_and_torch_adam_layout():
    """WatermarkAdam.state_dict() carries the message tables' moments and step counts in the layout of
    torch.optim.Adam(model.get_params(lr)).state_dict() (what the reference Trainer checkpoints); a resumed optimizer
    continues exactly like the uninterrupted one, and a torch-Adam checkpoint loads."""
    from nerf_signature_b200.optim import WatermarkAdam
    a, b = _scene(optimizer="fused"), _scene(optimizer="fused")
    batches = _batches(a, 2)
    gen = torch.Generator().manual_seed(9)
    msgs = [a.new_message(gen) for _ in range(6)]
    for i in range(3):
        a.train_step(batches[i % 2], msgs[i]); b.train_step(batches[i % 2], msgs[i])
    sd = a.optimizer.state_dict()
    steps3 = a.optimizer.steps.cpu().numpy().copy()                       # state after exactly 3 steps, for the torch-Adam check
    md2 = 2 * SMALL["message_dim"]
    assert sd["param_groups"][0]["params"] == list(range(md2))
    touched = {t for t in range(md2) if float(a.optimizer.steps[t]) > 0}
    exp_avg3 = {t: a.optimizer.exp_avg[t].clone() for t in touched}
    assert touched and touched == {k for k in sd["state"] if k < md2}
    assert all({"step", "exp_avg", "exp_avg_sq"} <= set(sd["state"][t]) for t in touched)
    n_dec = len(list(a.model.msg_decoder.parameters()))
    assert sorted(k for k in sd["state"] if k >= md2) == list(range(md2, md2 + n_dec))
    # resume: a fresh optimizer on a copy of the model state
    c = _scene(optimizer="fused")
    c.model.load_state_dict(a.model.state_dict())
    c.optimizer.load_state_dict(sd)
    c.scaler.load_state_dict(a.scaler.state_dict())
    assert torch.equal(c.optimizer.steps, a.optimizer.steps)
    for p, q in zip(a.model.msg_decoder.parameters(), c.model.msg_decoder.parameters()):
        assert torch.equal(p.detach(), q.detach())
    # yardstick: `b` is an identical scene that was never interrupted - how far two runs of the same schedule drift apart
    # (scatter-add order + Adam's sign-like updates of near-zero gradients) bounds what a correct resume may deviate by
    for i in range(3, 6):
        la = [float(x) for x in a.train_step(batches[i % 2], msgs[i])]
        lb = [float(x) for x in b.train_step(batches[i % 2], msgs[i])]
        lc = [float(x) for x in c.train_step(batches[i % 2], msgs[i])]
        for k in range(3):
            floor = abs(lb[k] - la[k]) / abs(la[k])
            assert abs(lc[k] - la[k]) / abs(la[k]) <= max(1e-4, 3.0 * floor), (i, k, la, lb, lc)
    floor = max(_bad_frac(x, y) for x, y in zip(_msg_tables(b), _msg_tables(a)))
    worst = max(_bad_frac(x, y) for x, y in zip(_msg_tables(c), _msg_tables(a)))
    assert worst <= max(1e-4, 3.0 * floor), (worst, floor)
    # a torch.optim.Adam checkpoint over get_params (the reference's optimizer) has the same layout and loads
    t = _scene(optimizer="torch")
    for i in range(3):
        t.train_step(batches[i % 2], msgs[i])
    tsd = t.optimizer.state_dict()
    assert tsd["param_groups"][0]["params"] == list(range(md2))
    d = _scene(optimizer="fused")
    d.optimizer.load_state_dict(tsd)
    np.testing.assert_allclose(d.optimizer.steps.cpu().numpy(), steps3)
    for t_ in touched:
        torch.testing.assert_close(d.optimizer.exp_avg[t_], exp_avg3[t_], rtol=1e-3, atol=1e-7)


def test_lr_schedule_acts_on_graph_replays():
    """the reference steps a LambdaLR every iteration (scheduler_update_every_step); under CUDA-graph replay the
    learning rate must be read from a device scalar.  lr -> 0 must freeze the tables and the decoder; and the cached
    summed table must not survive a replay."""
    g = _scene(optimizer="fused", graph=True, merged_render=True, fused_decoder=True, fused_losses=True)
    batches = _batches(g, 2)
    gen = torch.Generator().manual_seed(3)
    g.train_step(batches[0], g.new_message(gen))            # capture + first replay at lr 1e-2
    assert g.model._S_cache is None                          # dropped after the replay
    before = _msg_tables(g)
    dec_before = [p.detach().clone() for p in g.model.msg_decoder.parameters()]
    sched = torch.optim.lr_scheduler.LambdaLR(g.optimizer, lambda it: 0.0)   # sets every group's lr to 0 immediately
    g.train_step(batches[1], g.new_message(gen))
    torch.cuda.synchronize()
    for x, y in zip(before, _msg_tables(g)):
        assert torch.equal(x, y)
    for x, y in zip(dec_before, g.model.msg_decoder.parameters()):
        assert torch.equal(x, y.detach())
    for grp in g.optimizer.param_groups:
        grp["lr"] = 1e-2                                     # what a scheduler step does
    g.train_step(batches[0], g.new_message(gen))
    torch.cuda.synchronize()
    assert any(not torch.equal(x, y) for x, y in zip(before, _msg_tables(g)))
    del sched


@pytest.mark.parametrize("graph,lookahead", [(False, True), (True, True), (True, False), (False, False)])
def test_deferred_optimizer_equals_sequential_optimizer(graph, lookahead):
    """harness defer_optimizer: the Adam update of step t issued at the start of step t+1 (next to its march) leaves, after
    flush_optimizer(), the same tables, decoder and scaler state as the sequential schedule, step for step; a flush in the
    middle (what update_extra_state needs) does not apply an update twice.  The scatter-add order of dL/dS is not
    deterministic and Adam's update is sign-like for near-zero gradients (see _close_frac), so the yardstick is the
    run-to-run difference of two IDENTICAL sequential scenes.
    lookahead: S for the next step comes from nsig_msg_adam_lookahead_sum and the table update itself runs next to
    the decoder (between the composite forward and the field backward).  Without it (default) the pending update's kernel
    also accumulates S for the step that issues it (nsig_msg_adam_step_sum)."""
    kw = dict(optimizer="fused", merged_render=True, fused_decoder=True, fused_losses=True, graph=graph)
    a, a2, b = _scene(**kw), _scene(**kw), _scene(defer_optimizer=True, lookahead=lookahead, **kw)
    assert b.defer_optimizer and b.lookahead == lookahead and b.fuse_table_sum == (not lookahead)
    batches = _batches(a, 2)
    gen = torch.Generator().manual_seed(21)
    msgs = [a.new_message(gen) for _ in range(7)]

    def check_tables():
        floor = max(_bad_frac(x, y) for x, y in zip(_msg_tables(a2), _msg_tables(a)))
        worst = max(_bad_frac(x, y) for x, y in zip(_msg_tables(b), _msg_tables(a)))
        assert worst <= max(1e-4, 3.0 * floor), (worst, floor)

    for i, m in enumerate(msgs):
        la = [float(x) for x in a.train_step(batches[i % 2], m)]
        la2 = [float(x) for x in a2.train_step(batches[i % 2], m)]
        lb = [float(x) for x in b.train_step(batches[i % 2], m)]
        for k in range(3):
            floor = abs(la2[k] - la[k]) / abs(la[k])
            assert abs(lb[k] - la[k]) / abs(la[k]) <= max(2e-4, 3.0 * floor), (i, k, la, la2, lb)
        if i == 3:
            b.flush_optimizer()
            b.flush_optimizer()      # idempotent
            check_tables()
    b.flush_optimizer()
    torch.cuda.synchronize()
    check_tables()
    floor = max(_bad_frac(p.detach(), q.detach()) for p, q in zip(a2.model.msg_decoder.parameters(), a.model.msg_decoder.parameters()))
    worst = max(_bad_frac(p.detach(), q.detach()) for p, q in zip(b.model.msg_decoder.parameters(), a.model.msg_decoder.parameters()))
    assert worst <= max(1e-3, 3.0 * floor), (worst, floor)
    np.testing.assert_allclose(b.optimizer.steps.cpu().numpy(), a.optimizer.steps.cpu().numpy())
    assert a.scaler.get_scale() == b.scaler.get_scale()


@pytest.mark.parametrize("skip,shard", [(False, False), (True, False), (False, True)])
def test_lookahead_sum_is_bit_identical_to_update_then_sum(skip, shard):
    """nsig_msg_adam_lookahead_sum(applied, next) == nsig_msg_adam_step(applied) followed by nsig_msg_table_sum(next), bit for
    bit (same adam_elem, same accumulation order), without having touched tables or moments; the update that follows with
    steps_prepared=1 leaves exactly what a stand-alone nsig_msg_adam_step leaves (steps, tables, moments).  Also with a
    skipped step (found_inf = 1) and on a slice of the tables (sharded optimizer)."""
    from nerf_signature_b200 import _lib
    P = _lib.ptr
    dev = torch.device("cuda:0")
    md, log2_T = 6, 12
    n = 2 << log2_T
    g = torch.Generator(device="cuda").manual_seed(3)

    def state():
        gg = torch.Generator(device="cuda").manual_seed(11)
        tabs = [torch.randn(n, device=dev, generator=gg) * 1e-2 for _ in range(2 * md)]
        ms = [torch.randn(n, device=dev, generator=gg) * 1e-3 for _ in range(2 * md)]
        vs = [torch.rand(n, device=dev, generator=gg) * 1e-6 for _ in range(2 * md)]
        ptrs = torch.tensor([[t.data_ptr() for t in grp] for grp in (tabs, ms, vs)], dtype=torch.int64, device=dev)
        steps = torch.arange(2 * md, dtype=torch.float32, device=dev) + 3.0
        coef = torch.zeros(2 * md, 2, dtype=torch.float32, device=dev)
        return tabs, ms, vs, ptrs, steps, coef

    G = torch.randn(n, device=dev, generator=g) * 65536.0 * 1e-3
    G[torch.rand(n, device=dev, generator=g) < 0.5] = 0.0
    applied = torch.tensor([1, 0, 1, 1, 0, 0], dtype=torch.float32, device=dev)
    nxt = torch.tensor([1, 1, 0, 1, 0, 1], dtype=torch.float32, device=dev)
    scale = torch.tensor([65536.0], device=dev)
    finf = torch.tensor([1.0 if skip else 0.0], device=dev)
    lo, cnt = (1024, 2048) if shard else (0, 0)
    hyper = (1e-2, 0.9, 0.99, 1e-15)

    # reference order: update, then sum
    tabs, ms, vs, ptrs, steps, coef = state()
    _lib.call("nsig_msg_adam_step", P(ptrs), 2 * md, md, P(applied), P(G), P(steps), P(coef), P(scale), P(finf), *hyper,
              log2_T, None, lo, cnt, 0)
    S_ref = torch.full((n,), 7.0, device=dev)
    _lib.call("nsig_msg_table_sum", _lib.pointer_array(tabs), md, P(nxt), log2_T, P(S_ref), lo, cnt)
    # look-ahead order
    tabs2, ms2, vs2, ptrs2, steps2, coef2 = state()
    before = [t.clone() for t in tabs2 + ms2 + vs2]
    S = torch.full((n,), 7.0, device=dev)
    _lib.call("nsig_msg_adam_lookahead_sum", P(ptrs2), 2 * md, md, P(applied), P(nxt), P(G), P(steps2), P(coef2), P(scale),
              P(finf), *hyper, log2_T, None, lo, cnt, P(S))
    for x, y in zip(before, tabs2 + ms2 + vs2):
        assert torch.equal(x, y)                      # read-only
    assert torch.equal(S, S_ref)                      # incl. the untouched 7.0 outside the slice
    if not skip:
        assert not torch.equal(S_ref[lo:lo + (cnt or n)], sum(tabs2[2 * i + int(nxt[i])] for i in range(md))[lo:lo + (cnt or n)])
    _lib.call("nsig_msg_adam_step", P(ptrs2), 2 * md, md, P(applied), P(G), P(steps2), P(coef2), P(scale), P(finf), *hyper,
              log2_T, None, lo, cnt, 1)
    assert torch.equal(steps, steps2) and torch.equal(coef, coef2)
    for x, y in zip(tabs + ms + vs, tabs2 + ms2 + vs2):
        assert torch.equal(x, y)


@pytest.mark.parametrize("skip,shard", [(False, False), (True, False), (False, True)])
def test_adam_step_sum_is_bit_identical_to_update_then_sum(skip, shard):
    """nsig_msg_adam_step_sum(applied, next) == nsig_msg_adam_step(applied) followed by nsig_msg_table_sum(next): tables,
    moments, steps, coefficients and S, bit for bit - also with a skipped step (found_inf = 1: tables untouched, S of the
    tables as they are) and on a slice (sharded optimizer: nothing outside the slice is written, S included)."""
    from nerf_signature_b200 import _lib
    P = _lib.ptr
    dev = torch.device("cuda:0")
    md, log2_T = 7, 12          # odd message_dim: the unrolled loop's tail
    n = 2 << log2_T
    g = torch.Generator(device="cuda").manual_seed(5)

    def state():
        gg = torch.Generator(device="cuda").manual_seed(13)
        tabs = [torch.randn(n, device=dev, generator=gg) * 1e-2 for _ in range(2 * md)]
        ms = [torch.randn(n, device=dev, generator=gg) * 1e-3 for _ in range(2 * md)]
        vs = [torch.rand(n, device=dev, generator=gg) * 1e-6 for _ in range(2 * md)]
        ptrs = torch.tensor([[t.data_ptr() for t in grp] for grp in (tabs, ms, vs)], dtype=torch.int64, device=dev)
        steps = torch.arange(2 * md, dtype=torch.float32, device=dev) + 3.0
        coef = torch.zeros(2 * md, 2, dtype=torch.float32, device=dev)
        return tabs, ms, vs, ptrs, steps, coef

    G = torch.randn(n, device=dev, generator=g) * 65536.0 * 1e-3
    G[torch.rand(n, device=dev, generator=g) < 0.5] = 0.0
    applied = torch.tensor([1, 0, 1, 1, 0, 0, 1], dtype=torch.float32, device=dev)
    nxt = torch.tensor([1, 1, 0, 1, 0, 1, 1], dtype=torch.float32, device=dev)
    scale = torch.tensor([65536.0], device=dev)
    finf = torch.tensor([1.0 if skip else 0.0], device=dev)
    lo, cnt = (1024, 2048) if shard else (0, 0)
    hyper = (1e-2, 0.9, 0.99, 1e-15)

    tabs, ms, vs, ptrs, steps, coef = state()
    before = [t.clone() for t in tabs + ms + vs]
    _lib.call("nsig_msg_adam_step", P(ptrs), 2 * md, md, P(applied), P(G), P(steps), P(coef), P(scale), P(finf), *hyper,
              log2_T, None, lo, cnt, 0)
    S_ref = torch.full((n,), 7.0, device=dev)
    _lib.call("nsig_msg_table_sum", _lib.pointer_array(tabs), md, P(nxt), log2_T, P(S_ref), lo, cnt)

    tabs2, ms2, vs2, ptrs2, steps2, coef2 = state()
    S = torch.full((n,), 7.0, device=dev)
    _lib.call("nsig_msg_adam_step_sum", P(ptrs2), 2 * md, md, P(applied), P(nxt), P(G), P(steps2), P(coef2), P(scale),
              P(finf), *hyper, log2_T, None, lo, cnt, P(S))
    assert torch.equal(S, S_ref)
    assert torch.equal(steps, steps2) and torch.equal(coef, coef2)
    for x, y in zip(tabs + ms + vs, tabs2 + ms2 + vs2):
        assert torch.equal(x, y)
    changed = sum(int(not torch.equal(x, y)) for x, y in zip(before, tabs2 + ms2 + vs2))
    assert changed == (0 if skip else 3 * md)
    # argument validation
    with pytest.raises(_lib.NsigError):
        _lib.call("nsig_msg_adam_step_sum", P(ptrs2), 2 * md, md, P(applied), None, P(G), P(steps2), P(coef2), P(scale),
                  P(finf), *hyper, log2_T, None, lo, cnt, P(S))
    with pytest.raises(_lib.NsigError):
        _lib.call("nsig_msg_adam_step_sum", P(ptrs2), 2 * md, md, P(applied), P(nxt), P(G), P(steps2), P(coef2), P(scale),
                  P(finf), *hyper, log2_T, None, 2, 6, P(S))


def test_premarched_render_is_bit_identical():
    """NeRFRenderer.march_ahead + run_cuda(premarched=...) == run_cuda marching by itself: same kernels, same arguments."""
    from nerf_signature_b200 import harness
    s = _scene(optimizer="fused", merged_render=True)
    model = s.model
    b = harness.make_batch(s.cfg, seed=77)
    o = torch.from_numpy(b["rays_o"]).cuda()
    d = torch.from_numpy(b["rays_d"]).cuda()
    msg = s.new_message(torch.Generator().manual_seed(1)).cuda()
    kw = dict(staged=False, bg_color=1, perturb=False, force_all_rays=True, dt_gamma=0.0, max_steps=1024, T_thresh=1e-4)
    with torch.no_grad():
        ref = model.render(o, d, msg, **kw)
        rows = model.step_counter[(model.local_step - 1) % 16].tolist()
        bufs = model.march_buffers(o.shape[1])
        model.march_ahead(o, d, bufs)
        out = model.render(o, d, msg, premarched=bufs, **kw)
        # grid-limited march (nsig_march_rays_train_limited: 5 CTAs walk all the rays) - the same bits
        lim = model.march_buffers(o.shape[1])
        model.march_ahead(o, d, lim, max_blocks=5)
    assert bufs["counter"].tolist() == rows
    for k in ("image", "depth", "weights_sum"):
        assert torch.equal(out[k], ref[k]), k
    m = rows[0]
    assert lim["counter"].tolist() == rows and torch.equal(lim["rays"], bufs["rays"])
    for k in ("nears", "fars"):
        assert torch.equal(lim[k], bufs[k]), k
    for k in ("xyzs", "dirs", "deltas"):
        assert torch.equal(lim[k][:m], bufs[k][:m]), k
    with pytest.raises(ValueError):
        model.render(o[:, :100], d[:, :100], msg, premarched=bufs, **kw)
    with pytest.raises(ValueError):
        model.march_ahead(o[:, :100], d[:, :100], bufs)


def test_march_ahead_schedule_equals_plain_schedule():
    """harness march_ahead: the samples of batch t+1 are generated inside step t (next to the decoder), from a second input
    buffer, by a second graph - losses, sample counts, tables and decoder must be those of the plain captured step on the same
    batch / message sequence, whether a batch was announced (overlapped march) or not (marched on the spot).  Yardstick: two
    identical plain scenes (the scatter-add order of dL/dS is not deterministic)."""
    kw = dict(optimizer="fused", merged_render=True, fused_decoder=True, fused_losses=True, graph=True, defer_optimizer=True)
    a, a2, b = _scene(**kw), _scene(**kw), _scene(march_ahead=True, **kw)
    batches = _batches(a, 3)
    gen = torch.Generator().manual_seed(31)
    msgs = [a.new_message(gen) for _ in range(9)]
    # capture warm-ups: 3 steps (plain) / 4 steps (march_ahead, both buffers) on (batches[0], msgs[0])
    a._capture(batches[0], msgs[0]); a2._capture(batches[0], msgs[0])
    b._capture(batches[0], msgs[0], batches[0], msgs[0])
    a.train_step(batches[0], msgs[0]); a2.train_step(batches[0], msgs[0])
    for i, m in enumerate(msgs):
        cur = batches[i % 3]
        la = [float(x) for x in a.train_step(cur, m)]
        la2 = [float(x) for x in a2.train_step(cur, m)]
        if i in (4, 5) or i + 1 == len(msgs):   # nothing announced for steps 5 and 6: they are staged + marched on the spot
            lb = b.train_step(cur, m)
        else:
            lb = b.train_step(cur, m, next_batch=batches[(i + 1) % 3], next_message=msgs[i + 1])
        lb = [float(x) for x in lb]
        assert b.samples_per_step() == a.samples_per_step(), i
        for k in range(3):
            floor = abs(la2[k] - la[k]) / abs(la[k])
            assert abs(lb[k] - la[k]) / abs(la[k]) <= max(2e-4, 3.0 * floor), (i, k, la, la2, lb)
    for s in (a, a2, b):
        s.flush_optimizer()
    torch.cuda.synchronize()
    floor = max(_bad_frac(x, y) for x, y in zip(_msg_tables(a2), _msg_tables(a)))
    worst = max(_bad_frac(x, y) for x, y in zip(_msg_tables(b), _msg_tables(a)))
    assert worst <= max(1e-4, 3.0 * floor), (worst, floor)
    with pytest.raises(ValueError):
        _scene(march_ahead=True, optimizer="fused", merged_render=True, graph=False)
