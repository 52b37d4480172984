"""ELBO train step on the GPU (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerances (stated): the CUDA path computes in bf16 operands / fp32 accumulate.
 * against the oracle in bf16-emulation mode (rounds where the CUDA path rounds): logits within
   1.5e-2 * max|logit| (a bf16 ulp is 0.8 %), mu/std 5e-3 rel, rec_loss 2e-3 rel, KL 5e-3 rel;
 * against the exact fp64 oracle: logits 2e-2 * max(1, max|logit|), rec_loss 5e-3 rel, KL 1e-2 rel;
 * gradients (bf16 dlogits / gate gradients): 4e-2 of the tensor's max-abs, global norm 2e-2 rel.
"""
import numpy as np
import pytest
import torch

from helpers import O, SMALL, TINY, engine_for, feed_of, make_case, rel_err, rng_for

pytestmark = pytest.mark.gpu


def run_forward(sizes, B, T, ragged, **kw):
    cfg, params, batch = make_case(sizes, B, T, seed=3, ragged=ragged, **kw)
    eng = engine_for(cfg, params, B, T)
    N = B * cfg.num_captions
    out = eng.eval_step(rng=rng_for(batch), **feed_of(batch))
    taps = eng.debug_taps(N, T, z=True)
    with torch.no_grad():
        ref_e = O.forward(params, cfg, batch, emulate=True)
        ref_x = O.forward(params, cfg, batch, emulate=False)
    eng.close()
    ref_x["n_tokens"] = float((batch["cap_lbl"] != 0).sum())
    return cfg, out, taps, ref_e, ref_x


@pytest.mark.parametrize("sizes,B,T,ragged", [(TINY, 2, 5, False), (TINY, 3, 6, True), (SMALL, 4, 7, True),
                                              (SMALL, 30, 9, False)])
def test_forward_normal_prior(sizes, B, T, ragged):
    cfg, out, taps, ref_e, ref_x = run_forward(sizes, B, T, ragged)
    lg = ref_e["logits"].numpy()
    assert rel_err(taps["logits"], lg) <= 1.5e-2
    assert rel_err(taps["mu"], ref_e["mu"].numpy()) <= 5e-3
    assert rel_err(taps["std"], ref_e["std"].numpy()) <= 5e-3
    assert abs(out["rec_loss"] - float(ref_e["rec_loss"])) <= 2e-3 * abs(float(ref_e["rec_loss"]))
    assert abs(out["kld"] - float(ref_e["kld"])) <= 5e-3 * abs(float(ref_e["kld"])) + 1e-6
    # exact fp64 oracle, stated bf16 tolerance
    lx = ref_x["logits"].numpy()
    assert np.max(np.abs(taps["logits"] - lx)) <= 2e-2 * max(1.0, np.max(np.abs(lx)))
    assert abs(out["rec_loss"] - float(ref_x["rec_loss"])) <= 5e-3 * abs(float(ref_x["rec_loss"]))
    assert abs(out["kld"] - float(ref_x["kld"])) <= 1e-2 * abs(float(ref_x["kld"])) + 1e-6
    assert abs(out["lower_bound"] - float(ref_x["lower_bound"])) <= 5e-3 * abs(float(ref_x["lower_bound"]))
    assert out["n_tokens"] == ref_x["n_tokens"]


@pytest.mark.parametrize("prior,use_c_v,sizes,B,T", [("GMM", True, TINY, 3, 5), ("AG", True, TINY, 3, 5), ("AG", False, SMALL, 4, 6),
                                                     ("GMM", False, SMALL, 4, 6), ("AG", True, SMALL, 30, 7)])
def test_forward_gmm_ag_priors(prior, use_c_v, sizes, B, T):
    """90 (mu_k, logstd_k) heads; GMM gathers the drawn cluster, AG mixes with the cluster vector (encoder.py:71-107);
    AG's KL is a per-row vector against N(c_v @ c_means, 0.1) (main.py:136-145, Q2)."""
    cfg, out, taps, ref_e, ref_x = run_forward(sizes, B, T, True, prior=prior, use_c_v=use_c_v)
    assert rel_err(taps["mu"], ref_e["mu"].numpy()) <= 5e-3
    assert rel_err(taps["std"], ref_e["std"].numpy()) <= 5e-3
    assert rel_err(taps["logits"], ref_e["logits"].numpy()) <= 1.5e-2
    kl_ref = ref_x["kld"].numpy()
    if prior == "AG":
        assert rel_err(taps["kl_rows"], kl_ref) <= 1e-2
    assert abs(out["kld"] - float(kl_ref.mean())) <= 1e-2 * abs(float(kl_ref.mean())) + 1e-6
    assert abs(out["rec_loss"] - float(ref_x["rec_loss"])) <= 5e-3 * abs(float(ref_x["rec_loss"]))
    lb = float(ref_x["lower_bound"].mean())
    assert abs(out["lower_bound"] - lb) <= 5e-3 * abs(lb)


def test_padding_semantics():
    """dynamic_rnn(sequence_length): logits past the caption end equal the logits bias exactly."""
    cfg, params, batch = make_case(TINY, 3, 6, seed=5, ragged=True)
    eng = engine_for(cfg, params, 3, 6)
    N, T = 9, 6
    eng.eval_step(rng=rng_for(batch), **feed_of(batch))
    lg = eng.debug_taps(N, T)["logits"].reshape(N, T, -1)
    b_o = params["decoder/rnn_logits/bias"].to(torch.bfloat16).float().numpy()
    ln = batch["lengths"].numpy()
    for n in range(N):
        for t in range(int(ln[n]), T):
            np.testing.assert_array_equal(lg[n, t], b_o)
    eng.close()


def test_no_encoder_and_dropout_masks():
    for kw in (dict(no_encoder=True), dict(dec_keep_rate=0.7, dec_lstm_drop=0.8), dict(use_c_v=True)):
        cfg, out, taps, ref_e, ref_x = run_forward(SMALL, 3, 6, True, **kw)
        assert rel_err(taps["logits"], ref_e["logits"].numpy()) <= 1.5e-2, kw
        assert abs(out["rec_loss"] - float(ref_x["rec_loss"])) <= 5e-3 * abs(float(ref_x["rec_loss"])), kw


def grads_case(sizes, B, T, ragged, **kw):
    cfg, params, batch = make_case(sizes, B, T, seed=11, ragged=ragged, **kw)
    eng = engine_for(cfg, params, B, T)
    f = feed_of(batch)
    dev = lambda a, dt: torch.tensor(np.ascontiguousarray(a)).to(dt).cuda()
    eng.forward_backward_device(dev(f["image_f_inputs"], torch.float32), dev(f["ann_inputs_enc"], torch.int32),
                                dev(f["ann_inputs_dec"], torch.int32), dev(f["ann_lengths"], torch.int32), 0,
                                c_i=dev(f["c_i"], torch.float32) if f["c_i"] is not None else None, rng=rng_for(batch))
    torch.cuda.synchronize()
    res, grads, gnorm = O.compute_grads(params, cfg, batch, emulate=False)
    return cfg, eng, grads, gnorm


@pytest.mark.parametrize("sizes,B,T,ragged,kw", [
    (TINY, 2, 5, False, {}), (SMALL, 4, 7, True, {}), (SMALL, 3, 6, True, dict(use_c_v=True)),
    (SMALL, 3, 5, True, dict(no_encoder=True)), (SMALL, 3, 5, True, dict(dec_keep_rate=0.7, dec_lstm_drop=0.8)),
    (SMALL, 4, 6, False, dict(ann_param=3.0)), (TINY, 3, 5, True, dict(prior="GMM", use_c_v=True)),
    (TINY, 3, 5, True, dict(prior="AG", use_c_v=True)), (SMALL, 4, 6, True, dict(prior="AG")),
    (SMALL, 6, 6, False, dict(prior="GMM", use_c_v=True, ann_param=2.0))])
def test_gradients(sizes, B, T, ragged, kw):
    cfg, eng, grads, gnorm = grads_case(sizes, B, T, ragged, **kw)
    worst = {}
    for name, g in grads.items():
        if g is None:
            continue
        got = eng.get_gradient(name)
        ref = g.numpy()
        scale = max(np.max(np.abs(ref)), 1e-12)
        worst[name] = float(np.max(np.abs(got - ref)) / scale)
    out = eng.apply_gradients(1.0)
    eng.close()
    bad = {k: v for k, v in worst.items() if v > 4e-2}
    assert not bad, bad
    assert abs(out["global_norm"] - gnorm) <= 2e-2 * gnorm


def test_adam_three_steps():
    """Post-update parameters after 1 and 3 steps (TF-form Adam, beta1 = 0.8, clip 5.0; Q4/Q5)."""
    cfg, params, batch = make_case(SMALL, 4, 6, seed=21, ragged=True)
    eng = engine_for(cfg, params, 4, 6)
    p_ref = {k: v.clone() for k, v in params.items()}
    opt = {"t": 0, "m": {}, "v": {}}
    for step in range(3):
        out = eng.train_step(anneal=step, rng=rng_for(batch), **feed_of(batch))
        b2 = dict(batch)
        b2["global_step"] = step
        ref = O.train_step(p_ref, opt, cfg, b2)
        assert abs(out["rec_loss"] - ref["rec_loss"]) <= 5e-3 * abs(ref["rec_loss"]), step
        assert abs(out["kld"] - float(ref["kld"])) <= 2e-2 * abs(float(ref["kld"])) + 1e-6, step
        assert abs(out["global_norm"] - ref["global_norm"]) <= 3e-2 * ref["global_norm"], step
    # Adam's first steps move every weight by ~lr * sign(g): compare the *update* direction and size
    lr = cfg.learning_rate
    for name in ("decoder/rnn_logits/kernel", "decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel", "imf_emb/kernel",
                 "encoder/dense/kernel", "decoder/net/z_rnn/kernel", "encoder/enc_embeddings"):
        got = eng.get_variable(name).astype(np.float64) - params[name].numpy()
        want = p_ref[name].numpy() - params[name].numpy()
        # updates are bounded by ~3 * lr each; gradient noise from bf16 perturbs small-|g| entries, so compare in L2
        num = np.linalg.norm(got - want)
        den = np.linalg.norm(want)
        assert num <= 0.15 * den, (name, num, den)
        assert np.max(np.abs(got)) <= 3.5 * 3 * lr
    eng.close()


@pytest.mark.parametrize("kind", ["SGD", "Momentum"])
def test_sgd_and_momentum_three_steps(kind):
    """--optimizer SGD / Momentum (ops/optimizers.py:33-36, 41-46): plain / momentum-0.9 updates of the clipped gradient with
    the staircase-halved rate (decay period of 2 steps here, so steps 1-2 use lr and step 3 uses lr / 2). Exact host
    replay of the device's own fetched gradients pins the update rule; the oracle's 3-step trajectory pins the whole step."""
    cfg, params, batch = make_case(SMALL, 4, 6, seed=23, ragged=True, optimizer=kind, learning_rate=0.05,
                                   batch_size=4, num_ex_per_epoch=8, num_epochs_per_decay=1)
    assert O.decay_steps(cfg) == 1 or O.decay_steps(cfg) == 2
    eng = engine_for(cfg, params, 4, 6)
    assert eng.cfg.lr_decay_steps == O.decay_steps(cfg)
    p_ref = {k: v.clone() for k, v in params.items()}
    opt = {"t": 0, "m": {}, "v": {}}
    name = "decoder/rnn_logits/bias"
    acc = np.zeros(params[name].shape)
    host = eng.get_variable(name).astype(np.float64)
    for step in range(3):
        out = eng.train_step(anneal=step, rng=rng_for(batch), **feed_of(batch))
        b2 = dict(batch)
        b2["global_step"] = step
        ref = O.train_step(p_ref, opt, cfg, b2)
        assert abs(out["rec_loss"] - ref["rec_loss"]) <= 5e-3 * abs(ref["rec_loss"]), step
        assert abs(out["global_norm"] - ref["global_norm"]) <= 3e-2 * ref["global_norm"], step
        g = eng.get_gradient(name).astype(np.float64) * min(1.0, cfg.lstm_clip_by_norm / out["global_norm"])
        lr_d = cfg.learning_rate * 0.5 ** (step // O.decay_steps(cfg))
        acc = 0.9 * acc + g if kind == "Momentum" else g
        host = host - lr_d * acc
        np.testing.assert_allclose(eng.get_variable(name), host, rtol=0, atol=1e-5 * max(1.0, np.max(np.abs(host))))
    for nm in ("decoder/rnn_logits/kernel", "decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel", "imf_emb/kernel",
               "encoder/dense/kernel", "encoder/enc_embeddings"):
        got = eng.get_variable(nm).astype(np.float64) - params[nm].numpy()
        want = p_ref[nm].numpy() - params[nm].numpy()
        assert np.linalg.norm(got - want) <= 0.08 * np.linalg.norm(want), (nm, kind)
    eng.close()


def test_unknown_optimizer_is_rejected():
    cfg, params, batch = make_case(TINY, 2, 5, seed=2)
    cfg.optimizer = "RMSProp"
    with pytest.raises(ValueError):
        engine_for(cfg, params, 2, 5)


def test_decoder_dropout_flags_draw_philox_masks():
    """--dec_drop / --dec_lstm_drop without explicit masks (the CLI path): keep masks come from Philox(seed, step)
    on the device. Same seed and step -> same step; another seed -> another draw; the loss stays close to the oracle's
    with ITS explicit masks (same distribution, different draw) and the gradient check of the explicit-mask case above
    covers the arithmetic."""
    cfg, params, batch = make_case(SMALL, 6, 7, seed=31, ragged=True, dec_keep_rate=0.7, dec_lstm_drop=0.8)
    outs = []
    for seed in (5, 5, 6):
        eng = engine_for(cfg, params, 6, 7)
        outs.append(eng.train_step(anneal=0, rng={"seed": seed, "eps": rng_for(batch)["eps"]}, **feed_of(batch)))
        eng.close()
    assert abs(outs[0]["rec_loss"] - outs[1]["rec_loss"]) <= 1e-5 * abs(outs[0]["rec_loss"])
    assert abs(outs[0]["rec_loss"] - outs[2]["rec_loss"]) > 1e-6 * abs(outs[0]["rec_loss"])
    with torch.no_grad():
        ref = O.forward(params, cfg, batch)
    assert abs(outs[0]["rec_loss"] - float(ref["rec_loss"])) <= 0.05 * abs(float(ref["rec_loss"]))
    assert np.isfinite(outs[0]["global_norm"]) and outs[0]["global_norm"] > 0


def test_vocabulary_wider_than_the_register_resident_ce_kernel():
    """V = 13001 > 12288 takes the looped cross-entropy kernel: loss, logits and the vocabulary gradients as
    in the small cases."""
    sizes = dict(TINY)
    sizes["vocab_size"] = 13001
    cfg, eng, grads, gnorm = grads_case(sizes, 2, 4, True)
    for name in ("decoder/rnn_logits/kernel", "decoder/rnn_logits/bias", "decoder/net/dec_embeddings", "imf_emb/kernel"):
        ref = grads[name].numpy()
        got = eng.get_gradient(name)
        assert np.max(np.abs(got - ref)) <= 4e-2 * max(np.max(np.abs(ref)), 1e-12), name
    out = eng.apply_gradients(1.0)
    assert abs(out["global_norm"] - gnorm) <= 2e-2 * gnorm
    eng.close()


def test_clip_identity_and_adam_first_step_kat():
    """Known-answer: with |g| <= clip the clip is the identity, and Adam's first step is
    lr_t * (1-b1) g / (sqrt((1-b2) g^2) + eps) with lr_t = lr sqrt(1-b2)/(1-b1)  ~=  lr * sign(g)."""
    cfg, params, batch = make_case(TINY, 2, 5, seed=2)
    eng = engine_for(cfg, params, 2, 5)
    before = eng.get_variable("decoder/rnn_logits/bias").astype(np.float64)
    out = eng.train_step(anneal=0, rng=rng_for(batch), **feed_of(batch))
    assert out["global_norm"] < cfg.lstm_clip_by_norm
    g = eng.get_gradient("decoder/rnn_logits/bias").astype(np.float64)
    after = eng.get_variable("decoder/rnn_logits/bias").astype(np.float64)
    b1, b2, eps, lr = 0.8, 0.999, 1e-8, cfg.learning_rate
    lr_t = lr * np.sqrt(1 - b2) / (1 - b1)
    want = before - lr_t * (1 - b1) * g / (np.sqrt((1 - b2) * g * g) + eps)
    np.testing.assert_allclose(after, want, rtol=0, atol=2e-3 * lr)
    eng.close()


def test_error_behaviour():
    cfg, params, batch = make_case(TINY, 2, 5, seed=2)
    eng = engine_for(cfg, params, 2, 5)
    f = feed_of(batch)
    with pytest.raises(ValueError):
        eng.set_variable("no/such/variable", np.zeros(3))
    with pytest.raises(ValueError):
        eng.set_variable("imf_emb/bias", np.zeros(3))
    bad = dict(f)
    bad["ann_inputs_enc"] = f["ann_inputs_enc"][:, :3]
    with pytest.raises(ValueError):
        eng.train_step(anneal=0, rng=rng_for(batch), **bad)
    big = make_case(TINY, 4, 5, seed=2)[2]
    with pytest.raises(ValueError):  # exceeds the handle's max_batch
        eng.train_step(anneal=0, rng=rng_for(big), **feed_of(big))
    eng.close()


@pytest.mark.parametrize("kw", [dict(), dict(prior="AG", use_c_v=True)])
def test_double_buffered_feed_equals_plain_steps(kw):
    """vc_stage_batch / vc_train_step_staged (H2D of batch i+1 on a copy stream while step i computes, two staging
    slots) must give what vc_train_step gives on the same sequence of batches: same scalars every step, same
    variables after 4 steps (identical kernels on identical inputs; fp32 atomics order is the only freedom)."""
    B, T = 4, 7
    cfg, params, _ = make_case(SMALL, B, T, seed=11, **kw)
    batches = [O.synthetic_batch(cfg, B, T, seed=20 + i, dtype=torch.float64, ragged=True) for i in range(4)]
    feeds = [feed_of(b) for b in batches]
    if "c_i" in feeds[0] and feeds[0]["c_i"] is None:
        for f in feeds:
            f.pop("c_i")
    plain = engine_for(cfg, params, B, T)
    want = [plain.train_step(anneal=i, rng={"seed": 5 + i}, **feeds[i]) for i in range(4)]
    staged = engine_for(cfg, params, B, T)
    pinned = [{k: (staged.pinned(v) if v is not None else None) for k, v in f.items()} for f in feeds]
    got = []
    staged.stage_batch(0, **pinned[0])
    for i in range(4):
        if i + 1 < 4:
            staged.stage_batch((i + 1) & 1, **pinned[i + 1])
        if i % 2 == 0:
            got.append(staged.train_step_staged(i & 1, anneal=i, rng={"seed": 5 + i}))
        else:  # the data-parallel split of the same step (the all-reduce would sit between the two calls)
            staged.forward_backward_staged(i & 1, anneal=i, rng={"seed": 5 + i})
            got.append(staged.apply_gradients(1.0))
    for a, b in zip(got, want):
        for k in ("rec_loss", "kld", "lower_bound", "global_norm", "n_tokens", "annealing"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k])), (k, a[k], b[k])
    for name in ("decoder/rnn_logits/kernel", "encoder/enc_embeddings", "imf_emb/kernel"):
        # Adam turns summation-order noise on near-zero gradients into O(lr) differences on single entries, so the
        # updates are compared in L2
        d_staged = staged.get_variable(name).astype(np.float64) - params[name].numpy()
        d_plain = plain.get_variable(name).astype(np.float64) - params[name].numpy()
        assert np.linalg.norm(d_staged - d_plain) <= 2e-2 * np.linalg.norm(d_plain), name
    with pytest.raises(Exception):
        staged.train_step_staged(0, anneal=9)  # slot 0 was consumed and not refilled
    with pytest.raises(ValueError):
        staged.stage_batch(2, **pinned[0])
    plain.close()
    staged.close()


def test_deferred_step_results_equal_synchronous_ones():
    """vc_step_result_queue / vc_step_result: the host enqueues step i + 1 before it reads the scalars of step i. Every
    step's result must be the one the synchronous call returns, in order; at most two may be outstanding; popping with
    nothing queued is an error, not a hang."""
    B, T = 4, 6
    cfg, params, _ = make_case(SMALL, B, T, seed=13, prior="GMM", use_c_v=True)
    batches = [O.synthetic_batch(cfg, B, T, seed=40 + i, dtype=torch.float64, ragged=True) for i in range(5)]
    feeds = [feed_of(b) for b in batches]
    sync = engine_for(cfg, params, B, T)
    want = [sync.train_step(anneal=i, rng={"seed": 9 + i}, **feeds[i]) for i in range(5)]
    lazy = engine_for(cfg, params, B, T)
    with pytest.raises(Exception):
        lazy.pop_result()  # nothing queued (and no step has run)
    pinned = [{k: (lazy.pinned(v) if v is not None else None) for k, v in f.items()} for f in feeds]
    got = []
    lazy.stage_batch(0, **pinned[0])
    for i in range(5):
        if i + 1 < 5:
            lazy.stage_batch((i + 1) & 1, **pinned[i + 1])
        assert lazy.train_step_staged(i & 1, anneal=i, rng={"seed": 9 + i}, fetch=False) is None
        lazy.queue_result()
        if i >= 1:
            got.append(lazy.pop_result())  # the result of step i - 1, while step i is in flight
    # two outstanding are allowed, a third is refused
    lazy.queue_result()
    with pytest.raises(Exception):
        lazy.queue_result()
    got.append(lazy.pop_result())
    dup = lazy.pop_result()  # the second queued copy of the last step
    with pytest.raises(Exception):
        lazy.pop_result()
    assert len(got) == 5
    for a, b in zip(got + [dup], want + [want[-1]]):
        for k in ("rec_loss", "kld", "lower_bound", "global_norm", "n_tokens", "annealing"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k])), (k, a[k], b[k])
    sync.close()
    lazy.close()
