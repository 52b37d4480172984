"""Seeded synthetic inputs and weights for benchmarks and smoke runs.

There is no network for MSCOCO or the ImageNet VGG16 weights, so benchmarks feed data with the
shapes, dtypes and value ranges the reference's batch generator yields (utils/batch_gen.py:296-345,
utils/image_utils.py:5-13) and glorot-uniform weights as TF would initialise them.
"""
import math

import numpy as np

PAD, BOS, EOS = 0, 1, 2
# object-category columns that never occur in obj_vectors/c_v.pickle (91-column numbering)
_UNUSED_CV = {0, 12, 26, 29, 30, 45, 66, 68, 69, 71, 83}


def init_weights(variables, seed=1):
    """{name: fp32 array}: glorot-uniform kernels (He-uniform for the CNN so 13 random ReLU layers stay alive),
    zero biases -- what tf.global_variables_initializer gives the reference graph."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, shape, _ in variables:
        if len(shape) == 1:
            out[name] = np.zeros(shape, np.float32)
            continue
        if len(shape) == 4:
            fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
        else:
            fan_in, fan_out = shape[0], shape[1]
        lim = math.sqrt(6.0 / fan_in) if name.startswith("cnn/") else math.sqrt(6.0 / (fan_in + fan_out))
        out[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
    return out


def make_batch(B, C, T, V, seed=0, feature_size=4096, images=False, cluster_vectors=False, ragged=False):
    """One feed dict worth of host arrays. Row n = b*C + c (utils/caption_utils.py:16-21).

    Every caption has T+1 tokens: cap_in = <BOS> w1..w_{T-1}, cap_lbl = w1..w_{T-1} <EOS>, length T
    (so every position carries loss), unless ragged."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = B * C
    feed = {}
    if images:
        feed["image_f_inputs"] = rng.integers(0, 256, size=(B, 224, 224, 3), dtype=np.uint8).astype(np.float32)
    else:
        feed["image_f_inputs"] = np.maximum(0, rng.standard_normal((B, feature_size), dtype=np.float32))
    body = (rng.zipf(1.1, size=(N, T)) % max(V - 3, 1) + 3).astype(np.int32)
    lengths = np.full((N,), T, np.int32)
    if ragged:
        lengths = rng.integers(0, T + 1, size=(N,)).astype(np.int32)
    cap_in = np.zeros((N, T), np.int32)
    cap_lbl = np.zeros((N, T), np.int32)
    for n in range(N):
        L = int(lengths[n])
        if L == 0:
            continue
        seq = [BOS] + list(body[n, :L - 1]) + [EOS]
        cap_in[n, :L] = seq[:L]
        cap_lbl[n, :L] = seq[1:L + 1]
    feed["ann_inputs_enc"] = cap_lbl  # the encoder consumes the label sequence (main.py:230, Q8)
    feed["ann_inputs_dec"] = cap_in
    feed["ann_lengths"] = lengths
    if cluster_vectors:
        used = [k for k in range(91) if k not in _UNUSED_CV]
        cv = np.zeros((B, 91), np.float32)
        for b in range(B):
            k = int(np.clip(rng.geometric(0.34), 1, 18))
            cv[b, rng.choice(used, size=k, replace=False)] = 1.0 / k
        feed["c_i"] = np.repeat(cv[:, None, :], C, axis=1).reshape(N, 91)[:, 1:].copy()  # main.py:236 drops column 0
    return feed


def init_clusters(num_clusters, latent_size, seed=2, c_m_file=None):
    """Cluster means of the AG prior (utils/vae_utils.py:6-31): num_clusters random vectors in [-1, 1]^Z scaled to unit
    norm, persisted in (and re-read from) `c_m_file` (the reference's ./pickles/cluster_means.pickle) when given."""
    import os
    import pickle
    if c_m_file and os.path.exists(c_m_file):
        with open(c_m_file, "rb") as rf:
            return np.squeeze(np.asarray(pickle.load(rf), dtype=np.float32))
    rng = np.random.Generator(np.random.PCG64(seed))
    m = 2 * rng.random((num_clusters, latent_size)) - 1
    m = (m / np.sqrt(np.sum(m ** 2, axis=1, keepdims=True))).astype(np.float32)
    if c_m_file:
        d = os.path.dirname(c_m_file)
        if d and not os.path.exists(d):
            os.makedirs(d)
        with open(c_m_file, "wb") as wf:
            pickle.dump(m[:, None, :], wf)  # the reference stacks [1, Z] rows: shape [K, 1, Z]
    return m
