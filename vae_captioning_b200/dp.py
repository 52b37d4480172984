"""Data-parallel train step: one process per GPU, minibatch sharded by image, ONE sum all-reduce of the flat
gradient buffer, identical clip + Adam on every rank.

The reference is single-device (`main.py --gpu 0`); the step shards naturally on the batch because images and their
captions are independent units. Contract (stated in DESIGN.md): a W-way step equals ONE reference process whose
gradient is the average of W towers, each tower seeing B/W images -- the z reshape of decoder.py:109-110 mixes rows of
the *local* batch (Q1), so a tower is a reference replica at batch B/W, not a slice of one big batch. The Q4 global
norm treats the embedding gradients as the concatenation of every tower's IndexedSlices (their squared norms travel
in the 64-float tail of the gradient buffer and are summed by the same all-reduce).

`DataParallelStep` is transport-agnostic: `backend` is anything with forward_backward(feed), grad_tensor() and
apply(scale) -- the libvaecap Engine on a GPU (NCCL), or a CPU stand-in in the gloo tests.
"""
import os


def world_from_env():
    """(rank, local_rank, world_size) as torchrun exports them; (0, 0, 1) outside torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_images, rank, world):
    """Images [lo, hi) owned by `rank`: contiguous blocks, the first n_images % world ranks take one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_PER_IMAGE = ("image_f_inputs",)
_PER_CAPTION = ("ann_inputs_enc", "ann_inputs_dec", "ann_lengths", "c_i")


def shard_feed(feed, rank, world, num_captions):
    """The slice of a reference feed dict (main.py:229-236) owned by `rank`. Caption rows follow their image:
    row n = b * num_captions + c (utils/caption_utils.py:16-21). Works on numpy arrays and torch tensors."""
    B = feed["image_f_inputs"].shape[0]
    lo, hi = shard_range(B, rank, world)
    if hi == lo:
        raise ValueError("rank %d would get an empty shard (batch %d over %d ranks)" % (rank, B, world))
    out = {}
    for k, v in feed.items():
        if v is None:
            out[k] = None
        elif k in _PER_IMAGE:
            out[k] = v[lo:hi]
        elif k in _PER_CAPTION:
            if v.shape[0] != B * num_captions:
                raise ValueError("%s has %d rows, expected %d" % (k, v.shape[0], B * num_captions))
            out[k] = v[lo * num_captions:hi * num_captions]
        else:
            out[k] = v
    return out


class DataParallelStep(object):
    """forward+backward on the local shard -> sum all-reduce of the flat gradient -> clip + Adam with scale 1/W."""

    def __init__(self, backend, group=None):
        import torch.distributed as dist
        self.backend = backend
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self, local_feed, fetch=True):
        self.backend.forward_backward(local_feed)
        # an Engine with its own communicator (Engine.attach_comm) has already summed the buckets inside forward_backward
        if self.world > 1 and not getattr(self.backend, "reduces_internally", False):
            self.dist.all_reduce(self.backend.grad_tensor(), op=self.dist.ReduceOp.SUM, group=self.group)
        return self.backend.apply(1.0 / self.world, fetch)


class EngineBackend(object):
    """Adapter of vae_captioning_b200.engine.Engine (device-resident feed: torch CUDA tensors) for DataParallelStep."""

    def __init__(self, engine, device_index, seed=0):
        from . import lib as L
        import torch
        self.engine = engine
        self.reduces_internally = bool(getattr(engine, "reduces_internally", False))
        self.seed = seed + world_from_env()[0]  # every tower draws its own noise
        self.step = 0
        ptr, count = engine.grad_buffer()
        self._grad = L.alias_tensor(ptr, count, torch.float32, device_index)

    def forward_backward(self, feed):
        rng = feed.get("rng") or {"seed": self.seed}
        self.engine.forward_backward_device(feed["image_f_inputs"], feed["ann_inputs_enc"], feed["ann_inputs_dec"],
                                            feed["ann_lengths"], feed.get("anneal", self.step), c_i=feed.get("c_i"), rng=rng)
        self.step += 1

    def grad_tensor(self):
        return self._grad

    def apply(self, scale, fetch=True):
        return self.engine.apply_gradients(scale, fetch=fetch)
