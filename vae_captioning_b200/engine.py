"""Host-side mirror of the reference's per-step `sess.run` surface (main.py:229-244, 262-284) on
top of the libvaecap C ABI (include/vaecap.h).

`Engine` plays the role of the TF graph + session the reference builds in main.py:43-196: it is
constructed from a `Parameters` object (+ vocab size), owns the variables by TF name, and exposes
train / eval / debug-tap calls whose arguments are the reference's feed-dict entries. Arrays
crossing the boundary are numpy (host) or torch CUDA tensors (device containers only).
"""
import ctypes

import numpy as np

from . import lib as L

PRIORS = {"Normal": 0, "GMM": 1, "AG": 2}
OPTIMIZERS = {"Adam": 0, "SGD": 1, "Momentum": 2}


class VcConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "vocab_size", "embed_size", "encoder_hidden", "decoder_hidden", "latent_size", "gen_z_samples", "num_clusters",
        "num_captions", "cnn_feature_size", "prior", "use_c_v", "no_encoder", "fine_tune", "restore", "with_cnn",
        "max_batch", "max_len", "optimizer", "cnn_optimizer", "lr_decay_steps")] + [(n, ctypes.c_float) for n in (
            "dec_keep_rate", "dec_lstm_drop", "cnn_dropout", "weight_decay", "learning_rate", "cnn_lr", "clip_norm",
            "ann_param", "std", "temperature")]


class VcRng(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("eps_dev", ctypes.c_void_p), ("emb_keep_dev", ctypes.c_void_p),
                ("out_keep_dev", ctypes.c_void_p), ("gmm_cluster_dev", ctypes.c_void_p),
                ("cnn_keep_dev", ctypes.c_void_p)]


class VcStepOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("kld", "rec_loss", "lower_bound", "annealing", "global_norm", "n_tokens")]

    def as_dict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


def config_from_params(params, vocab_size=None, max_batch=None, max_len=30, with_cnn=None):
    """vc_config from a reference-style Parameters object (utils/parameters.py:3-66)."""
    c = VcConfig()
    c.vocab_size = int(vocab_size if vocab_size is not None else params.vocab_size)
    c.embed_size = int(params.embed_size)
    c.encoder_hidden = int(params.encoder_hidden)
    c.decoder_hidden = int(params.decoder_hidden)
    c.latent_size = int(params.latent_size)
    c.gen_z_samples = int(params.gen_z_samples)
    c.num_clusters = int(getattr(params, "num_clusters", 90))
    # main.py:84-89: features are tiled x num_captions only in training mode
    c.num_captions = int(params.num_captions)
    c.cnn_feature_size = int(getattr(params, "cnn_feature_size", 4096))
    if params.prior not in PRIORS:
        raise ValueError("unknown prior %r (expected one of Normal, GMM, AG)" % (params.prior,))
    c.prior = PRIORS[params.prior]
    c.use_c_v = int(bool(params.use_c_v))
    c.no_encoder = int(bool(params.no_encoder))
    c.fine_tune = int(bool(params.fine_tune))
    c.restore = int(bool(params.restore))
    c.with_cnn = int(bool(params.fine_tune if with_cnn is None else with_cnn))
    c.max_batch = int(max_batch if max_batch is not None else params.batch_size)
    c.max_len = int(max_len)
    for field, attr in (("optimizer", "optimizer"), ("cnn_optimizer", "cnn_optimizer")):
        kind = getattr(params, attr, "Adam")
        if kind not in OPTIMIZERS:  # argparse restricts --optimizer to these three (utils/parameters.py:104-106)
            raise ValueError("unknown %s %r (expected one of SGD, Adam, Momentum)" % (attr, kind))
        setattr(c, field, OPTIMIZERS[kind])
    # ops/optimizers.py:24-26
    c.lr_decay_steps = int(getattr(params, "num_ex_per_epoch", 150000) / (getattr(params, "batch_size", c.max_batch) + 0.001) *
                           getattr(params, "num_epochs_per_decay", 5))
    c.dec_keep_rate = float(params.dec_keep_rate)
    c.dec_lstm_drop = float(params.dec_lstm_drop)
    c.cnn_dropout = float(getattr(params, "cnn_dropout", 0.5))
    c.weight_decay = float(getattr(params, "weight_decay", 0.00004))
    c.learning_rate = float(params.learning_rate)
    c.cnn_lr = float(getattr(params, "cnn_lr", 0.00001))
    c.clip_norm = float(getattr(params, "lstm_clip_by_norm", 5.0))
    c.ann_param = float(params.ann_param)
    c.std = float(params.std)
    c.temperature = float(params.temperature)
    return c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)


class Engine(object):
    """One handle per GPU per process. Not re-entrant (like a tf.Session used from one thread)."""

    def __init__(self, params, vocab_size=None, max_batch=None, max_len=30, device=0, with_cnn=None):
        self.lib = L.load()
        self._declare()
        self.cfg = config_from_params(params, vocab_size, max_batch, max_len, with_cnn)
        self.device = device
        self._h = ctypes.c_void_p(0)
        L.check(self.lib.vc_create(ctypes.byref(self.cfg), device, ctypes.byref(self._h)))
        self._keep = []
        self._info = None

    def _declare(self):
        lib = self.lib
        if getattr(lib, "_vc_declared", False):
            return
        vp, ci, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        lib.vc_create.argtypes = [ctypes.POINTER(VcConfig), ci, ctypes.POINTER(vp)]
        lib.vc_destroy.argtypes = [vp]
        lib.vc_num_params.argtypes = [vp]
        lib.vc_param_info.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)]
        for f in (lib.vc_param_get, lib.vc_param_set, lib.vc_grad_get):
            f.argtypes = [vp, ctypes.c_char_p, vp]
        step_args = [vp, vp, vp, vp, vp, vp, ci, ci, cll, ctypes.POINTER(VcRng)]
        lib.vc_train_step.argtypes = step_args + [ctypes.POINTER(VcStepOut), vp]
        lib.vc_train_step_dev.argtypes = step_args + [ctypes.POINTER(VcStepOut), vp]
        lib.vc_train_step_images.argtypes = step_args + [ctypes.POINTER(VcStepOut), vp]
        lib.vc_train_step_images_u8.argtypes = step_args + [ctypes.POINTER(VcStepOut), vp]
        lib.vc_vgg_forward_u8.argtypes = [vp, vp, vp, ci, vp]
        lib.vc_stage_batch.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, ci, ci, vp]
        lib.vc_train_step_staged.argtypes = [vp, ci, cll, ctypes.POINTER(VcRng), ctypes.POINTER(VcStepOut), vp]
        lib.vc_forward_backward_staged.argtypes = [vp, ci, cll, ctypes.POINTER(VcRng), vp]
        lib.vc_forward_backward_dev.argtypes = step_args + [vp]
        lib.vc_eval_step.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ctypes.POINTER(VcRng), ctypes.POINTER(VcStepOut), vp]
        lib.vc_set_cluster_means.argtypes = [vp, vp]
        lib.vc_grad_buffer.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int64)]
        lib.vc_apply_gradients.argtypes = [vp, ctypes.c_float, ctypes.POINTER(VcStepOut), vp]
        lib.vc_forward_debug.argtypes = [vp] * 7
        lib.vc_vgg_forward.argtypes = [vp, vp, vp, ci, vp]
        lib.vc_vgg_forward_dev.argtypes = [vp, vp, vp, ci, vp]
        lib.vc_vgg_activation.argtypes = [vp, ctypes.c_char_p, vp]
        lib.vc_vgg_keep_activations.argtypes = [vp, ci]
        lib.vc_comm_unique_id.argtypes = [vp]
        lib.vc_comm_init.argtypes = [vp, vp, ci, ci]
        lib.vc_allreduce_gradients.argtypes = [vp, vp]
        lib.vc_comm_set_mode.argtypes = [vp, ci]
        lib.vc_comm_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_longlong),
                                      ctypes.POINTER(ctypes.c_int)]
        lib._vc_declared = True

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self._h:
            self.lib.vc_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ data parallelism
    def attach_comm(self, rank=None, world=None, group=None):
        """vc_comm_init: gives this handle its NCCL communicator. Rank 0 draws the unique id, torch.distributed (any
        backend; only used as the bootstrap channel) broadcasts it. From then on every train_step* call is the
        data-parallel step: bucketed gradient all-reduce overlapped with the backward pass, mean of the towers."""
        import torch.distributed as dist
        rank = dist.get_rank(group) if rank is None else rank
        world = dist.get_world_size(group) if world is None else world
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            L.check(self.lib.vc_comm_unique_id(buf))
        box = [bytes(buf.raw) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        L.check(self.lib.vc_comm_init(self._h, box[0], int(rank), int(world)))
        self.world = int(world)
        return self

    @property
    def reduces_internally(self):
        return getattr(self, "world", 1) > 1

    def comm_set_mode(self, mode):
        """1 bucketed + overlapped (default), 2 one all-reduce behind the backward pass, 0 none (timing only)."""
        L.check(self.lib.vc_comm_set_mode(self._h, int(mode)))

    def comm_stats(self):
        ms, nb, calls = ctypes.c_float(), ctypes.c_longlong(), ctypes.c_int()
        L.check(self.lib.vc_comm_stats(self._h, ctypes.byref(ms), ctypes.byref(nb), ctypes.byref(calls)))
        return {"span_ms": float(ms.value), "bytes": int(nb.value), "buckets": int(calls.value)}

    def allreduce_gradients(self):
        L.check(self.lib.vc_allreduce_gradients(self._h, self._stream()))

    # ------------------------------------------------------------------ variables (tf.train.Saver surface)
    def variables(self):
        """[(name, shape, trainable)] in creation order -- tf.trainable_variables() of the reference graph."""
        if self._info is None:
            out = []
            n = self.lib.vc_num_params(self._h)
            for i in range(n):
                name = ctypes.c_char_p()
                ndim = ctypes.c_int32()
                shape = (ctypes.c_int64 * 4)()
                tr = ctypes.c_int32()
                L.check(self.lib.vc_param_info(self._h, i, ctypes.byref(name), ctypes.byref(ndim), shape, ctypes.byref(tr)))
                out.append((name.value.decode(), tuple(int(shape[j]) for j in range(ndim.value)), bool(tr.value)))
            self._info = out
        return self._info

    def _shape(self, name):
        for n, s, _ in self.variables():
            if n == name:
                return s
        raise ValueError("unknown variable %r" % name)

    def set_variable(self, name, value):
        shape = self._shape(name)
        a = _f32(value)
        if tuple(a.shape) != shape:
            raise ValueError("variable %s expects shape %s, got %s" % (name, shape, tuple(a.shape)))
        L.check(self.lib.vc_param_set(self._h, name.encode(), _np_ptr(a)))

    def get_variable(self, name):
        a = np.empty(self._shape(name), dtype=np.float32)
        L.check(self.lib.vc_param_get(self._h, name.encode(), _np_ptr(a)))
        return a

    def get_gradient(self, name):
        a = np.empty(self._shape(name), dtype=np.float32)
        L.check(self.lib.vc_grad_get(self._h, name.encode(), _np_ptr(a)))
        return a

    def set_cluster_means(self, c_means):
        """c_means [num_clusters, latent_size]: init_clusters() of the reference (utils/vae_utils.py:6-31), AG prior."""
        a = _f32(c_means)
        if a.shape != (self.cfg.num_clusters, self.cfg.latent_size):
            raise ValueError("cluster means must be [%d, %d], got %s" % (self.cfg.num_clusters, self.cfg.latent_size, a.shape))
        L.check(self.lib.vc_set_cluster_means(self._h, _np_ptr(a)))

    def load_state(self, state):
        """state: {tf_name: array}. Unknown names are rejected, missing names keep their value."""
        for name, value in state.items():
            self.set_variable(name, value.detach().cpu().numpy() if hasattr(value, "detach") else value)

    def state(self):
        return {n: self.get_variable(n) for n, _, _ in self.variables()}

    # ------------------------------------------------------------------ randomness
    def _rng(self, rng):
        """rng: None or dict(seed=, eps=, emb_keep=, out_keep=, gmm_cluster=, cnn_keep=) of torch CUDA tensors."""
        r = VcRng()
        keep = []
        if rng:
            r.seed = int(rng.get("seed", 0))
            for key, field in (("eps", "eps_dev"), ("emb_keep", "emb_keep_dev"), ("out_keep", "out_keep_dev"),
                               ("gmm_cluster", "gmm_cluster_dev"), ("cnn_keep", "cnn_keep_dev")):
                t = rng.get(key)
                if t is not None:
                    if not t.is_cuda or not t.is_contiguous():
                        raise ValueError("rng[%s] must be a contiguous CUDA tensor" % key)
                    keep.append(t)
                    setattr(r, field, t.data_ptr())
        return r, keep

    # ------------------------------------------------------------------ the sess.run calls
    def _stream(self):
        return L.stream_ptr()

    def train_step(self, image_f_inputs, ann_inputs_enc, ann_inputs_dec, ann_lengths, anneal, c_i=None, rng=None,
                   fetch=True, images=False):
        """sess.run([kld, rec_loss, lower_bound, optimize, optimize_cnn, annealing], feed) (main.py:229-244).

        Host (numpy) inputs: copied to the device inside the call. Returns dict(kld, rec_loss, lower_bound,
        annealing, global_norm, n_tokens), or None when fetch=False (no synchronisation)."""
        # uint8 pixel arrays (the HDF5 image store's dtype, utils/batch_gen.py:278-294) are fed as they are
        u8 = getattr(image_f_inputs, "dtype", None) == np.uint8 and (images or self.cfg.fine_tune)
        feats = np.ascontiguousarray(image_f_inputs) if u8 else _f32(image_f_inputs)
        lbl, inp = _i32(ann_inputs_enc), _i32(ann_inputs_dec)
        ln = _i32(np.asarray(ann_lengths).ravel())
        cv = _f32(c_i) if c_i is not None else None
        B, T = feats.shape[0], lbl.shape[1]
        self._check_feed(feats, lbl, inp, ln, cv, B, T, images)
        r, keep = self._rng(rng)
        out = VcStepOut()
        self._keep = [feats, lbl, inp, ln, cv, keep]
        fn = self.lib.vc_train_step_images_u8 if u8 else (self.lib.vc_train_step_images if images else self.lib.vc_train_step)
        L.check(fn(self._h, _np_ptr(feats), _np_ptr(lbl), _np_ptr(inp), _np_ptr(ln), _np_ptr(cv), B, T,
                                       int(anneal), ctypes.byref(r), ctypes.byref(out) if fetch else None, self._stream()))
        return out.as_dict() if fetch else None

    # ------------------------------------------------------------------ double-buffered feed
    def stage_batch(self, slot, image_f_inputs, ann_inputs_enc, ann_inputs_dec, ann_lengths, c_i=None, images=False):
        """Starts the H2D copy of one step's feed into staging slot 0/1 on the engine's copy stream and returns at
        once (asynchronous for pinned host arrays, see `pinned`). Run it with `train_step_staged(slot, ...)`; staging
        batch i+1 before launching step i overlaps its copy with step i's compute."""
        import torch
        u8 = getattr(image_f_inputs, "dtype", None) == np.uint8 and (images or self.cfg.fine_tune)
        feats = np.ascontiguousarray(image_f_inputs) if u8 else _f32(image_f_inputs)
        lbl, inp = _i32(ann_inputs_enc), _i32(ann_inputs_dec)
        ln = _i32(np.asarray(ann_lengths).ravel())
        cv = _f32(c_i) if c_i is not None else None
        B, T = feats.shape[0], lbl.shape[1]
        self._check_feed(feats, lbl, inp, ln, cv, B, T, images)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staged = {}
        kind = 2 if u8 else (1 if images and not self.cfg.fine_tune else 0)
        self._staged[slot] = (feats, lbl, inp, ln, cv)  # the host arrays must outlive the asynchronous copy
        L.check(self.lib.vc_stage_batch(self._h, int(slot), _np_ptr(feats), kind, _np_ptr(lbl), _np_ptr(inp), _np_ptr(ln),
                                        _np_ptr(cv), B, T, ctypes.c_void_p(self._copy_stream.cuda_stream)))

    def train_step_staged(self, slot, anneal, rng=None, fetch=True):
        """The train step (main.py:229-244) on the batch staged in `slot`. Same return value as train_step."""
        r, keep = self._rng(rng)
        out = VcStepOut()
        self._keep = [keep]
        L.check(self.lib.vc_train_step_staged(self._h, int(slot), int(anneal), ctypes.byref(r),
                                              ctypes.byref(out) if fetch else None, self._stream()))
        return out.as_dict() if fetch else None

    def queue_result(self):
        """Queue the read-back of the last step's scalars (a step called with fetch=False) without waiting for it."""
        L.check(self.lib.vc_step_result_queue(self._h, self._stream()))

    def pop_result(self):
        """Wait for the OLDEST queued read-back; same dict as train_step returns. Lets the host enqueue step i + 1 before
        it reads the loss of step i (at most two results outstanding)."""
        out = VcStepOut()
        L.check(self.lib.vc_step_result(self._h, ctypes.byref(out)))
        return out.as_dict()

    def forward_backward_staged(self, slot, anneal, rng=None):
        """Data-parallel half of train_step_staged: gradients stay in grad_buffer() for the all-reduce; finish with
        apply_gradients(1 / world)."""
        r, keep = self._rng(rng)
        self._keep = [keep]
        L.check(self.lib.vc_forward_backward_staged(self._h, int(slot), int(anneal), ctypes.byref(r), self._stream()))

    @staticmethod
    def pinned(array):
        """A page-locked copy of a numpy array (torch as the allocator), for truly asynchronous stage_batch copies."""
        import torch
        return torch.from_numpy(np.ascontiguousarray(array)).pin_memory().numpy()

    def train_step_device(self, feats, cap_lbl, cap_in, lengths, anneal, c_i=None, rng=None, fetch=True):
        """Same step with inputs already resident on the device (torch CUDA tensors as containers)."""
        B, T = feats.shape[0], cap_lbl.shape[1]
        r, keep = self._rng(rng)
        out = VcStepOut()
        self._keep = [feats, cap_lbl, cap_in, lengths, c_i, keep]
        L.check(self.lib.vc_train_step_dev(self._h, L.ptr(feats), L.ptr(cap_lbl), L.ptr(cap_in), L.ptr(lengths), L.ptr(c_i),
                                           B, T, int(anneal), ctypes.byref(r), ctypes.byref(out) if fetch else None,
                                           self._stream()))
        return out.as_dict() if fetch else None

    def forward_backward_device(self, feats, cap_lbl, cap_in, lengths, anneal, c_i=None, rng=None):
        B, T = feats.shape[0], cap_lbl.shape[1]
        r, keep = self._rng(rng)
        self._keep = [feats, cap_lbl, cap_in, lengths, c_i, keep]
        L.check(self.lib.vc_forward_backward_dev(self._h, L.ptr(feats), L.ptr(cap_lbl), L.ptr(cap_in), L.ptr(lengths),
                                                 L.ptr(c_i), B, T, int(anneal), ctypes.byref(r), self._stream()))

    def grad_buffer(self):
        """(device pointer, element count) of the flat fp32 gradient buffer the data-parallel all-reduce sums."""
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        L.check(self.lib.vc_grad_buffer(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def apply_gradients(self, grad_scale=1.0, fetch=True):
        out = VcStepOut()
        L.check(self.lib.vc_apply_gradients(self._h, float(grad_scale), ctypes.byref(out) if fetch else None, self._stream()))
        return out.as_dict() if fetch else None

    def eval_step(self, image_f_inputs, ann_inputs_enc, ann_inputs_dec, ann_lengths, c_i=None, rng=None):
        """sess.run([rec_loss], feed) of validate() (main.py:262-284): forward only, training graph (Q12)."""
        feats = _f32(image_f_inputs)
        lbl, inp = _i32(ann_inputs_enc), _i32(ann_inputs_dec)
        ln = _i32(np.asarray(ann_lengths).ravel())
        cv = _f32(c_i) if c_i is not None else None
        B, T = feats.shape[0], lbl.shape[1]
        self._check_feed(feats, lbl, inp, ln, cv, B, T)
        r, keep = self._rng(rng)
        out = VcStepOut()
        self._keep = [feats, lbl, inp, ln, cv, keep]
        L.check(self.lib.vc_eval_step(self._h, _np_ptr(feats), _np_ptr(lbl), _np_ptr(inp), _np_ptr(ln), _np_ptr(cv), B, T,
                                      ctypes.byref(r), ctypes.byref(out), self._stream()))
        return out.as_dict()

    def _check_feed(self, feats, lbl, inp, ln, cv, B, T, images=False):
        N = B * self.cfg.num_captions
        if lbl.shape != (N, T) or inp.shape != (N, T):
            raise ValueError("captions must be [B*num_captions=%d, T] (got %s and %s)" % (N, lbl.shape, inp.shape))
        if ln.shape != (N,):
            raise ValueError("ann_lengths must have %d entries, got %s" % (N, ln.shape))
        if cv is not None and cv.shape != (N, self.cfg.num_clusters):
            raise ValueError("c_i must be [%d, %d], got %s" % (N, self.cfg.num_clusters, cv.shape))
        per_image = 224 * 224 * 3 if (self.cfg.fine_tune or images) else self.cfg.cnn_feature_size
        if feats.size != B * per_image:
            raise ValueError("image_f_inputs has %d values per row, expected %d" % (feats.size // max(B, 1), per_image))

    # ------------------------------------------------------------------ VGG16 feature extractor
    def vgg_forward(self, images):
        """sess.run(features, {input_img: images}) of Data.extract_features_from_dir (utils/data.py:120-125), batched:
        images [B,224,224,3] RGB 0..255 (host) -> fc2 features fp32 [B,4096] (host)."""
        u8 = getattr(images, "dtype", None) == np.uint8
        img = np.ascontiguousarray(images) if u8 else _f32(images)
        if img.ndim != 4 or img.shape[1:] != (224, 224, 3):
            raise ValueError("images must be [B,224,224,3], got %s" % (img.shape,))
        out = np.empty((img.shape[0], 4096), np.float32)
        fn = self.lib.vc_vgg_forward_u8 if u8 else self.lib.vc_vgg_forward
        L.check(fn(self._h, _np_ptr(img), _np_ptr(out), img.shape[0], self._stream()))
        return out

    def vgg_forward_device(self, images, out=None):
        """Device-resident form: images is a CUDA float32 tensor [B,224,224,3]; returns a CUDA tensor [B,4096]
        (`out` when given). Runs on torch's current stream."""
        import torch
        B = images.shape[0]
        if out is None:
            out = torch.empty((B, 4096), dtype=torch.float32, device=images.device)
        self._keep_vgg = (images, out)
        L.check(self.lib.vc_vgg_forward_dev(self._h, L.ptr(images), L.ptr(out), B, self._stream()))
        return out

    def vgg_keep_activations(self, on=True):
        L.check(self.lib.vc_vgg_keep_activations(self._h, int(bool(on))))

    def vgg_activation(self, layer, B):
        shapes = {"conv1_1": (224, 64), "conv1_2": (224, 64), "pool1": (112, 64), "conv2_1": (112, 128),
                  "conv2_2": (112, 128), "pool2": (56, 128), "conv3_1": (56, 256), "conv3_2": (56, 256),
                  "conv3_3": (56, 256), "pool3": (28, 256), "conv4_1": (28, 512), "conv4_2": (28, 512),
                  "conv4_3": (28, 512), "pool4": (14, 512), "conv5_1": (14, 512), "conv5_2": (14, 512),
                  "conv5_3": (14, 512), "pool5": (7, 512)}
        if layer in ("fc1", "fc2"):  # post-ReLU (+ dropout when fine-tuning) [B, 4096]
            a = np.empty((B, 4096), np.float32)
            L.check(self.lib.vc_vgg_activation(self._h, layer.encode(), _np_ptr(a)))
            return a
        if layer not in shapes:
            raise ValueError("unknown VGG layer %r" % layer)
        hw, c = shapes[layer]
        a = np.empty((B, hw, hw, c), np.float32)
        L.check(self.lib.vc_vgg_activation(self._h, layer.encode(), _np_ptr(a)))
        return a

    def debug_taps(self, N, T, logits=True, z=False):
        """x_logits (main.py:150), qz mean/std (main.py:122-124), per-row KL and CE of the last forward-only pass."""
        V, Z, S = self.cfg.vocab_size, self.cfg.latent_size, self.cfg.gen_z_samples
        out = {}
        lg = np.empty((N * T, V), np.float32) if logits else None
        ce = np.empty((N, T), np.float32)
        if self.cfg.no_encoder:
            L.check(self.lib.vc_forward_debug(self._h, _np_ptr(lg), None, None, None, None, _np_ptr(ce)))
        else:
            mu = np.empty((N, Z), np.float32)
            sd = np.empty((N, Z), np.float32)
            kl = np.empty((N,), np.float32)
            zz = np.empty((S, N, Z), np.float32) if z else None
            L.check(self.lib.vc_forward_debug(self._h, _np_ptr(lg), _np_ptr(mu), _np_ptr(sd), _np_ptr(zz), _np_ptr(kl),
                                              _np_ptr(ce)))
            out.update(mu=mu, std=sd, kl_rows=kl)
            if z:
                out["z"] = zz
        if logits:
            out["logits"] = lg
        out["ce"] = ce
        return out
