"""`main.py` of the reference (main.py:14-300) on the B200 path: same flags (`Parameters.parse_args`), same loop
structure, prints and file outputs (`./checkpoints/{checkpoint}.ckpt*`, `./val_{gen_name}.json`, `./test_{gen_name}.json`),
with the TF graph + session replaced by `Engine` / `Decoder`.

The data side mirrors utils/data.py / utils/batch_gen.py (`data.Data`, `batch_gen.Batch_Generator`) and is
used when `params.coco_dir` exists; `feeder` is anything with the reference generators' interface. Without a COCO
tree, a seeded synthetic feeder with the reference's shapes and dtypes is used (there is no COCO on the bench box):

    python -m vae_captioning_b200.main --gpu 0 --bs 32 --epochs 1 [--prior AG --c_v] [--mode inference]
"""
import os
import sys

import numpy as np

from . import checkpoint, synthetic
from .caption_utils import preprocess_captions
from .parameters import Parameters


class SyntheticFeeder(object):
    """Stands in for Data.load_train_data_generator / get_valid_data / get_test_data (utils/data.py, batch_gen.py:164-345):
    yields (f_images_batch, (inputs, labels), lengths, c_v) with captions as [B, C, T] like the reference's generator."""

    def __init__(self, params, vocab_size, batches, seed=0, T=20):
        self.params, self.V, self.batches, self.seed, self.T = params, vocab_size, batches, seed, T

    def _batch(self, i):
        p = self.params
        f = synthetic.make_batch(p.batch_size, p.num_captions, self.T, self.V, seed=self.seed + i, images=p.fine_tune,
                                 cluster_vectors=True, ragged=True)
        B, C, T = p.batch_size, p.num_captions, self.T
        # one 91-wide cluster vector per IMAGE, like the generator's cl_v (batch_gen.py:101-130); preprocess_captions tiles it
        c_v = np.concatenate([np.zeros((B * C, 1), np.float32), f["c_i"]], axis=1).reshape(B, C, 91)[:, 0, :]
        lengths = f["ann_lengths"].astype(np.float64).reshape(B, C)  # float64 like batch_gen.py:317; 0 = missing caption
        return f["image_f_inputs"], (f["ann_inputs_dec"].reshape(B, C, T), f["ann_inputs_enc"].reshape(B, C, T)), lengths, c_v

    def next_batch(self, use_obj_vectors=False, num_captions=1):
        for i in range(self.batches):
            yield self._batch(i)

    def next_val_batch(self, get_image_ids=False, use_obj_vectors=False):
        for i in range(self.batches):
            feats, caps, lens, c_v = self._batch(1000 + i)
            ids = list(range(i * len(feats), (i + 1) * len(feats)))
            yield feats, caps, lens, ids, c_v

    def next_test_batch(self, use_obj_vectors=False):
        for i in range(self.batches):
            feats, _, _, c_v = self._batch(2000 + i)
            yield feats, list(range(i * len(feats), (i + 1) * len(feats))), c_v


class _Vocabulary(object):
    def __init__(self, vocab_size):
        self.idx2word = {i: "w%d" % i for i in range(vocab_size)}
        self.idx2word.update({0: "<PAD>", 1: "<BOS>", 2: "<EOS>"})
        self.word2idx = {w: i for i, w in self.idx2word.items()}
        self.vocab_size = vocab_size


CLUSTER_MEANS_FILE = "./pickles/cluster_means.pickle"  # utils/vae_utils.py:7


def _feed(params, f_images_batch, captions_batch, cl_batch, c_v):
    """main.py:225-236: flatten [B, C, T] captions, pick labels/inputs, drop the background cluster column."""
    if params.num_captions > 1:
        captions_batch, cl_batch, c_v = preprocess_captions(captions_batch, cl_batch, c_v)
    feed = dict(image_f_inputs=f_images_batch, ann_inputs_enc=captions_batch[1], ann_inputs_dec=captions_batch[0],
                ann_lengths=cl_batch)
    if params.use_c_v or params.prior in ("GMM", "AG"):
        feed["c_i"] = np.asarray(c_v)[:, 1:]
    return feed


def load_data(params):
    """main.py:19-39: Data -> (train, val, test) generators + the vocabulary. Features come from the ./pickles cache or
    the device VGG16 (Data.extract_features_from_dir); with --fine_tune the generators yield uint8 images instead."""
    from .data import Data
    repartiton = not params.gen_val_captions < 0
    data = Data(params, True, params.image_net_weights_path, repartiton=repartiton, gen_val_cap=params.gen_val_captions)
    batch_gen = data.load_train_data_generator(params.batch_size, params.fine_tune)
    pretrained = not params.fine_tune
    val_gen = data.get_valid_data(params.batch_size, val_tr_unused=batch_gen.unused_cap_in, pretrained=pretrained)
    test_gen = data.get_test_data(params.batch_size, pretrained=pretrained)
    if data.engine is not None:  # the feature-extraction handle (VGG16 workspace) is not needed any more
        data.engine.close()
        data.engine = None
    return batch_gen, val_gen, test_gen, data.dictionary


def run(params, feeder=None, val_feeder=None, test_feeder=None, vocabulary=None, device=0, out=print, max_len=64,
        report_every=500, prefetch=2):
    """The body of main() (main.py:14-289). Returns the Engine (its variables are the checkpoint)."""
    from .batch_gen import Prefetcher
    from .decode import Decoder
    from .engine import Engine
    from . import inference as inference_mod
    vocab = vocabulary or _Vocabulary(params.vocab_size or 11313)
    feeder = feeder or SyntheticFeeder(params, vocab.vocab_size, batches=8)
    val_feeder = val_feeder or feeder
    test_feeder = test_feeder or feeder
    eng = Engine(params, vocab_size=vocab.vocab_size, max_batch=params.batch_size, max_len=max_len, device=device,
                 with_cnn=bool(params.fine_tune))
    ckpt = checkpoint.checkpoint_path(params)
    if params.prior == "AG" and not params.no_encoder:
        # init_clusters (utils/vae_utils.py:6-31): the reference builds c_means in BOTH modes and hands them to the
        # decoder (main.py:136-140 -> decoder.cap_clusters, decoder.py:45-71); they live in ./pickles/cluster_means.pickle
        eng.set_cluster_means(synthetic.init_clusters(params.num_clusters, params.latent_size,
                                                      c_m_file=CLUSTER_MEANS_FILE))
    if params.mode == "training":
        if not params.restore:
            eng.load_state(synthetic.init_weights(eng.variables(), seed=1))  # tf.global_variables_initializer
            if params.fine_tune and os.path.exists(params.image_net_weights_path):
                out("Loading imagenet weights for futher usage")
                eng.load_state(checkpoint.vgg16_npz_state(params.image_net_weights_path))
        else:
            out("Restoring from checkpoint")
            checkpoint.restore(eng, ckpt)
        gs = 0
        pending = 0
        lb = rl = kl = ann = float("nan")
        for e in range(params.num_epochs):
            gs_epoch = 0
            stop = False
            while not stop:
                n_batches = 0
                batches = feeder.next_batch(use_obj_vectors=params.use_c_v, num_captions=params.num_captions)
                if prefetch:  # batch i+1 is assembled on a host thread while the device runs batch i
                    batches = Prefetcher(batches, depth=prefetch)
                # double-buffered feed: batch i+1 is copied to the device (copy stream) while step i computes
                it = iter(batches)
                nxt = next(it, None)
                slot = 0
                if nxt is not None:
                    eng.stage_batch(slot, **_feed(params, *nxt))
                while nxt is not None:
                    n_batches += 1
                    cur_slot = slot
                    nxt = next(it, None)
                    if nxt is not None:
                        slot ^= 1
                        eng.stage_batch(slot, **_feed(params, *nxt))
                    # the step's scalars come back one step late (vc_step_result_queue / vc_step_result): the next step is
                    # enqueued before the host waits for this one's loss; at a report point the queue is drained first, so
                    # what is printed is the loss of the step named in the line, as in the reference (main.py:240-250)
                    eng.train_step_staged(cur_slot, anneal=gs, rng={"seed": gs}, fetch=False)
                    eng.queue_result()
                    pending += 1
                    gs += 1
                    gs_epoch += 1
                    report = gs % report_every == 0
                    while pending > (0 if report else 1):
                        res = eng.pop_result()
                        pending -= 1
                        kl, rl, lb, ann = res["kld"], res["rec_loss"], res["lower_bound"], res["annealing"]
                    if report:
                        out("Epoch: {} Iteration: {} VLB: {} Rec Loss: {}".format(e, gs, lb, rl))
                        if not params.no_encoder:
                            out("Annealing coefficient:{} KLD: {}".format(ann, kl))
                    if gs_epoch * params.batch_size > params.num_ex_per_epoch:
                        stop = True
                        break
                while pending:
                    res = eng.pop_result()
                    pending -= 1
                    kl, rl, lb, ann = res["kld"], res["rec_loss"], res["lower_bound"], res["annealing"]
                if hasattr(batches, "close"):
                    batches.close()
                if n_batches == 0 or not getattr(feeder, "endless", False):
                    stop = True
            out("Epoch: {} Iteration: {} VLB: {} Rec Loss: {}".format(e, gs, lb, rl))
            val_rec = []  # validate(): rec_loss of the training graph on the validation batches (main.py:262-284)
            for vi, (f_images_batch, captions_batch, cl_batch, c_v) in enumerate(val_feeder.next_batch(
                    use_obj_vectors=params.use_c_v, num_captions=params.num_captions)):
                # fresh noise per validation batch, as every sess.run of the reference draws its own (main.py:279)
                val_rec.append(eng.eval_step(rng={"seed": (gs << 20) + vi}, **_feed(params, f_images_batch, captions_batch, cl_batch, c_v))["rec_loss"])
            out("Validation reconstruction loss: {}".format(np.mean(val_rec) if val_rec else float("nan")))
            out("-----------------------------------------------")
            save_path = checkpoint.save(ckpt, eng.state())
            out("Model saved in file: %s" % save_path)
    if params.mode == "inference":
        checkpoint.restore(eng, ckpt)
        decoder = Decoder(eng, params, vocab)
        inference_mod.inference(params, decoder, val_feeder, test_feeder)
    return eng


def main(argv=None):
    params = Parameters()
    params.parse_args(argv)
    if params.save_params:  # main.py:295-304
        import pickle
        if not os.path.exists("./pickles"):
            os.makedirs("./pickles")
        fn = "./pickles/params_{}_{}_{}_{}.pickle".format(params.prior, params.no_encoder, params.checkpoint, params.use_c_v)
        with open(fn, "wb") as wf:
            pickle.dump(params, wf)
    if os.path.isdir(params.coco_dir):
        batch_gen, val_gen, test_gen, dictionary = load_data(params)
        batch_gen.endless = True  # the reference loops `while True` over next_batch until num_ex_per_epoch (main.py:216-246)
        run(params, feeder=batch_gen, val_feeder=val_gen, test_feeder=test_gen, vocabulary=dictionary)
    else:
        print("No COCO tree at {}: running on the seeded synthetic feeder".format(params.coco_dir))
        run(params)
    return 0


if __name__ == "__main__":
    sys.exit(main())
