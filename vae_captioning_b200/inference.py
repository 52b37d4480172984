"""Inference driver (reference: ops/inference.py:4-56): iterate the validation and test generators, decode every
batch on the device, and write `val_{gen_name}.json` / `test_{gen_name}.json` in the COCO-caption result format
[{'image_id': ..., 'caption': ...}].

`decoder` is a `vae_captioning_b200.decode.Decoder` (or anything with its beam_search / online_inference methods):
the batched device decode replaces the reference's one-`sess.run`-per-token-per-beam loops.

The reference always sends the test split through online_inference, which ignores sample_gen='beam_search' and
emits 30 x <PAD>. By default this driver decodes the test split greedily instead;
`reference_test_split_bug=True` keeps the reference's call pattern (used by the parity test of the json layout)."""
import json
import os


def _drop_background_column(params, c_v, val):
    uses = params.use_c_v or (val and params.prior in ("GMM", "AG"))
    return c_v[:, 1:] if uses else c_v  # column 0 of the 91-wide cluster vectors is never used (ops/inference.py:19-21)


def inference(params, decoder, val_gen, test_gen, out_dir=".", reference_test_split_bug=False, verbose=True, seed=0):
    say = print if verbose else (lambda *a, **k: None)
    captions_gen = []
    n_call = [0]

    def rng():  # every sess.run of the reference draws fresh z noise / multinomial samples: one Philox stream per batch
        n_call[0] += 1
        return {"seed": (int(seed) << 24) + n_call[0]}

    def kw():  # reference-style decoders (no rng argument) still plug in
        return {"rng": rng()} if getattr(decoder, "accepts_rng", False) else {}

    say("Generating captions for val file")
    for feats, _, _, image_ids, c_v in val_gen.next_val_batch(get_image_ids=True, use_obj_vectors=params.use_c_v):
        c_v = _drop_background_column(params, c_v, True)
        if params.sample_gen == "beam_search":
            sent = decoder.beam_search(image_ids, feats, c_v, beam_size=params.beam_size, **kw())
        else:
            sent, _ = decoder.online_inference(image_ids, feats, c_v=c_v, **kw())
        captions_gen += sent
    say("Generated {} captions".format(len(captions_gen)))
    val_file = os.path.join(out_dir, "val_{}.json".format(params.gen_name))
    with open(val_file, "w") as f:
        json.dump(captions_gen, f)
    captions_gen = []
    say("Generating captions for test file")
    for feats, image_ids, c_v in test_gen.next_test_batch(params.use_c_v):
        c_v = _drop_background_column(params, c_v, False)
        if reference_test_split_bug or params.sample_gen != "beam_search":
            sent, _ = decoder.online_inference(image_ids, feats, c_v=c_v, **kw())
        else:
            sent, _ = decoder.online_inference(image_ids, feats, c_v=c_v, sample_gen="greedy", **kw())
        captions_gen += sent
    test_file = os.path.join(out_dir, "test_{}.json".format(params.gen_name))
    with open(test_file, "w") as f:
        json.dump(captions_gen, f)
    return val_file, test_file
