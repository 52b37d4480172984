"""Generation-side mirror of the reference's `Decoder` (vae_model/decoder.py:145-320) on top of the C ABI.

`Decoder.online_inference` and `Decoder.beam_search` keep the reference's arguments (minus the TF session and
placeholders) and return values (`cap_list` of {'image_id', 'caption'} dicts, `cap_raw`), but decode the whole batch
of images on the device in one call instead of one `sess.run` per token per beam."""
import ctypes

import numpy as np

from . import lib as L
from .engine import VcRng, _f32, _np_ptr


class Decoder(object):
    accepts_rng = True  # online_inference / beam_search take rng= (inference.inference passes one seed per batch)

    def __init__(self, engine, params, data_dict):
        """engine: vae_captioning_b200.engine.Engine; params: Parameters; data_dict: Dictionary (word2idx / idx2word)."""
        self.engine = engine
        self.params = params
        self.data_dict = data_dict
        self._declare()

    def _declare(self):
        lib = self.engine.lib
        if getattr(lib, "_vc_decode_declared", False):
            return
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.vc_decode_greedy.argtypes = [vp, vp, vp, ci, ci, ci, ctypes.POINTER(VcRng), ci, ci, vp, vp, vp]
        lib.vc_decode_beam.argtypes = [vp, vp, vp, ci, ci, ci, cf, ctypes.POINTER(VcRng), ci, ci, vp, vp, vp, vp, vp]
        lib.vc_decode_begin.argtypes = [vp, vp, vp, ci, ctypes.POINTER(VcRng), vp]
        lib.vc_decode_step.argtypes = [vp, vp, ci, vp, vp]
        lib.vc_decode_state_get.argtypes = [vp, vp, vp, vp]
        lib.vc_decode_state_set.argtypes = [vp, vp, vp, vp]
        lib._vc_decode_declared = True

    # ------------------------------------------------------------------ raw device loops
    def _feed(self, in_pictures, c_v):
        pix = np.asarray(in_pictures)
        if pix.ndim == 4 and pix.shape[1:] == (224, 224, 3):
            # raw images: generation after --fine_tune runs them through the fine-tuned VGG16 first ("caption generation
            # after fine_tune must be made with params.fine_tune=True", main.py:46-51 + decoder.py:174-183 feed images)
            if not self.engine.cfg.with_cnn:
                raise ValueError("in_pictures are images but this engine was created without the CNN (fine_tune / with_cnn)")
            mb = int(self.engine.cfg.max_batch)
            in_pictures = np.concatenate([self.engine.vgg_forward(pix[i:i + mb]) for i in range(0, pix.shape[0], mb)])
        feats = _f32(in_pictures)
        if feats.ndim != 2 or feats.shape[1] != self.engine.cfg.cnn_feature_size:
            raise ValueError("in_pictures must be [B, %d] features, got %s" % (self.engine.cfg.cnn_feature_size, feats.shape))
        cv = None
        feeds_c_i = self.params.use_c_v or self.params.prior in ("GMM", "AG")  # decoder.c_i is set only then (main.py:103-110)
        if feeds_c_i and c_v is not None and len(c_v) != 0:
            cv = _f32(c_v)
            if cv.shape != (feats.shape[0], self.engine.cfg.num_clusters):
                raise ValueError("c_v must be [%d, %d], got %s" % (feats.shape[0], self.engine.cfg.num_clusters, cv.shape))
        return feats, cv

    def greedy_tokens(self, in_pictures, c_v=None, mode="greedy", rng=None, max_len=None):
        """-> (tokens int32 [B, gen_max_len] zero padded, lengths [B]): cap_raw of online_inference."""
        feats, cv = self._feed(in_pictures, c_v)
        B = feats.shape[0]
        max_len = int(max_len or self.params.gen_max_len)
        toks = np.zeros((B, max_len), np.int32)
        lens = np.zeros((B,), np.int32)
        r, keep = self.engine._rng(rng)
        w2i = self.data_dict.word2idx
        L.check(self.engine.lib.vc_decode_greedy(self.engine._h, _np_ptr(feats), _np_ptr(cv), B, max_len,
                                                 {"greedy": 0, "sample": 1}[mode], ctypes.byref(r), w2i["<BOS>"], w2i["<EOS>"],
                                                 _np_ptr(toks), _np_ptr(lens), self.engine._stream()))
        return toks, lens

    def beam_tokens(self, in_pictures, c_v=None, beam_size=2, len_norm_f=0.7, rng=None, max_len=None):
        """-> (tokens [B, beam, gen_max_len] incl. <BOS>/<EOS>, lengths [B, beam], scores [B, beam], n_beams [B])."""
        feats, cv = self._feed(in_pictures, c_v)
        B = feats.shape[0]
        max_len = int(max_len or self.params.gen_max_len)
        toks = np.zeros((B, beam_size, max_len), np.int32)
        lens = np.zeros((B, beam_size), np.int32)
        scores = np.zeros((B, beam_size), np.float32)
        nb = np.zeros((B,), np.int32)
        r, keep = self.engine._rng(rng)
        w2i = self.data_dict.word2idx
        L.check(self.engine.lib.vc_decode_beam(self.engine._h, _np_ptr(feats), _np_ptr(cv), B, int(beam_size), max_len,
                                               float(len_norm_f), ctypes.byref(r), w2i["<BOS>"], w2i["<EOS>"], _np_ptr(toks),
                                               _np_ptr(lens), _np_ptr(scores), _np_ptr(nb), self.engine._stream()))
        return toks, lens, scores, nb

    # ------------------------------------------------------------------ the reference's granularity (one sess.run)
    def begin(self, in_pictures, c_v=None, rng=None):
        feats, cv = self._feed(in_pictures, c_v)
        r, keep = self.engine._rng(rng)
        L.check(self.engine.lib.vc_decode_begin(self.engine._h, _np_ptr(feats), _np_ptr(cv), feats.shape[0], ctypes.byref(r),
                                                self.engine._stream()))
        self._rows = feats.shape[0]

    def step(self, tokens, probs=True):
        """sess.run([sample, out_state], {captions: tokens[:, None], in_state: <current>}) -> softmax probs [M, V]."""
        tok = np.ascontiguousarray(tokens, dtype=np.int32).ravel()
        out = np.empty((tok.size, self.engine.cfg.vocab_size), np.float32) if probs else None
        L.check(self.engine.lib.vc_decode_step(self.engine._h, _np_ptr(tok), tok.size, _np_ptr(out), self.engine._stream()))
        return out

    def get_state(self):
        H = self.engine.cfg.decoder_hidden
        c = np.empty((self._rows, H), np.float32)
        h = np.empty((self._rows, H), np.float32)
        L.check(self.engine.lib.vc_decode_state_get(self.engine._h, _np_ptr(c), _np_ptr(h), self.engine._stream()))
        return c, h

    def set_state(self, c, h):
        c, h = _f32(c), _f32(h)
        L.check(self.engine.lib.vc_decode_state_set(self.engine._h, _np_ptr(c), _np_ptr(h), self.engine._stream()))

    # ------------------------------------------------------------------ reference-shaped methods
    def _caption(self, ids):
        i2w = self.data_dict.idx2word
        return " ".join(i2w[int(w)] for w in ids if i2w[int(w)] not in ("<BOS>", "<EOS>"))

    def online_inference(self, picture_ids, in_pictures, c_v=None, stop_word="<EOS>", sample_gen=None, rng=None):
        """decoder.py:145-201 -> (cap_list, cap_raw). sample_gen defaults to params.sample_gen; any value other than
        'greedy' / 'sample' reproduces the reference's behaviour for that case: gen_word_idx stays 0, i.e. gen_max_len
        x <PAD> -- no device work is needed to produce that."""
        mode = sample_gen or self.params.sample_gen
        n = len(in_pictures)
        if mode not in ("greedy", "sample"):
            raw = [[0] * self.params.gen_max_len for _ in range(n)]
        else:
            toks, lens = self.greedy_tokens(in_pictures, c_v, mode, rng)
            raw = [[int(w) for w in toks[i, :lens[i]]] for i in range(n)]
        cap_list = [{"image_id": picture_ids[i], "caption": self._caption(raw[i])} for i in range(n)]
        return cap_list, raw

    def beam_search(self, picture_ids, in_pictures, c_v=None, beam_size=2, ret_beams=False, len_norm_f=0.7, rng=None):
        """decoder.py:203-320 -> cap_list (caption is a list of strings when ret_beams)."""
        toks, lens, scores, nb = self.beam_tokens(in_pictures, c_v, beam_size, len_norm_f, rng)
        cap_list = []
        for i in range(len(in_pictures)):
            beams = [self._caption(toks[i, b, :lens[i, b]]) for b in range(int(nb[i]))]
            cap_list.append({"image_id": picture_ids[i], "caption": beams if ret_beams else beams[0]})
        return cap_list
