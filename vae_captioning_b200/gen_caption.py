"""Single-image caption generator (reference: gen_caption.py:19-160): load the pickled `Parameters` and the caption
vocabulary, restore the trained variables by name, extract the image's VGG16 fc2 feature and decode greedily or by beam
search.

What changes against the reference is only where the arithmetic runs: the fc2 feature comes from this library's
device VGG16 (`Engine.vgg_forward`, weights = the checkpoint's `cnn/*` variables or, if the checkpoint holds none,
`params.image_net_weights_path`) instead of a Keras ImageNet model, and decoding is the batched device decode behind
`Decoder.online_inference` / `Decoder.beam_search`. `ret_beams=True` returns all beams (decoder.py:311-319), and
`c_v_generator` fills the stub the reference left at gen_caption.py:40-42: a callable image -> [91] or [90] cluster vector.

    python -m vae_captioning_b200.gen_caption --img_path x.jpg --checkpoint ./checkpoints/default.ckpt \\
        --params_path ./pickles/params_Normal_False_default_False.pickle --gen_method beam_search --beam_size 5
"""
import argparse
import io
import os
import pickle

import numpy as np

from . import checkpoint
from .captions import Dictionary
from .image_utils import load_image
from .parameters import Parameters


class _ParamsUnpickler(pickle.Unpickler):
    """Parameter pickles written by the reference name the class `utils.parameters.Parameters` (main.py:305-313)."""

    def find_class(self, module, name):
        if name == "Parameters" and module in ("utils.parameters", "parameters"):
            return Parameters
        return pickle.Unpickler.find_class(self, module, name)


def load_params(params_path):
    with open(params_path, "rb") as rf:
        return _ParamsUnpickler(io.BytesIO(rf.read())).load()


class Generator(object):
    def __init__(self, checkpoint_path, params_path, vocab_path, gen_method="greedy", c_v_generator=None, device=0):
        self.checkpoint_path = checkpoint_path
        self.params = load_params(params_path)
        self.gen_method = gen_method
        self.c_v_generator = c_v_generator
        self.device = device
        if not vocab_path or not os.path.exists(vocab_path):
            raise ValueError("No caption vocabulary path specified, Usually it can be found in the ./pickles foulder "
                             "after model training")
        with open(vocab_path, "rb") as rf:
            data_dict = pickle.load(rf)
        # the pickle holds the raw training caption dict (captions.py:122-125); ids are re-derived from it
        self.data_dict = Dictionary(data_dict, getattr(self.params, "keep_words", 3), save_path=None)
        self.params.vocab_size = self.data_dict.vocab_size
        self._engine = None
        self._decoder = None

    def _c_v_generator(self, image):
        if self.c_v_generator is None:
            return None
        v = np.asarray(self.c_v_generator(image), dtype=np.float32).reshape(1, -1)
        if v.shape[1] == 91:
            v = v[:, 1:]  # column 0 (background) is never fed (main.py:236)
        if v.shape[1] != 90:
            raise ValueError("cluster vector generator must return 91 or 90 values, got %d" % v.shape[1])
        return v

    def _build(self):
        """The graph of gen_caption.py:75-115 (imf_emb, [cv_emb,] decoder/*) + the feature extractor, restored by name."""
        if self._engine is not None:
            return
        from .decode import Decoder
        from .engine import Engine
        self.params.sample_gen = self.gen_method
        eng = Engine(self.params, vocab_size=self.data_dict.vocab_size, max_batch=1, max_len=self.params.gen_max_len,
                     device=self.device, with_cnn=True)
        state = checkpoint.load(self.checkpoint_path)
        names = {n for n, _, _ in eng.variables()}
        eng.load_state({k: v for k, v in state.items() if k in names})
        if not any(k.startswith("cnn/") for k in state):
            eng.load_state(checkpoint.vgg16_npz_state(self.params.image_net_weights_path))
        missing = sorted(n for n in names if n not in state and not n.startswith("cnn/") and not n.startswith("encoder/"))
        if missing:  # Saver.restore raises NotFoundError for a variable the checkpoint lacks
            raise KeyError("checkpoint %s has no variable %s" % (self.checkpoint_path, missing[0]))
        if self.params.prior == "AG" and not self.params.no_encoder:
            # decoder.cap_clusters = c_means (decoder.py:45-71): the generation prior's mean is built from the cluster
            # means the model was trained with (./pickles/cluster_means.pickle, utils/vae_utils.py:7)
            from . import synthetic
            from .main import CLUSTER_MEANS_FILE
            eng.set_cluster_means(synthetic.init_clusters(self.params.num_clusters, self.params.latent_size,
                                                          c_m_file=CLUSTER_MEANS_FILE))
        self._engine = eng
        self._decoder = Decoder(eng, self.params, self.data_dict)

    def _get_features(self, img_path):
        """-> (fc2 feature [1, 4096], uint8 RGB image [224, 224, 3])."""
        img = load_image(img_path, (224, 224))
        return self._engine.vgg_forward(img[None]), img

    def generate_caption(self, img_path, beam_size=2, ret_beams=False):
        if not os.path.exists(img_path):
            raise ValueError("Image not found")
        self._build()
        im_id = [img_path.split("/")[-1]]
        feature_vector, image = self._get_features(img_path)
        c_v = self._c_v_generator(image) if self.params.use_c_v else None
        if self.params.use_c_v and c_v is None:
            raise ValueError("the model was trained with cluster vectors (--c_v): pass c_v_generator")
        if self.gen_method == "beam_search":
            return self._decoder.beam_search(im_id, feature_vector, c_v, beam_size=int(beam_size), ret_beams=ret_beams)
        sent, _ = self._decoder.online_inference(im_id, feature_vector, c_v=c_v, sample_gen=self.gen_method)
        return sent

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


def main(argv=None):
    ap = argparse.ArgumentParser(description="Specify generation parameters")
    ap.add_argument("--img_path", help="Path to the image")
    ap.add_argument("--checkpoint", help="Model checkpoint path")
    ap.add_argument("--vocab_path", default="./pickles/capt_vocab.pickle", help="Indices to words dictionary")
    ap.add_argument("--gpu", default="", help="Specify GPU number if use GPU")
    ap.add_argument("--c_v_generator", default=None, help="If use cluster vectors: module:function returning the vector")
    ap.add_argument("--gen_method", default="greedy", help="greedy, beam_search or sample")
    ap.add_argument("--params_path", default=None, help="specify params pickle file")
    ap.add_argument("--beam_size", default=2, help="If using beam_search, specify beam_size")
    args = ap.parse_args(argv)
    os.environ["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
    if args.gpu != "":
        os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu
    cvg = None
    if args.c_v_generator:
        import importlib
        mod, fn = args.c_v_generator.split(":")
        cvg = getattr(importlib.import_module(mod), fn)
    generator = Generator(checkpoint_path=args.checkpoint, params_path=args.params_path, vocab_path=args.vocab_path,
                          gen_method=args.gen_method, c_v_generator=cvg)
    caption = generator.generate_caption(args.img_path, args.beam_size)
    print(caption[0]["caption"])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
