"""Dataset front end (reference: utils/data.py:17-172): caption json -> vocabulary -> indexed captions, batch
generators for the three splits, and the VGG16 fc2 feature cache `./pickles/{train2014,val2014,test2014}.pickle`.

`extract_features_from_dir` is the immediate caller of the VGG16 forward in non-fine-tune mode. The reference runs one
image per `sess.run` (data.py:120-125); here files are decoded and resized on a thread pool into uint8 NHWC batches
and each batch goes through ONE `vc_vgg_forward_u8` call (uint8 H2D, mean subtraction and the whole conv/fc stack on
the device), the next batch being decoded while the device works. The returned / pickled dictionary is the
reference's: {file name: float32 [1, 4096]}."""
import concurrent.futures
import glob
import os
import pickle

import numpy as np

from . import checkpoint
from .batch_gen import Batch_Generator
from .captions import Captions, Dictionary
from .image_utils import load_image


def extract_features(engine, paths, batch_size=None, workers=8, im_shape=(224, 224), progress=None):
    """fc2 features of image files through the device VGG16: -> {basename: float32 [1, 4096]} in `paths` order."""
    bs = int(batch_size or engine.cfg.max_batch)
    if bs > engine.cfg.max_batch:
        raise ValueError("batch_size %d exceeds the engine's max_batch %d" % (bs, engine.cfg.max_batch))
    out = {}
    chunks = [paths[s:s + bs] for s in range(0, len(paths), bs)]
    with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as pool:
        def decode(chunk):
            return np.stack(list(pool.map(lambda p: load_image(p, im_shape), chunk)))

        with concurrent.futures.ThreadPoolExecutor(max_workers=1) as stage:
            pending = stage.submit(decode, chunks[0]) if chunks else None
            for i, chunk in enumerate(chunks):
                images = pending.result()
                pending = stage.submit(decode, chunks[i + 1]) if i + 1 < len(chunks) else None
                feats = engine.vgg_forward(images)  # uint8 [b,224,224,3] -> fp32 [b,4096]
                for p, f in zip(chunk, feats):
                    out[p.split("/")[-1]] = f[None, :].copy()
                if progress:
                    progress(len(out), len(paths))
    return out


class Data(object):
    def __init__(self, params, extract_features=False, weights_path=None, repartiton=False, gen_val_cap=None,
                 engine=None):
        """engine: an `Engine` created with with_cnn=True; only needed when features have to be extracted (no pickle
        yet). Without one it is created on first use from `params` and `weights_path`."""
        coco = params.coco_dir
        self.params = params
        self.train_cap_json = coco + "annotations/captions_train2014.json"
        self.valid_cap_json = coco + "annotations/captions_val2014.json"
        self.test_cap_json = coco + "annotations/image_info_test2014.json"
        self.train_dir = coco + "images/train2014/"
        self.valid_dir = coco + "images/val2014/"
        self.test_dir = coco + "images/test2014/"
        self.captions_tr = Captions(self.train_cap_json, params.cap_max_length)
        self.captions_val = Captions(self.valid_cap_json, params.cap_max_length)
        if not os.path.exists("./pickles"):  # Dictionary pickles the caption dict there (captions.py:122-125)
            os.makedirs("./pickles")
        self.dictionary = Dictionary(self.captions_tr.captions, params.keep_words)
        self.captions_tr.index_captions(self.dictionary.word2idx)
        self.captions_val.index_captions(self.dictionary.word2idx)
        self.train_feature_dict = None
        self.num_examples = self.captions_tr.num_captions
        self.repartiton = repartiton
        self.gen_val_cap = gen_val_cap
        self.engine = engine
        self.weights_path = weights_path
        if repartiton and not gen_val_cap:
            raise ValueError("If using repartition must specify how many val images to use")
        if extract_features:
            if not weights_path:
                raise ValueError("Specify imagenet weights path")
            self.train_feature_dict = self.extract_features_from_dir(self.train_dir)

    # ------------------------------------------------------------------ generators
    def load_train_data_generator(self, batch_size, fine_tune=False, usehdf5=True):
        feature_dict = self.train_feature_dict
        val_cap = valid_feature_dict = None
        if self.repartiton:
            val_cap = self.captions_val
            valid_feature_dict = self.extract_features_from_dir(self.valid_dir)
        if fine_tune or not feature_dict:
            self.train_batch_gen = Batch_Generator(self.train_dir, self.train_cap_json, self.captions_tr, batch_size,
                                                   use_hdf5=self.params.use_hdf5, hdf5_file=self.params.hdf5_file,
                                                   feature_dict=None)
        else:
            self.train_batch_gen = Batch_Generator(self.train_dir, self.train_cap_json, self.captions_tr, batch_size,
                                                   feature_dict=feature_dict)
        if self.repartiton:
            self.train_batch_gen.repartiton(val_cap, valid_feature_dict, self.gen_val_cap)
        return self.train_batch_gen

    def get_valid_data(self, val_batch_size=None, val_tr_unused=None, pretrained=True):
        valid_feature_dict = self.extract_features_from_dir(self.valid_dir) if pretrained else None
        self.valid_batch_gen = Batch_Generator(self.valid_dir, self.valid_cap_json, self.captions_val, val_batch_size,
                                               feature_dict=valid_feature_dict, get_image_ids=True,
                                               val_tr_unused=val_tr_unused, use_hdf5=self.params.use_hdf5,
                                               hdf5_file=self.params.hdf5_file)
        return self.valid_batch_gen

    def get_test_data(self, test_batch_size=None, pretrained=True):
        test_feature_dict = self.extract_features_from_dir(self.test_dir) if pretrained else None
        self.train_batch_gen = Batch_Generator(self.test_dir, train_cap_json=self.test_cap_json,
                                               batch_size=test_batch_size, feature_dict=test_feature_dict,
                                               get_image_ids=True, get_test_ids=True)
        return self.train_batch_gen

    # ------------------------------------------------------------------ feature cache
    def _vgg_engine(self, batch_size):
        if self.engine is None:
            from .engine import Engine
            eng = Engine(self.params, vocab_size=self.dictionary.vocab_size, max_batch=batch_size, with_cnn=True)
            if not self.weights_path:
                raise ValueError("Specify imagenet weights path")
            eng.load_state(checkpoint.vgg16_npz_state(self.weights_path))  # vgg16.load_weights (image_embeddings.py:240-246)
            self.engine = eng
        return self.engine

    def extract_features_from_dir(self, data_dir, save_pickle=True, im_shape=(224, 224), batch_size=256):
        """{image file name: fc2 feature [1, 4096]} of every *.jpg in data_dir; cached in ./pickles/{split}.pickle."""
        pkl = "./pickles/" + data_dir.split("/")[-2] + ".pickle"
        try:
            with open(pkl, "rb") as rf:
                print("Loading prepared feature vector from {}".format(pkl))
                return pickle.load(rf)
        except (OSError, EOFError, pickle.UnpicklingError):
            pass
        print("Extracting features")
        if not os.path.exists("./pickles"):
            os.makedirs("./pickles")
        paths = list(glob.glob(data_dir + "*.jpg"))
        eng = self._vgg_engine(batch_size)
        feature_dict = extract_features(eng, paths, min(batch_size, eng.cfg.max_batch), im_shape=im_shape)
        if save_pickle:
            with open(pkl, "wb") as wf:
                pickle.dump(feature_dict, wf)
        return feature_dict
