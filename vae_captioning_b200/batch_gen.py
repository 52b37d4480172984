"""Host batcher either side of the train step (reference: utils/batch_gen.py:16-369).

`Batch_Generator` keeps the reference's constructor, generators and yields -- `next_batch` / `next_val_batch` /
`next_test_batch` give `(images_or_features, (input_captions, label_captions), lengths[, image_ids], cluster_vectors)`
with captions padded to the batch maximum and `lengths = len(caption) - 1` as float64 -- and the same consumption of
the two random streams (`random.shuffle` of the file list, `numpy.random.randint` for the caption pick, seed 42), so a
run is reproducible against the reference batch by batch (tests/golden/batch_gen.json was produced by the reference).

Differences that do not change what is yielded:
  * the uint8 image store (`use_hdf5`) is read through `ImageStore`: an HDF5 file when h5py is importable, or a `.npy`
    array written by `vae_captioning_b200.preprocess` (memory-mapped); rows come back as uint8 NHWC, which is what the
    device entry points take (`Engine.train_step(..., images=True)`), a quarter of the fp32 H2D bytes;
  * `Prefetcher` runs any of the generators one batch ahead on a thread so file reads overlap the device step.
"""
import glob
import json
import os
import pickle
import queue
import random
import threading

import numpy as np

from .image_utils import load_image


class ImageStore(object):
    """uint8 [N, 224, 224, 3] image array on disk + file name -> row (preprocess.py:26-45, batch_gen.py:37-43, 278-287)."""

    def __init__(self, path, index_pickle="./pickles/itoi.pickle"):
        with open(index_pickle, "rb") as rf:
            self.imtoi = pickle.load(rf)
        if path.endswith(".npy"):
            self._file = None
            self.images = np.load(path, mmap_mode="r")
        else:
            try:
                import h5py
            except ImportError:
                raise ImportError("reading %s needs h5py; vae_captioning_b200.preprocess can write a .npy store instead" % path)
            self._file = h5py.File(path, "r")
            self.images = self._file["images"]

    def rows(self, indices):
        """indices must be increasing (an HDF5 fancy-indexing rule the reference sorts for, batch_gen.py:151-161)."""
        return np.asarray(self.images[list(indices)])


class Batch_Generator(object):
    def __init__(self, train_dir, train_cap_json=None, captions=None, batch_size=None, use_hdf5=False, hdf5_file=None,
                 feature_dict=None, get_image_ids=False, get_test_ids=False, val_tr_unused=None):
        self.use_hdf5 = use_hdf5
        if use_hdf5:
            if not hdf5_file:
                raise ValueError("Specify hdf5 file path")
            store = ImageStore(hdf5_file)
            self.imtoi, self.images = store.imtoi, store.images
            self._store = store
        self._batch_size = batch_size
        if val_tr_unused is None:
            self._iterable = list(glob.glob(train_dir + "*.jpg"))
        else:
            print("Val captions for generation : ", len(val_tr_unused))
            self._iterable = val_tr_unused
        self._train_dir = train_dir
        if not batch_size:
            print("use all data")
            self._batch_size = len(self._iterable)
        if len(self._iterable) == 0:
            print("Check images files avaliability")
            print("Coco dir: ", train_dir)
            raise FileNotFoundError
        self._train_cap_json = train_cap_json
        if get_test_ids:  # the test split has no captions, only image ids
            self._fn_to_id = self._test_images_to_imid()
        if captions:
            self.cap_instance = captions
            self.captions = captions.captions_indexed
        self.random_seed = 42
        np.random.seed(self.random_seed)
        self.feature_dict = feature_dict
        self.get_image_ids = get_image_ids
        self.unused_cap_in = None

    # ------------------------------------------------------------------ train + val repartition (batch_gen.py:70-99)
    def repartiton(self, val_cap_instance, val_feature_dict, gen_val_cap):
        self.gen_val_cap = gen_val_cap
        if not val_cap_instance:
            raise ValueError("If use validation set images for training need to specify val_cap instance")
        self.val_cap_instance = val_cap_instance
        self.val_captions = val_cap_instance.captions_indexed
        val_dir = "/".join(self._train_dir.split("/")[:-2] + ["val2014/"])
        val_list = list(glob.glob(val_dir + "*.jpg"))
        random.shuffle(val_list)
        if self.gen_val_cap is not None and self.gen_val_cap < 0:
            self.gen_val_cap = None
        if self.gen_val_cap:  # the last gen_val_cap validation images stay out of training
            self._iterable.extend(val_list[:-self.gen_val_cap])
            self.unused_cap_in = val_list[-self.gen_val_cap:]
        else:
            self._iterable.extend(val_list)
        print("Train + Validation set size (use repartition): ", len(self._iterable))
        if not val_feature_dict:
            raise ValueError("If use validation set images for training need to specify val_feature_dict")
        self.val_feature_dict = val_feature_dict

    # ------------------------------------------------------------------ per-batch assembly
    @staticmethod
    def _base(name):
        return name.split("/")[-1]

    def _lookup(self, first, second_attr, key):
        """first[key], falling back to the validation-split container when training on train + val."""
        try:
            return first[key]
        except Exception:
            return getattr(self, second_attr)[key]

    def _images_c_v(self, imn_batch, c_v, indices=None):
        """-> ([B, 4096] features or [B, 224, 224, 3] images, [B, 91] cluster vectors or an empty array)."""
        feats, cl_v = [], []
        if c_v or self.feature_dict:
            for imn in imn_batch:
                key = self._base(imn)
                if c_v:
                    cl_v.append(c_v[key] if key in c_v else np.zeros(91))
                if self.feature_dict:
                    feats.append(self._lookup(self.feature_dict, "val_feature_dict", key))
        cl_v = np.array(cl_v)
        if self.feature_dict:
            images = np.squeeze(np.array(feats), 1)  # the feature pickles hold [1, 4096] rows (data.py:123-125)
        else:
            images = self._get_images(imn_batch, indices)
        return images, cl_v

    def _get_imid(self, imn_batch, test=False):
        if not test:
            return [self._lookup(self.cap_instance.filename_to_imid, "_val_imid", self._base(fn)) for fn in imn_batch]
        ids = [self._fn_to_id[self._base(fn)] for fn in imn_batch]
        # batch_gen.py:139-142 nests the test-split loop inside the per-file loop, so the id list comes back B times
        # over; consumers index it by row, so only the first B entries are ever read. Kept for identical yields.
        return ids * len(imn_batch)

    @property
    def _val_imid(self):
        return self.val_cap_instance.filename_to_imid

    def _next_imn(self):
        idx = np.random.choice(range(len(self._iterable)), self._batch_size, replace=False)
        return np.array(self._iterable)[idx]

    def _get_indices(self, imn_batch):
        """Rows of the image store for a batch, reordered by (row, name) as HDF5 wants increasing indices."""
        pairs = sorted(((self._base(n), self.imtoi[self._base(n)]) for n in imn_batch), key=lambda p: (p[1], p[0]))
        return [p[0] for p in pairs], [p[1] for p in pairs]

    def _get_images(self, names, indices=None):
        if self.use_hdf5:
            return self._store.rows(indices)
        return np.stack([load_image(n) for n in names])

    def _chunks(self):
        """The file list cut into batches of _batch_size; the tail batch is shorter (batch_gen.py:183-206)."""
        bs = self._batch_size
        for s in range(0, len(self._iterable), bs):
            yield list(self._iterable[s:s + bs])

    def _assemble(self, chunk, c_v, with_captions=True, **cap_kw):
        indices = None
        if self.use_hdf5 and with_captions:
            chunk, indices = self._get_indices(chunk)
        images, cl_v = self._images_c_v(chunk, c_v, indices)
        if not with_captions:
            return chunk, images, cl_v, None
        return chunk, images, cl_v, self._form_captions_batch(chunk, **cap_kw)

    # ------------------------------------------------------------------ the three generators
    def next_batch(self, use_obj_vectors=False, num_captions=1):
        """Training batches: shuffled file list; num_captions == 1 picks one random caption per image, otherwise the
        first num_captions captions of every image are returned as [B, C, T]."""
        self.use_obj_vectors = use_obj_vectors
        c_v = self._get_cluster_vectors() if use_obj_vectors else None
        random.shuffle(self._iterable)
        for chunk in self._chunks():
            _, images, cl_v, (inp, lbl, lengths) = self._assemble(chunk, c_v, random_select=num_captions == 1,
                                                                  num_captions=num_captions)
            yield images, (inp, lbl), lengths, cl_v

    def next_val_batch(self, get_image_ids=False, use_obj_vectors=False):
        self.get_image_ids = get_image_ids
        self.use_obj_vectors = use_obj_vectors
        c_v = self._get_cluster_vectors() if use_obj_vectors else None
        for chunk in self._chunks():
            chunk, images, cl_v, (inp, lbl, lengths) = self._assemble(chunk, c_v)
            if self.get_image_ids:
                yield images, (inp, lbl), lengths, self._get_imid(chunk), cl_v
            else:
                yield images, (inp, lbl), lengths, cl_v

    def next_test_batch(self, use_obj_vectors=False):
        self.use_obj_vectors = use_obj_vectors
        c_v = self._get_cluster_vectors(True) if use_obj_vectors else None
        for chunk in self._chunks():
            chunk, images, cl_v, _ = self._assemble(chunk, c_v, with_captions=False)
            yield images, self._get_imid(chunk, True), cl_v

    def _test_images_to_imid(self):
        with open(self._train_cap_json) as rf:
            j = json.loads(rf.read())
        return {img["file_name"]: img["id"] for img in j["images"]}

    # ------------------------------------------------------------------ captions -> padded id arrays
    def _form_captions_batch(self, imn_batch, random_select=True, num_captions=1):
        """-> (inputs, labels, lengths): inputs = caption[:-1] (<BOS> ...), labels = caption[1:] (... <EOS>), both
        zero-padded to the longest caption of the batch; [B, C, T] / lengths [B, C], or [B, T] / [B] when C == 1."""
        if random_select:
            num_captions = 1
        B = len(imn_batch)
        inputs = [[[0] for _ in range(num_captions)] for _ in range(B)]
        labels = [[[0] for _ in range(num_captions)] for _ in range(B)]
        lengths = np.zeros((B, num_captions))
        for b, fn in enumerate(imn_batch):
            key = self._base(fn)
            caps = self.captions[key]
            if len(caps) == 0:  # an image of the validation split (train + val repartition)
                caps = self.val_captions[key]
            if random_select:
                caps = [caps[np.random.randint(len(caps))]]
            for c, cap in enumerate(caps[:num_captions]):
                inputs[b][c] = cap[:-1]
                labels[b][c] = cap[1:]
                lengths[b][c] = len(cap) - 1
        pad = max(len(cap) for caps in inputs for cap in caps)
        inputs = np.array([[cap + [0] * (pad - len(cap)) for cap in caps] for caps in inputs])
        labels = np.array([[cap + [0] * (pad - len(cap)) for cap in caps] for caps in labels])
        if inputs.shape[1] == 1:
            inputs, labels, lengths = np.squeeze(inputs, 1), np.squeeze(labels, 1), np.squeeze(lengths)
        return inputs, labels, lengths

    def _get_cluster_vectors(self, load_test=False):
        """{file name: [91] cluster vector} from ./obj_vectors (batch_gen.py:347-362)."""
        path = "./obj_vectors/c_v_test.pickle" if load_test else "./obj_vectors/c_v.pickle"
        with open(path, "rb") as rf:
            c_v = pickle.load(rf)
        assert type(c_v) == dict, "cluster vector pickle must contain dict"
        return c_v

    @property
    def cap_dict(self):
        return self._cap_dict

    def set_bs(self, batch_size):
        self._batch_size = batch_size


class Prefetcher(object):
    """Runs a batch generator `depth` batches ahead on a worker thread, so json/pickle lookups, padding and image-store
    reads of batch i+1 overlap the device step of batch i. Yields exactly what the wrapped generator yields, in order;
    an exception in the worker is re-raised at the consumer."""

    _END = object()

    def __init__(self, generator, depth=2):
        self._q = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._work, args=(generator,), daemon=True)
        self._t.start()

    def _put(self, entry):
        while not self._stop.is_set():
            try:
                self._q.put(entry, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _work(self, generator):
        try:
            for item in generator:
                if not self._put((item, None)):
                    return
            self._put((self._END, None))
        except BaseException as e:  # noqa: B036 -- handed to the consumer
            self._put((self._END, e))

    def close(self):
        """Stops the worker (the consumer left the loop early, e.g. main.py's num_ex_per_epoch stop condition)."""
        self._stop.set()

    def __iter__(self):
        try:
            while True:
                item, err = self._q.get()
                if err is not None:
                    raise err
                if item is self._END:
                    return
                yield item
        finally:
            self.close()
