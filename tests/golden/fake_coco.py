"""Deterministic miniature MSCOCO tree used by both the fixture generator (which feeds it to the REFERENCE's
Captions / Dictionary / Batch_Generator) and the CPU tests (which feed the same tree to this repo's mirrors)."""
import json
import os
import pickle

import numpy as np

WORDS = ("a man dog cat hat riding wave surfboard holding hot laptop sitting on the of in his hand two and red big small "
         "table pizza train bus street person woman umbrella kite").split()


def _sentence(rng, n):
    s = " ".join(WORDS[int(i)] for i in rng.integers(0, len(WORDS), size=n))
    return s.capitalize() + rng.choice([".", "!", "", " ."])


def build(root, n_train=7, n_val=5, n_test=3, write_images=False, seed=21):
    """Creates {root}/coco/{annotations,images/{train,val,test}2014}, {root}/obj_vectors, {root}/pickles.
    Returns the split -> file name lists. Image files are empty unless write_images (then tiny real JPEGs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    coco = os.path.join(root, "coco")
    names = {}
    next_id = 100
    ann_id = 1
    for split, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        d = os.path.join(coco, "images", "%s2014" % split)
        os.makedirs(d)
        files = ["COCO_%s2014_%012d.jpg" % (split, 7 * i + 3) for i in range(n)]
        names[split] = files
        images, anns = [], []
        for fn in files:
            p = os.path.join(d, fn)
            if write_images:
                import cv2
                h, w = int(rng.integers(20, 40)), int(rng.integers(20, 40))
                img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
                cv2.imwrite(p, img)
            else:
                open(p, "wb").close()
            images.append({"id": next_id, "file_name": fn})
            if split != "test":
                for _ in range(int(rng.integers(2, 7))):  # 2..6 captions per image, 3..11 words
                    anns.append({"image_id": next_id, "id": ann_id, "caption": _sentence(rng, int(rng.integers(3, 12)))})
                    ann_id += 1
            next_id += 3
        os.makedirs(os.path.join(coco, "annotations"), exist_ok=True)
        if split == "test":
            with open(os.path.join(coco, "annotations", "image_info_test2014.json"), "w") as f:
                json.dump({"images": images}, f)
        else:
            order = rng.permutation(len(anns))  # annotations are not grouped by image in the real files either
            with open(os.path.join(coco, "annotations", "captions_%s2014.json" % split), "w") as f:
                json.dump({"images": images, "annotations": [anns[i] for i in order]}, f)
    os.makedirs(os.path.join(root, "obj_vectors"))
    os.makedirs(os.path.join(root, "pickles"))
    for fname, splits, skip in (("c_v.pickle", ("train", "val"), 3), ("c_v_test.pickle", ("test",), 2)):
        cv = {}
        for s in splits:
            for i, fn in enumerate(names[s]):
                if i % skip == skip - 1:
                    continue  # some images have no detected objects: the generator substitutes zeros(91)
                v = np.zeros(91)
                k = int(rng.integers(1, 5))
                v[rng.choice(np.arange(1, 91), size=k, replace=False)] = 1.0 / k
                cv[fn] = v
        with open(os.path.join(root, "obj_vectors", fname), "wb") as f:
            pickle.dump(cv, f)
    return names


def feature_dict(names, seed=5, dim=8):
    """{file name: float32 [1, dim]} like ./pickles/train2014.pickle (dim 4096 there)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return {fn: rng.standard_normal((1, dim)).astype(np.float32) for fn in names}


def image_store(root, names, seed=6, hw=4):
    """A uint8 [N, hw, hw, 3] array + ./pickles/itoi.pickle, rows in a scrambled order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    arr = rng.integers(0, 256, size=(len(names), hw, hw, 3), dtype=np.uint8)
    order = rng.permutation(len(names))
    with open(os.path.join(root, "pickles", "itoi.pickle"), "wb") as f:
        pickle.dump({names[j]: int(i) for i, j in enumerate(order)}, f)
    np.save(os.path.join(root, "store.npy"), arr)
    return arr
