"""Packs the resized train + val images into one uint8 [N, 224, 224, 3] array and writes the file name -> row map
`./pickles/itoi.pickle` (reference: preprocess.py:10-56). `--output_h5 x.h5` writes HDF5 (dataset "images") when h5py
is importable; a name ending in `.npy` writes a numpy array file that `ImageStore` memory-maps -- same rows, same dtype.

    python -m vae_captioning_b200.preprocess --coco_dir /data/coco --output_h5 train_val.npy
"""
import argparse
import glob
import os
import pickle

import numpy as np

from .image_utils import load_image


def main(params):
    coco_dir, out = params["coco_dir"], params["output_h5"]
    imgs = list(glob.glob(coco_dir + "/images/train2014/" + "*.jpg")) + list(glob.glob(coco_dir + "/images/val2014/" + "*jpg"))
    if len(imgs) == 0:
        raise ValueError
    N = len(imgs)
    h5 = None
    if out.endswith(".npy"):
        dset = np.lib.format.open_memmap(out, mode="w+", dtype=np.uint8, shape=(N, 224, 224, 3))
    else:
        import h5py
        h5 = h5py.File(out, "w")
        dset = h5.create_dataset("images", (N, 224, 224, 3), dtype="uint8")
    imtoi = {}
    for i, path in enumerate(imgs):
        imtoi[path.split("/")[-1]] = i
        dset[i] = load_image(path, shape=(224, 224))
        if i % 1000 == 0:
            print("processing %d/%d (%.2f%% done)" % (i, N, i * 100.0 / N))
    if not os.path.exists("./pickles"):
        os.makedirs("./pickles")
    with open("./pickles/itoi.pickle", "wb") as wf:
        pickle.dump(obj=imtoi, file=wf)
        print("Saved hdf5 imname to indices pickle")
    if h5 is not None:
        h5.close()
    else:
        dset.flush()
    print("wrote ", out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--output_h5", default="train_val.h5", help="output h5 file")
    ap.add_argument("--coco_dir", help="MSCOCO directory")
    main(vars(ap.parse_args()))
