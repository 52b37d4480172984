"""Multi-caption batch flattening (reference: utils/caption_utils.py:4-25).

The batch generator yields captions as [B, C, T]; the train step wants one row per caption, ordered
n = b*C + c -- the same order tf.tile gives the image features (main.py:84-89) -- and the cluster
vector of an image repeated for each of its C captions."""
import numpy as np


def preprocess_captions(captions_batch, cl_batch, cv):
    inputs, labels = captions_batch
    B, C, T = inputs.shape
    flat = (np.asarray(inputs).reshape(B * C, T), np.asarray(labels).reshape(B * C, T))
    lengths = np.asarray(cl_batch).reshape(-1)
    if len(cv) != 0:
        cv = np.repeat(np.asarray(cv), C, axis=0)  # row b -> rows b*C .. b*C+C-1
    return flat, lengths, cv
