"""Checkpoint surface of the reference (`tf.train.Saver` by variable name, main.py:186-191, 211, 288; the ImageNet
`vgg16_weights.npz` loader, utils/image_embeddings.py:240-246).

Variables are exchanged by their TensorFlow names and layouts, so a state dict produced here maps one to
one onto the reference's checkpoint. The on-disk container is `.npz` (one array per variable name); reading and writing
TF's V2 tensor-bundle files is a "next" row of the scope table. Optimiser state is not saved -- the
reference's Saver holds only the trainable variables (+ the cnn/ variables when they are frozen).
"""
import os

import numpy as np

# creation order of vgg16.parameters (utils/image_embeddings.py:36-238): conv kernels/biases, then fc1, fc2
VGG_VARIABLES = []
for _name in ("conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
              "conv5_1", "conv5_2", "conv5_3"):
    _suffix = "_conv" if _name.startswith("conv5") else ""  # image_embeddings.py:176-201 names them weights_conv / biases_conv
    VGG_VARIABLES += ["cnn/%s/weights%s" % (_name, _suffix), "cnn/%s/biases%s" % (_name, _suffix)]
VGG_VARIABLES += ["cnn/fc1/weights", "cnn/fc1/biases", "cnn/fc2/weights", "cnn/fc2/biases"]


def checkpoint_path(params, directory="./checkpoints"):
    """'./checkpoints/{checkpoint}.ckpt' of the reference (main.py:211, 288) with the .npz container suffix."""
    return os.path.join(directory, "{}.ckpt.npz".format(params.checkpoint))


def save(path, state):
    """state: {tf_variable_name: array}. Names contain '/', which np.savez keeps verbatim as archive member names."""
    d = os.path.dirname(path)
    if d and not os.path.exists(d):
        os.makedirs(d)
    np.savez(path, **{k: np.asarray(v, dtype=np.float32) for k, v in state.items()})
    return path


def load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def restore(engine, path, strict=True):
    """saver.restore: every variable of the engine must be present with its shape (TF raises otherwise)."""
    state = load(path)
    for name, shape, _ in engine.variables():
        if name not in state:
            if strict:
                raise KeyError("checkpoint %s has no variable %s" % (path, name))
            continue
        if tuple(state[name].shape) != tuple(shape):
            raise ValueError("variable %s: checkpoint shape %s, graph shape %s" % (name, state[name].shape, shape))
        engine.set_variable(name, state[name])
    return state


def vgg16_npz_state(weight_file):
    """vgg16.load_weights: the npz keys are sorted and assigned, in that order, to the first 30 graph parameters
    ('original file contains weights we dont need': fc8 is skipped). -> {tf name: array}."""
    with np.load(weight_file) as z:
        keys = sorted(z.files)
        if len(keys) < 30:
            raise ValueError("%s holds %d arrays, expected at least 30" % (weight_file, len(keys)))
        # sorted() puts conv1_1_W, conv1_1_b, ..., conv5_3_b, fc6_W, fc6_b, fc7_W, fc7_b, (fc8_W, fc8_b)
        return {VGG_VARIABLES[i]: z[k] for i, k in enumerate(keys[:30])}
