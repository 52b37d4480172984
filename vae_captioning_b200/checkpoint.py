"""Checkpoint surface of the reference (`tf.train.Saver` by variable name, main.py:186-191, 211, 288; the ImageNet
`vgg16_weights.npz` loader, utils/image_embeddings.py:240-246).

Variables are exchanged by their TensorFlow names and layouts, so a state dict produced here maps one to
one onto the reference's checkpoint. The on-disk container is TensorFlow's V2 tensor bundle
(`{name}.ckpt.index` + `{name}.ckpt.data-00000-of-00001`, written and read by `tf_bundle` without TensorFlow, SURVEY
8f-2) plus the `checkpoint` state file Saver keeps beside it; a path ending in `.npz` selects a plain numpy archive
instead (one array per variable name). Optimiser state is not saved -- the reference's Saver holds only the trainable
variables (+ the cnn/ variables when they are frozen).
"""
import os

import numpy as np

from . import tf_bundle

# creation order of vgg16.parameters (utils/image_embeddings.py:36-238): conv kernels/biases, then fc1, fc2
VGG_VARIABLES = []
for _name in ("conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
              "conv5_1", "conv5_2", "conv5_3"):
    _suffix = "_conv" if _name.startswith("conv5") else ""  # image_embeddings.py:176-201 names them weights_conv / biases_conv
    VGG_VARIABLES += ["cnn/%s/weights%s" % (_name, _suffix), "cnn/%s/biases%s" % (_name, _suffix)]
VGG_VARIABLES += ["cnn/fc1/weights", "cnn/fc1/biases", "cnn/fc2/weights", "cnn/fc2/biases"]


def checkpoint_path(params, directory="./checkpoints"):
    """'./checkpoints/{checkpoint}.ckpt' of the reference (main.py:211, 288): the prefix of the bundle files."""
    return "{}/{}.ckpt".format(directory.rstrip("/"), params.checkpoint)


def save(path, state):
    """saver.save(sess, path): state is {tf_variable_name: array}. Returns `path` like Saver.save does."""
    d = os.path.dirname(path)
    if d and not os.path.exists(d):
        os.makedirs(d)
    state = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
    if path.endswith(".npz"):
        np.savez(path, **state)  # names contain '/', which np.savez keeps verbatim as archive member names
        return path
    tf_bundle.write_bundle(path, state)
    tf_bundle.update_checkpoint_state(d or ".", path)
    return path


def load(path, names=None):
    """{tf_variable_name: array} of a bundle prefix (or .npz archive); `names` restricts what is read from a bundle."""
    if path.endswith(".npz"):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        with np.load(path) as z:
            return {k: z[k] for k in z.files if names is None or k in names}
    return tf_bundle.read_bundle(path, names)


def restore(engine, path, strict=True):
    """saver.restore: every variable of the engine must be present with its shape (TF raises otherwise)."""
    if path.endswith(".npz"):
        state = load(path)
    else:  # a reference-written bundle also carries cnn/ variables the engine may not hold: read only what it needs
        _, entries = tf_bundle.list_bundle(path)
        state = load(path, [n for n, _, _ in engine.variables() if n in entries])
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
