"""`Parameters` -- the configuration surface of the reference (utils/parameters.py:1-164).

Same attribute names, defaults and command-line flags, so scripts and pickled parameter objects
written for the reference keep working. The table below is data, not a transcription: every row is
(attribute, default, flag, kind). Rows without a flag are class-attribute-only knobs in the
reference as well.
"""
import argparse
import os

# (attribute, default, cli flag or None, kind) -- kind: float/int/str/flag
_FIELDS = [
    ("latent_size", 150, "--latent", int),
    ("num_clusters", 90, None, int),
    ("num_epochs", 20, "--epochs", int),
    ("learning_rate", 0.0005, "--lr", float),
    ("num_captions", 5, None, int),
    ("batch_size", 32, "--bs", int),
    ("cnn_feature_size", 4096, None, int),
    ("temperature", 1.0, "--temperature", float),
    ("sample_gen", "beam_search", "--sample_gen", str),
    ("beam_size", 10, None, int),
    ("encoder_rnn_layers", 1, None, int),
    ("encoder_hidden", 512, "--enc_hid", int),
    ("std", 0.1, "--std", float),
    ("decoder_hidden", 512, "--dec_hid", int),
    ("decoder_rnn_layers", 1, None, int),
    ("dec_keep_rate", 1.0, "--dec_drop", float),
    ("embed_size", 256, "--embed_dim", int),
    ("gen_max_len", 30, None, int),
    ("gen_z_samples", 100, "--gen_z_samples", int),
    ("ann_param", 0, "--ann_param", float),
    ("dec_lstm_drop", 1.0, "--dec_lstm_drop", float),
    ("optimizer", "Adam", "--optimizer", ("SGD", "Adam", "Momentum")),
    ("lstm_clip_by_norm", 5.0, None, float),
    ("restore", False, "--restore", "flag"),
    ("LOG_DIR", "./model_logs/", None, str),
    ("save_params", 0, "--save_params", "flag"),
    ("no_encoder", False, "--no_encoder", "flag"),
    ("vocab_size", None, None, int),
    ("coco_dir", "/home/luoyy16/datasets-large/mscoco/coco/", "--coco_dir", str),
    ("use_hdf5", True, None, bool),
    ("fine_tune", False, "--fine_tune", "flag"),
    ("fine_tune_top", True, None, bool),
    ("fine_tune_fe", True, None, bool),
    ("cnn_lr", 0.00001, None, float),
    ("cnn_optimizer", "Adam", None, str),
    ("cnn_dropout", 0.5, None, float),
    ("weight_decay", 0.00004, None, float),
    ("gen_name", "00", "--gen_name", str),
    ("checkpoint", "last_run", "--checkpoint", str),
    ("num_epochs_per_decay", 5, None, int),
    ("use_c_v", False, "--c_v", "flag"),
    ("gen_val_captions", 4000, None, int),
    ("keep_words", 3, None, int),
    ("cap_max_length", 100, None, int),
    ("prior", "Normal", "--prior", ("GMM", "AG", "Normal")),
    ("max_checkpoints_to_keep", 5, None, int),
    ("mode", "training", "--mode", ("training", "inference")),
    ("num_ex_per_epoch", 150000, None, int),
    ("image_net_weights_path", "./utils/vgg16_weights.npz", None, str),
    ("logging", False, None, bool),
]


class Parameters(object):
    """Class attributes are the defaults; `parse_args()` overrides them from the command line."""

    def parse_args(self, argv=None):
        ap = argparse.ArgumentParser(description="CVAE captioning parameters (reference-compatible flags)")
        for attr, default, flag, kind in _FIELDS:
            if flag is None:
                continue
            if kind == "flag":
                ap.add_argument(flag, dest=attr, action="store_true")
            elif isinstance(kind, tuple):
                ap.add_argument(flag, dest=attr, default=getattr(self, attr), choices=list(kind))
            else:
                ap.add_argument(flag, dest=attr, default=getattr(self, attr))
        # the reference makes --gpu mandatory in effect (os.environ[...] = None raises, parameters.py:163-164)
        ap.add_argument("--gpu", dest="gpu", default=None, help="GPU index (CUDA_VISIBLE_DEVICES)")
        ns = ap.parse_args(argv)
        for attr, default, flag, kind in _FIELDS:
            if flag is None:
                continue
            val = getattr(ns, attr)
            if kind in (int, float):
                val = kind(val)
            setattr(self, attr, val)
        self.hdf5_file = self.coco_dir + Parameters.hdf5_file.split("/")[-1]
        if ns.gpu is None:
            raise TypeError("--gpu must be given (the reference sets CUDA_VISIBLE_DEVICES from it)")
        os.environ["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
        os.environ["CUDA_VISIBLE_DEVICES"] = str(ns.gpu)
        return self


for _attr, _default, _flag, _kind in _FIELDS:
    setattr(Parameters, _attr, _default)
Parameters.hdf5_file = Parameters.coco_dir + "train_val.hdf5"
