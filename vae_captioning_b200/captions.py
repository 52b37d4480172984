"""Caption tokeniser and vocabulary (reference: utils/captions.py:37-40, 65-126).

Token ids define the rows of the embedding tables and the columns of the vocab projection, so the id
assignment must match the reference exactly for checkpoints to be interchangeable: words sorted by
(-count, word), kept when count >= keep_words (or '<UNK>'), ids from 1, '<PAD>' = 0."""
import collections
import pickle
import re

_SPLIT = re.compile(r"\W+")
_SPECIAL = ("<EOS>", "<BOS>", "<PAD>")


def tokenize_caption(caption):
    return ["<BOS>"] + [w for w in _SPLIT.split(caption.lower()) if w] + ["<EOS>"]


class Dictionary(object):
    def __init__(self, caption_dict, keep_words, save_path="./pickles/capt_vocab.pickle"):
        counts = collections.Counter()
        for caps in caption_dict.values():
            for cap in caps:
                counts.update(w if w in _SPECIAL else w.lower() for w in cap)
        counts["<UNK>"] += 1
        ranked = sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))
        words = [w for w, n in ranked if n >= keep_words or w == "<UNK>"]
        self._word2idx = {w: i for i, w in enumerate(words, start=1)}
        self._idx2word = {i: w for w, i in self._word2idx.items()}
        self._idx2word[0] = "<PAD>"
        self._word2idx["<PAD>"] = 0
        if save_path:  # the reference pickles the raw caption dict next to the checkpoints (captions.py:122-125)
            with open(save_path, "wb") as f:
                pickle.dump(caption_dict, f)

    @property
    def vocab_size(self):
        return len(self._idx2word)

    @property
    def word2idx(self):
        return self._word2idx

    @property
    def idx2word(self):
        return self._idx2word

    def seq2dx(self, sentence):
        return [self._word2idx[w] for w in sentence]

    def __len__(self):
        return len(self._idx2word)
