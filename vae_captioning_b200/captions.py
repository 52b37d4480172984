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


class Captions(object):
    """COCO caption annotations of one split (reference: utils/captions.py:5-63).

    `captions` maps file name -> list of token lists (`<BOS>` w.. `<EOS>`), in annotation order; `captions_indexed` is
    the same mapping after `index_captions` replaced the words by vocabulary ids (out-of-vocabulary -> `<UNK>`).
    The reference's length clip (`captions.py:34-36`) tests `len()` of the annotation *dict* (its key count), so it
    never triggers for COCO's three-key annotations and would raise on a longer one; captions are therefore never
    clipped here either and `max_length` is only recorded."""

    def __init__(self, captions_file, max_length=16, verbose=True):
        import json
        self._cap_json = captions_file
        self.max_length = max_length
        self.captions = collections.defaultdict(list)
        with open(captions_file) as rf:
            j = json.loads(rf.read())
        name_of = {img["id"]: img["file_name"] for img in j["images"]}
        self._fn_to_id = {img["file_name"]: img["id"] for img in j["images"]}
        for ann in j["annotations"]:
            self.captions[name_of[ann["image_id"]]].append(tokenize_caption(ann["caption"]))
        self.captions_indexed = self.captions.copy()  # shallow: index_captions rewrites the shared lists in place
        self.num_captions = len(self.captions)
        if verbose:
            print("Number of images in set", self.num_captions)

    def index_captions(self, word2idx):
        unk = word2idx["<UNK>"]
        for caps in self.captions_indexed.values():
            for i, cap in enumerate(caps):
                caps[i] = [word2idx.get(w, unk) for w in cap]

    @property
    def filename_to_imid(self):
        return self._fn_to_id
