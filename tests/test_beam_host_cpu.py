"""The product's beam bookkeeping (csrc/beam_core.h: heapq-exact TopN, <BOS> twice, length normalisation, complete /
partial separation, stable final sort) run on the HOST through vc_beam_search_host against the golden vectors that
the reference's own Decoder.beam_search produced (tests/golden/decode_loops.json). No GPU."""
import ctypes
import json
import os

import numpy as np

from oracle import decode_oracle as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float))


def run_host_beam(model, V, beam, max_len, bos=1, eos=2, len_norm=0.7):
    from vae_captioning_b200 import build, lib
    build.build()
    h = lib.load()
    h.vc_beam_search_host.argtypes = [STEP, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]
    states = []

    def step(user, token, state_in, probs_out):
        p, st = model.step(token, None if state_in < 0 else states[state_in])
        ctypes.memmove(probs_out, p.ctypes.data, V * 4)
        states.append(st)
        return len(states) - 1

    toks = np.zeros((beam, max_len), np.int32)
    lens = np.zeros((beam,), np.int32)
    scores = np.zeros((beam,), np.float32)
    n = ctypes.c_int32(0)
    cb = STEP(step)
    lib.check(h.vc_beam_search_host(cb, None, V, beam, max_len, bos, eos, len_norm, toks.ctypes.data, lens.ctypes.data,
                                    scores.ctypes.data, ctypes.byref(n)))
    return [[int(w) for w in toks[b, :lens[b]]] for b in range(n.value)], scores[:n.value]


def test_host_beam_matches_reference_golden():
    cases = [c for c in json.load(open(os.path.join(GOLD, "decode_loops.json"))) if c["mode"] == "beam_search"]
    assert len(cases) >= 12
    checked = 0
    for c in cases:
        V = c["V"]
        idx2word = {i: "w%d" % i for i in range(V)}
        idx2word.update({0: "<PAD>", 1: "<BOS>", 2: "<EOS>"})
        for i in range(c["n_img"]):
            beams, scores = run_host_beam(D.HashModel(V, c["seed"] * 1000 + i), V, c["beam"], c["max_len"])
            got = [" ".join(idx2word[w] for w in b if w not in (1, 2)) for b in beams]
            assert got[0] == c["captions"][i]["caption"], (V, c["beam"], i)
            assert got == c["ret_beams"][i]["caption"], (V, c["beam"], i)
            assert all(scores[j] >= scores[j + 1] for j in range(len(scores) - 1))
            checked += 1
    assert checked >= 36
