// Beam bookkeeping of Decoder.beam_search (reference: vae_model/decoder.py:237-319) with the exact container
// semantics of utils/top_n.py: TopN is a size-n binary min-heap driven by heapq.heappush / heapq.heappushpop,
// partial beams are processed in raw heap-array order, and the final list is sorted descending and stably.
// The heap routines below follow CPython's heapq algorithm (sift directions, strict '<' comparisons) so ties
// resolve exactly as in the reference. Compiled for host and device: the CUDA decode kernel runs one thread per
// image; vc_beam_search_host runs the same code against a callback (CPU parity tests vs the golden fixtures).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

namespace vc {

constexpr int kMaxBeam = 16;

struct BeamEntry {
  float score;      // TopN key (Beam.score)
  float logprob;    // Beam.logprob
  int node;         // history node of the last token (valid for entries of the current partial list)
  int parent_node;  // history node of the prefix
  int token;        // last token
  int len;          // len(sentence), <BOS> included
  int src_row;      // which of this iteration's rows produced the state (Beam.state)
};

struct BeamHeap {
  int n;
  BeamEntry e[kMaxBeam];
};

// heapq._siftdown
VC_HD void heap_siftdown(BeamEntry* h, int startpos, int pos) {
  const BeamEntry item = h[pos];
  while (pos > startpos) {
    const int parent = (pos - 1) >> 1;
    if (item.score < h[parent].score) {
      h[pos] = h[parent];
      pos = parent;
      continue;
    }
    break;
  }
  h[pos] = item;
}

// heapq._siftup
VC_HD void heap_siftup(BeamEntry* h, int n, int pos) {
  const int startpos = pos;
  const BeamEntry item = h[pos];
  int child = 2 * pos + 1;
  while (child < n) {
    const int right = child + 1;
    if (right < n && !(h[child].score < h[right].score)) child = right;
    h[pos] = h[child];
    pos = child;
    child = 2 * pos + 1;
  }
  h[pos] = item;
  heap_siftdown(h, startpos, pos);
}

// TopN.push (utils/top_n.py:15-21)
VC_HD void topn_push(BeamHeap* t, int cap, const BeamEntry& x) {
  if (t->n < cap) {
    t->e[t->n] = x;
    t->n += 1;
    heap_siftdown(t->e, 0, t->n - 1);
  } else if (t->n > 0 && t->e[0].score < x.score) {  // heapq.heappushpop
    t->e[0] = x;
    heap_siftup(t->e, t->n, 0);
  }
}

// list.sort(reverse=True): descending, equal keys keep their original order
VC_HD void stable_sort_desc(BeamEntry* e, int n) {
  for (int i = 1; i < n; ++i) {
    const BeamEntry x = e[i];
    int j = i - 1;
    while (j >= 0 && e[j].score < x.score) {
      e[j + 1] = e[j];
      --j;
    }
    e[j + 1] = x;
  }
}

struct BeamImage {
  BeamHeap partial, complete;
  int n_nodes;
  int done;
};

VC_HD void beam_image_init(BeamImage* im, int2* nodes, int bos) {
  im->partial.n = 1;
  BeamEntry& b = im->partial.e[0];
  b.score = 0.f; b.logprob = 0.f; b.node = 0; b.parent_node = -1; b.token = bos; b.len = 1; b.src_row = 0;
  im->complete.n = 0;
  im->n_nodes = 1;
  im->done = 0;
  nodes[0].x = bos;
  nodes[0].y = -1;
}

// One iteration of the loop at decoder.py:248-299 for one image. cand_idx / cand_p: [beam rows, beam] the top-`beam`
// words of each processed row in descending probability (ties: lower index first, list.sort stability) and their
// softmax probabilities. Writes, for every slot j of the new partial list, the row whose state it inherits and the
// token to feed next. Returns the new partial count.
VC_HD int beam_image_update(BeamImage* im, int2* nodes, int max_nodes, const int* cand_idx, const float* cand_p, int beam,
                            int eos, float len_norm, int* slot_src_row, int* slot_token) {
  if (im->done) return 0;
  BeamEntry plist[kMaxBeam];
  const int np = im->partial.n;
  for (int i = 0; i < np; ++i) plist[i] = im->partial.e[i];
  im->partial.n = 0;  // partial_captions.reset()
  for (int i = 0; i < np; ++i) {
    const BeamEntry& c = plist[i];
    for (int j = 0; j < beam; ++j) {
      const float p = cand_p[i * beam + j];
      if (p < 1e-12f) continue;  // "Avoid log(0)"
      BeamEntry b;
      b.token = cand_idx[i * beam + j];
      b.parent_node = c.node;
      b.node = -1;
      b.len = c.len + 1;
      b.logprob = c.logprob + logf(p);
      b.score = b.logprob;
      b.src_row = i;
      if (b.token == eos) {
        if (len_norm > 0.f) b.score = b.score / powf((float)b.len, len_norm);
        topn_push(&im->complete, beam, b);
      } else {
        topn_push(&im->partial, beam, b);
      }
    }
  }
  if (im->partial.n == 0) {
    im->done = 1;
    return 0;
  }
  for (int j = 0; j < im->partial.n; ++j) {
    BeamEntry& b = im->partial.e[j];
    int id = im->n_nodes;
    if (id < max_nodes) {
      nodes[id].x = b.token;
      nodes[id].y = b.parent_node;
      im->n_nodes = id + 1;
    } else {
      id = max_nodes - 1;  // cannot happen: max_nodes >= 1 + beam * iterations
    }
    b.node = id;
    slot_src_row[j] = b.src_row;
    slot_token[j] = b.token;
  }
  return im->partial.n;
}

// decoder.py:300-319: fall back to the partial list when nothing completed, sort, emit sentences.
// out_tokens [beam, max_len] (zero padded), out_len [beam], out_score [beam]; returns the number of beams.
VC_HD int beam_image_finish(BeamImage* im, const int2* nodes, int beam, int max_len, int* out_tokens, int* out_len,
                            float* out_score) {
  BeamHeap* src = im->complete.n > 0 ? &im->complete : &im->partial;
  stable_sort_desc(src->e, src->n);
  for (int b = 0; b < beam; ++b) {
    out_len[b] = 0;
    out_score[b] = 0.f;
    for (int t = 0; t < max_len; ++t) out_tokens[b * max_len + t] = 0;
  }
  for (int b = 0; b < src->n; ++b) {
    const BeamEntry& e = src->e[b];
    int len = e.len > max_len ? max_len : e.len;
    out_len[b] = len;
    out_score[b] = e.score;
    int pos = len - 1;
    if (e.node >= 0) {  // a partial entry: its own node is the last token
      int nd = e.node;
      while (nd >= 0 && pos >= 0) { out_tokens[b * max_len + pos] = nodes[nd].x; nd = nodes[nd].y; --pos; }
    } else {
      out_tokens[b * max_len + pos] = e.token;
      --pos;
      int nd = e.parent_node;
      while (nd >= 0 && pos >= 0) { out_tokens[b * max_len + pos] = nodes[nd].x; nd = nodes[nd].y; --pos; }
    }
  }
  return src->n;
}

}  // namespace vc
