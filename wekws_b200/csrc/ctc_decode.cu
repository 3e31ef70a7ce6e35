// CTC prefix beam search + keyword detection on the device (CTC models): what the reference does per
// utterance / per frame in pure Python --
//   wekws/model/loss.py:206-312 ctc_prefix_beam_search (whole utterance, wekws/bin/score_ctc.py:198-200) and its
//   streaming twin wekws/bin/stream_kws_ctc.py:124-215 (one frame per call, hypotheses carried), followed by the
//   keyword look-up of score_ctc.py:201-220 / stream_kws_ctc.py:411-434 (is_sublist, sqrt of the product of the
//   token probabilities).
// Bit-exact restatement, including the parts that only exist because of Python object semantics:
//   * probabilities are float32 values promoted to double, every update is the same sequence of double multiplies and
//     adds (no FMA contraction), `math.isclose(p, 0.0, abs_tol=1e-6)` is |p| <= 1e-6;
//   * path nodes are dict OBJECTS shared between hypotheses by the shallow `cur_nodes.copy()`: `nodes[-1]['prob'] = ps`
//     (loss.py:273-275) is visible through every list that holds the same dict.  Nodes therefore live in a pool and the
//     hypotheses hold node ids; "pop + append" (loss.py:296-297) allocates a fresh node;
//   * `next_hyps` is a dict in insertion order and `sorted(..., reverse=True)` is stable: ties keep insertion order;
//   * is_sublist never tests the last possible offset when the prefix is longer than the keyword (score_ctc.py:95).
// One warp per utterance: all lanes scan the frame's V probabilities for the top score_beam entries, lane 0 runs the
// (tiny, inherently sequential) hypothesis update in shared memory.  Utterances are independent -> B warps.
#include <math.h>

#include "common.cuh"
#include "ctc_decode.h"

namespace wekws {
namespace {

constexpr int ML = WEKWS_CTC_MAX_PREFIX;      // longest prefix / node list kept (overflow is flagged)
constexpr int PBM = WEKWS_CTC_MAX_PATH_BEAM;  // path_beam_size limit
constexpr int SBM = WEKWS_CTC_MAX_SCORE_BEAM; // score_beam_size limit
constexpr int NEXTM = PBM * (SBM + 1);        // keys next_hyps can hold: old prefixes + one extension per token
constexpr int POOLM = PBM * ML + PBM * SBM + 8;

struct Hyp {
  double pb, pnb;
  int32_t len, nlen;          // prefix length, node-list length (equal except transiently empty lists)
  int16_t tok[ML];
  int16_t node[ML];
};

struct Work {                 // one utterance's decoder state (shared memory)
  Hyp cur[PBM];
  Hyp next[NEXTM];
  int32_t nframe[POOLM];      // node pool: frame, prob, token
  float nprob[POOLM];
  int16_t ntok[POOLM];
  int16_t remap[POOLM];
  int32_t ncur, npool, overflow;
  uint8_t order[NEXTM];
};

__device__ __forceinline__ bool close0(double p) { return fabs(p) <= 1e-6; }   // math.isclose(p, 0.0, abs_tol=1e-6)

// find the entry with this prefix (tok[0..len) [+ extra]) or insert an empty one -- defaultdict((0.0, 0.0, []))
__device__ int find_or_insert(Work& w, int& nnext, const int16_t* tok, int len, int extra /* -1 = none */) {
  const int L = len + (extra >= 0 ? 1 : 0);
  for (int e = 0; e < nnext; ++e) {
    const Hyp& h = w.next[e];
    if (h.len != L) continue;
    bool same = true;
    for (int i = 0; i < len && same; ++i) same = h.tok[i] == tok[i];
    if (same && extra >= 0) same = h.tok[len] == (int16_t)extra;
    if (same) return e;
  }
  if (nnext >= NEXTM) return -1;
  Hyp& h = w.next[nnext];
  h.pb = 0.0; h.pnb = 0.0; h.len = L; h.nlen = 0;
  for (int i = 0; i < len; ++i) h.tok[i] = tok[i];
  if (extra >= 0) h.tok[len] = (int16_t)extra;
  return nnext++;
}

__device__ __forceinline__ void copy_nodes(Hyp& dst, const Hyp& src) {            // nodes = cur_nodes.copy()
  dst.nlen = src.nlen;
  for (int i = 0; i < src.nlen; ++i) dst.node[i] = src.node[i];
}

__device__ __forceinline__ int new_node(Work& w, int tok, int frame, float prob) {
  if (w.npool >= POOLM) { w.overflow = 1; return POOLM - 1; }
  const int id = w.npool++;
  w.ntok[id] = (int16_t)tok; w.nframe[id] = frame; w.nprob[id] = prob;
  return id;
}

// one frame of loss.py:229-306 / stream_kws_ctc.py:140-213 for the filtered tokens s[0..ns) with probabilities ps[]
__device__ void advance(Work& w, int t, const int* s_idx, const float* s_prob, int ns, int path_beam) {
  int nnext = 0;
  for (int k = 0; k < ns; ++k) {
    const int s = s_idx[k];
    const float psf = s_prob[k];
    const double ps = (double)psf;
    for (int hi = 0; hi < w.ncur; ++hi) {
      const Hyp& c = w.cur[hi];
      const int last = c.len > 0 ? c.tok[c.len - 1] : -1;
      const double pb = c.pb, pnb = c.pnb;
      if (s == 0) {                                               // blank
        const int e = find_or_insert(w, nnext, c.tok, c.len, -1);
        if (e < 0) { w.overflow = 1; continue; }
        Hyp& n = w.next[e];
        n.pb = __dadd_rn(__dadd_rn(n.pb, __dmul_rn(pb, ps)), __dmul_rn(pnb, ps));
        copy_nodes(n, c);
      } else if (s == last) {
        if (!close0(pnb)) {                                       // *ss -> *s
          const int e = find_or_insert(w, nnext, c.tok, c.len, -1);
          if (e < 0) { w.overflow = 1; continue; }
          Hyp& n = w.next[e];
          n.pnb = __dadd_rn(n.pnb, __dmul_rn(pnb, ps));
          copy_nodes(n, c);
          const int id = n.node[n.nlen - 1];
          if (psf > w.nprob[id]) { w.nprob[id] = psf; w.nframe[id] = t; }   // the shared dict is updated in place
        }
        if (!close0(pb)) {                                        // *s-s -> *ss
          if (c.len >= ML) { w.overflow = 1; continue; }
          const int e = find_or_insert(w, nnext, c.tok, c.len, s);
          if (e < 0) { w.overflow = 1; continue; }
          Hyp& n = w.next[e];
          n.pnb = __dadd_rn(n.pnb, __dmul_rn(pb, ps));
          copy_nodes(n, c);
          n.node[n.nlen++] = (int16_t)new_node(w, s, t, psf);
        }
      } else {
        if (c.len >= ML) { w.overflow = 1; continue; }
        const int e = find_or_insert(w, nnext, c.tok, c.len, s);
        if (e < 0) { w.overflow = 1; continue; }
        Hyp& n = w.next[e];
        if (n.nlen > 0) {
          if (psf > w.nprob[n.node[n.nlen - 1]]) n.node[n.nlen - 1] = (int16_t)new_node(w, s, t, psf);   // pop + append
        } else {
          copy_nodes(n, c);
          n.node[n.nlen++] = (int16_t)new_node(w, s, t, psf);
        }
        n.pnb = __dadd_rn(__dadd_rn(n.pnb, __dmul_rn(pb, ps)), __dmul_rn(pnb, ps));
      }
    }
  }
  // stable sort by pb + pnb, descending (insertion sort on an index array keeps ties in insertion order)
  for (int e = 0; e < nnext; ++e) {
    const double key = __dadd_rn(w.next[e].pb, w.next[e].pnb);
    int p = e;
    while (p > 0) {
      const Hyp& o = w.next[w.order[p - 1]];
      if (__dadd_rn(o.pb, o.pnb) >= key) break;
      w.order[p] = w.order[p - 1];
      --p;
    }
    w.order[p] = (uint8_t)e;
  }
  const int keep = nnext < path_beam ? nnext : path_beam;
  // garbage-collect the node pool: keep the nodes the surviving hypotheses reference, ids stay in ascending order
  for (int i = 0; i < w.npool; ++i) w.remap[i] = 0;
  for (int r = 0; r < keep; ++r) {
    const Hyp& h = w.next[w.order[r]];
    for (int i = 0; i < h.nlen; ++i) w.remap[h.node[i]] = 1;
  }
  int live = 0;
  for (int i = 0; i < w.npool; ++i) {
    if (w.remap[i]) {
      w.ntok[live] = w.ntok[i]; w.nframe[live] = w.nframe[i]; w.nprob[live] = w.nprob[i];
      w.remap[i] = (int16_t)live++;
    }
  }
  w.npool = live;
  for (int r = 0; r < keep; ++r) {
    const Hyp& h = w.next[w.order[r]];
    Hyp& d = w.cur[r];
    d.pb = h.pb; d.pnb = h.pnb; d.len = h.len; d.nlen = h.nlen;
    for (int i = 0; i < h.len; ++i) d.tok[i] = h.tok[i];
    for (int i = 0; i < h.nlen; ++i) d.node[i] = w.remap[h.node[i]];
  }
  w.ncur = keep;
}

__global__ void __launch_bounds__(32) ctc_prefix_beam_kernel(const CtcArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  Work& w = *reinterpret_cast<Work*>(smem);
  uint32_t* allow = reinterpret_cast<uint32_t*>(smem + sizeof(Work));      // keyword-token bitmap over V (optional)
  const int lane = threadIdx.x;
  const long long b = blockIdx.x;
  const int V = a.V, SB = a.score_beam;
  long long n = a.lens ? (long long)a.lens[b] : a.T;
  n = n < 0 ? 0 : (n > a.T ? a.T : n);

  const int nwords = (V + 31) / 32;
  for (int i = lane; i < nwords; i += 32) allow[i] = a.n_allowed > 0 ? 0u : 0xffffffffu;
  __syncwarp();
  if (lane == 0) {
    for (int i = 0; i < a.n_allowed; ++i) {
      const int tkn = a.allowed[i];
      if (tkn >= 0 && tkn < V) allow[tkn >> 5] |= 1u << (tkn & 31);
    }
    // hypotheses: carried state or the initial [(tuple(), (1.0, 0.0, []))]
    Work* st = a.state ? reinterpret_cast<Work*>(a.state + (size_t)b * sizeof(Work)) : nullptr;
    if (st && !a.reset_state) {
      w.ncur = st->ncur; w.npool = st->npool; w.overflow = st->overflow;
      for (int h = 0; h < st->ncur; ++h) w.cur[h] = st->cur[h];
      for (int i = 0; i < st->npool; ++i) { w.ntok[i] = st->ntok[i]; w.nframe[i] = st->nframe[i]; w.nprob[i] = st->nprob[i]; }
    } else {
      w.ncur = 1; w.npool = 0; w.overflow = 0;
      w.cur[0].pb = 1.0; w.cur[0].pnb = 0.0; w.cur[0].len = 0; w.cur[0].nlen = 0;
    }
  }
  __syncwarp();

  const float* P = a.probs + b * a.T * (long long)V;
  for (long long t = 0; t < n; ++t) {
    const float* p = P + t * V;
    // ---- probs.topk(score_beam): per-lane candidates, then SB rounds of warp arg-max (ties: lower index first)
    float bv[SBM];
    int bi[SBM];
#pragma unroll
    for (int k = 0; k < SBM; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
    for (int i = lane; i < V; i += 32) {
      const float v = __ldg(p + i);
      if (v > bv[SBM - 1]) {                     // strictly greater: an equal later index never displaces an earlier one
        bv[SBM - 1] = v; bi[SBM - 1] = i;
#pragma unroll
        for (int k = SBM - 1; k > 0; --k) {
          if (bv[k] > bv[k - 1]) {
            const float tv = bv[k]; bv[k] = bv[k - 1]; bv[k - 1] = tv;
            const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
          }
        }
      }
    }
    int s_idx[SBM];
    float s_prob[SBM];
    int ns = 0;
    for (int k = 0; k < SB; ++k) {
      float mv = bv[0];
      int mi = bi[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
      }
      if (bi[0] == mi && mi != 0x7fffffff) {     // the winning lane pops its head
#pragma unroll
        for (int q = 0; q < SBM - 1; ++q) { bv[q] = bv[q + 1]; bi[q] = bi[q + 1]; }
        bv[SBM - 1] = -INFINITY; bi[SBM - 1] = 0x7fffffff;
      }
      // filter: prob > 0.05 (Python float compare of the float32 value) and token in the keyword set
      if (mi != 0x7fffffff && (double)mv > 0.05 && ((allow[mi >> 5] >> (mi & 31)) & 1u)) {
        s_idx[ns] = mi; s_prob[ns] = mv; ++ns;
      }
    }
    if (ns == 0) continue;                       // loss.py:254-255: the frame is skipped entirely
    if (lane == 0) advance(w, (int)(a.frame_offset + t * a.frame_stride), s_idx, s_prob, ns, a.path_beam);
    __syncwarp();
  }

  if (lane == 0) {
    // hyps = [(prefix, pb + pnb, nodes)]
    a.nhyp[b] = w.ncur;
    a.overflow[b] = w.overflow;
    for (int h = 0; h < a.path_beam; ++h) {
      const long long o = b * a.path_beam + h;
      if (h < w.ncur) {
        const Hyp& c = w.cur[h];
        a.hyp_len[o] = c.len;
        a.hyp_score[o] = __dadd_rn(c.pb, c.pnb);
        for (int i = 0; i < ML; ++i) {
          a.hyp_tokens[o * ML + i] = i < c.len ? c.tok[i] : -1;
          a.node_frame[o * ML + i] = i < c.nlen ? w.nframe[c.node[i]] : -1;
          a.node_prob[o * ML + i] = i < c.nlen ? w.nprob[c.node[i]] : 0.f;
        }
      } else {
        a.hyp_len[o] = -1;
        a.hyp_score[o] = 0.0;
      }
    }
    if (a.state) {
      Work* st = reinterpret_cast<Work*>(a.state + (size_t)b * sizeof(Work));
      st->ncur = w.ncur; st->npool = w.npool; st->overflow = w.overflow;
      for (int h = 0; h < w.ncur; ++h) st->cur[h] = w.cur[h];
      for (int i = 0; i < w.npool; ++i) { st->ntok[i] = w.ntok[i]; st->nframe[i] = w.nframe[i]; st->nprob[i] = w.nprob[i]; }
    }
  }
}

// score_ctc.py:88-103 (identical copy in stream_kws_ctc.py:105-120), quirk included: for a longer main list the loop
// runs over range(len(main) - len(check)) and never tests the last offset
__device__ int is_sublist(const int32_t* main_list, int nm, const int32_t* check, int nc) {
  if (nm < nc) return -1;
  if (nm == nc) {
    for (int i = 0; i < nc; ++i)
      if (main_list[i] != check[i]) return -1;
    return 0;
  }
  for (int i = 0; i < nm - nc; ++i) {
    if (main_list[i] == check[0]) {
      int j = 0;
      while (j < nc && main_list[i + j] == check[j]) ++j;
      if (j == nc) return i;
    }
  }
  return -1;
}

// score_ctc.py:201-220: first hypothesis (in beam order) containing a keyword (in keyword order)
__global__ void ctc_keyword_hit_kernel(const int32_t* __restrict__ nhyp, const int32_t* __restrict__ hyp_len,
                                       const int32_t* __restrict__ hyp_tokens, const int32_t* __restrict__ node_frame,
                                       const float* __restrict__ node_prob, long long B, int path_beam,
                                       const int32_t* __restrict__ kw_tokens, const int32_t* __restrict__ kw_off, int nkw,
                                       int32_t* __restrict__ hit, double* __restrict__ hit_score,
                                       int32_t* __restrict__ start, int32_t* __restrict__ end) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int found = -1, st = 0, en = 0;
  double score = 1.0;
  for (int h = 0; h < nhyp[b] && found < 0; ++h) {
    const long long o = b * path_beam + h;
    const int32_t* pre = hyp_tokens + o * ML;
    for (int k = 0; k < nkw; ++k) {
      const int32_t* lab = kw_tokens + kw_off[k];
      const int nl = kw_off[k + 1] - kw_off[k];
      const int off = is_sublist(pre, hyp_len[o], lab, nl);
      if (off != -1) {
        found = k;
        st = node_frame[o * ML + off];
        en = node_frame[o * ML + off + nl - 1];
        for (int i = off; i < off + nl; ++i) score = __dmul_rn(score, (double)node_prob[o * ML + i]);
        break;
      }
    }
    if (found >= 0) score = sqrt(score);
  }
  hit[b] = found; hit_score[b] = score; start[b] = st; end[b] = en;
}

}  // namespace

size_t ctc_state_bytes() { return sizeof(Work); }

int ctc_launch(const CtcArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(Work) + (size_t)((a.V + 31) / 32) * 4;
  WEKWS_REQUIRE(smem <= 227 * 1024, "ctc decode: vocabulary %d too large for the shared-memory token bitmap", a.V);
  static size_t attr_bytes[64] = {0};                  // per device: the largest dynamic size opted into so far
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && attr_bytes[dev] < smem) {
    WEKWS_CUDA_OK(cudaFuncSetAttribute(ctc_prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes[dev] = smem;
  }
  ctc_prefix_beam_kernel<<<(unsigned)a.B, 32, smem, st>>>(a);
  return check_launch("ctc_prefix_beam_kernel");
}

int ctc_hit_launch(const int32_t* nhyp, const int32_t* hyp_len, const int32_t* hyp_tokens, const int32_t* node_frame,
                   const float* node_prob, long long B, int path_beam, const int32_t* kw_tokens, const int32_t* kw_off,
                   int nkw, int32_t* hit, double* hit_score, int32_t* start, int32_t* end, cudaStream_t st) {
  const int nt = 64;
  ctc_keyword_hit_kernel<<<(unsigned)((B + nt - 1) / nt), nt, 0, st>>>(nhyp, hyp_len, hyp_tokens, node_frame, node_prob, B,
                                                                     path_beam, kw_tokens, kw_off, nkw, hit, hit_score,
                                                                     start, end);
  return check_launch("ctc_keyword_hit_kernel");
}

}  // namespace wekws

using namespace wekws;

extern "C" int64_t wekws_ctc_state_bytes(void) { return (int64_t)ctc_state_bytes(); }

extern "C" int wekws_ctc_prefix_beam_search(const float* d_probs, const int32_t* d_lens, int64_t B, int64_t T, int V,
                                            const int32_t* d_keyword_tokens, int n_keyword_tokens, int score_beam_size,
                                            int path_beam_size, int64_t frame_offset, int frame_stride, void* d_state,
                                            int reset_state, int32_t* d_nhyp, int32_t* d_hyp_len, int32_t* d_hyp_tokens,
                                            double* d_hyp_score, int32_t* d_node_frame, float* d_node_prob,
                                            int32_t* d_overflow, void* stream) {
  WEKWS_REQUIRE(B >= 0 && T >= 0 && V >= 1 && V <= 32767, "wekws_ctc_prefix_beam_search: bad sizes (vocabulary <= 32767)");
  WEKWS_REQUIRE(score_beam_size >= 1 && score_beam_size <= WEKWS_CTC_MAX_SCORE_BEAM && score_beam_size <= V,
                "score_beam_size %d out of range (1..%d)", score_beam_size, WEKWS_CTC_MAX_SCORE_BEAM);
  WEKWS_REQUIRE(path_beam_size >= 1 && path_beam_size <= WEKWS_CTC_MAX_PATH_BEAM, "path_beam_size %d out of range (1..%d)",
                path_beam_size, WEKWS_CTC_MAX_PATH_BEAM);
  WEKWS_REQUIRE(n_keyword_tokens >= 0 && (n_keyword_tokens == 0 || d_keyword_tokens), "keyword token set is null");
  WEKWS_REQUIRE(frame_stride >= 1, "frame_stride must be >= 1");
  if (B == 0) return WEKWS_OK;
  WEKWS_REQUIRE((d_probs || T == 0) && d_nhyp && d_hyp_len && d_hyp_tokens && d_hyp_score && d_node_frame && d_node_prob &&
                    d_overflow,
                "wekws_ctc_prefix_beam_search: null argument");
  CtcArgs a;
  a.probs = d_probs; a.lens = d_lens; a.B = B; a.T = T; a.V = V;
  a.allowed = d_keyword_tokens; a.n_allowed = n_keyword_tokens;
  a.score_beam = score_beam_size; a.path_beam = path_beam_size;
  a.frame_offset = frame_offset; a.frame_stride = frame_stride;
  a.state = (uint8_t*)d_state; a.reset_state = reset_state;
  a.nhyp = d_nhyp; a.overflow = d_overflow; a.hyp_len = d_hyp_len; a.hyp_tokens = d_hyp_tokens; a.hyp_score = d_hyp_score;
  a.node_frame = d_node_frame; a.node_prob = d_node_prob;
  return ctc_launch(a, (cudaStream_t)stream);
}

extern "C" int wekws_ctc_keyword_hit(const int32_t* d_nhyp, const int32_t* d_hyp_len, const int32_t* d_hyp_tokens,
                                     const int32_t* d_node_frame, const float* d_node_prob, int64_t B, int path_beam_size,
                                     const int32_t* d_kw_tokens, const int32_t* d_kw_offsets, int num_keywords,
                                     int32_t* d_hit, double* d_hit_score, int32_t* d_start, int32_t* d_end, void* stream) {
  WEKWS_REQUIRE(B >= 0 && path_beam_size >= 1 && num_keywords >= 1, "wekws_ctc_keyword_hit: bad sizes");
  if (B == 0) return WEKWS_OK;
  WEKWS_REQUIRE(d_nhyp && d_hyp_len && d_hyp_tokens && d_node_frame && d_node_prob && d_kw_tokens && d_kw_offsets && d_hit &&
                    d_hit_score && d_start && d_end,
                "wekws_ctc_keyword_hit: null argument");
  return ctc_hit_launch(d_nhyp, d_hyp_len, d_hyp_tokens, d_node_frame, d_node_prob, B, path_beam_size, d_kw_tokens,
                        d_kw_offsets, num_keywords, d_hit, d_hit_score, d_start, d_end, (cudaStream_t)stream);
}
