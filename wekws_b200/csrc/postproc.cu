// On-device detection statistics for max-pooling keyword models: what the reference computes on the
// host from the text score file -- wekws/bin/score.py:128-137 writes every frame's posterior with '{:.6f}',
// wekws/bin/compute_det.py:76-105 then takes, per threshold, (a) max(score_list) < threshold over keyword
// utterances (false rejects) and (b) a left-to-right scan over filler utterances that counts a trigger whenever
// score >= threshold and then skips `window_shift` frames (false alarms).
//
// Bit-exact with that pipeline: a float posterior x in [0, 1] times 1e6 is exact in double (24 x 20 significant
// bits), so rint(x * 1e6) / 1e6 is precisely the double Python parses back from '{:.6f}'.format(x), and the
// thresholds are supplied by the host as the doubles the reference accumulates (threshold += step).
#include <stdint.h>

#include "common.cuh"

namespace wekws {
namespace {

__device__ __forceinline__ double text_round6(float x) { return rint((double)x * 1e6) / 1e6; }

// one thread per (stream b, keyword k, threshold i); a warp covers consecutive thresholds of one (b, k), so the
// posteriors it scans are broadcast loads
__global__ void det_stats_kernel(const float* __restrict__ post, const int32_t* __restrict__ lens, long long B,
                                 long long T, int K, const double* __restrict__ thr, int nthr, int window_shift,
                                 double* __restrict__ max_score, int32_t* __restrict__ triggers) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = B * K * (long long)nthr;
  if (idx >= total) return;
  const int i = (int)(idx % nthr);
  const long long bk = idx / nthr;
  const int k = (int)(bk % K);
  const long long b = bk / K;
  long long n = lens ? (long long)lens[b] : T;
  n = n < 0 ? 0 : (n > T ? T : n);
  const float* p = post + (b * T) * K + k;
  const double th = thr[i];
  int count = 0;
  long long t = 0;
  while (t < n) {                                   // compute_det.py:91-97
    if (text_round6(p[t * K]) >= th) { ++count; t += window_shift; }
    else ++t;
  }
  triggers[idx] = count;
  if (i == 0) {                                     // compute_det.py:83: max(score_list) (as rounded text values)
    float m = -INFINITY;
    for (long long u = 0; u < n; ++u) m = fmaxf(m, p[u * K]);
    max_score[bk] = n > 0 ? text_round6(m) : (double)m;     // the double Python parses back (compared in double)
  }
}

}  // namespace
}  // namespace wekws

using namespace wekws;

extern "C" int wekws_det_stats(const float* d_post, const int32_t* d_lens, int64_t B, int64_t T, int K,
                               const double* d_thresholds, int nthr, int window_shift, double* d_max_score,
                               int32_t* d_triggers, void* stream) {
  WEKWS_REQUIRE(B >= 0 && T >= 0 && K >= 1 && nthr >= 1, "wekws_det_stats: bad sizes");
  WEKWS_REQUIRE(window_shift >= 1, "wekws_det_stats: window_shift must be >= 1 (got %d)", window_shift);
  if (B == 0) return WEKWS_OK;
  WEKWS_REQUIRE((d_post || T == 0) && d_thresholds && d_max_score && d_triggers, "wekws_det_stats: null argument");
  const long long total = (long long)B * K * nthr;
  const int nt = 128;
  det_stats_kernel<<<(unsigned)((total + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
      d_post, d_lens, B, T, K, d_thresholds, nthr, window_shift, d_max_score, d_triggers);
  return check_launch("det_stats_kernel");
}
