// Kernel argument block of the CTC prefix beam search (ctc_decode.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/wekws_b200.h"

namespace wekws {

struct CtcArgs {
  const float* probs;        // (B, T, V) softmax posteriors
  const int32_t* lens;       // (B) valid frames or nullptr
  long long B, T;
  int V;
  const int32_t* allowed;    // keyword token set (n_allowed ids) or nullptr / 0 = every token
  int n_allowed;
  int score_beam, path_beam;
  long long frame_offset;    // absolute frame number of row 0 (streaming: total_frames)
  int frame_stride;          // frames per row (streaming: downsampling)
  uint8_t* state;            // B x ctc_state_bytes() carried hypotheses, or nullptr
  int reset_state;
  int32_t* nhyp;             // (B)
  int32_t* overflow;         // (B)
  int32_t* hyp_len;          // (B, path_beam)      -1 = unused slot
  int32_t* hyp_tokens;       // (B, path_beam, WEKWS_CTC_MAX_PREFIX)
  double* hyp_score;         // (B, path_beam)      pb + pnb
  int32_t* node_frame;       // (B, path_beam, WEKWS_CTC_MAX_PREFIX)
  float* node_prob;          // (B, path_beam, WEKWS_CTC_MAX_PREFIX)
};

size_t ctc_state_bytes();
int ctc_launch(const CtcArgs& a, cudaStream_t st);
int ctc_hit_launch(const int32_t* nhyp, const int32_t* hyp_len, const int32_t* hyp_tokens, const int32_t* node_frame,
                   const float* node_prob, long long B, int path_beam, const int32_t* kw_tokens, const int32_t* kw_off,
                   int nkw, int32_t* hit, double* hit_score, int32_t* start, int32_t* end, cudaStream_t st);

}  // namespace wekws
