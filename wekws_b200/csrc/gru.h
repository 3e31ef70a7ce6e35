// Kernel argument block of the fused GRU kernel (gru.cu).
#pragma once
#include <cuda_runtime.h>

namespace wekws {

struct GruArgs {
  const float* feats;      // (B, T, idim)
  const float* in_cache;   // (L, B, H) or nullptr (== zeros)
  float* out;              // (B, T, odim)
  float* out_cache;        // (L, B, H)
  const float* vec;        // packed weights (see model_host.cu: pack_gru)
  int B, T, L, H, idim, odim, act, has_cmvn;
  int v_mean, v_istd, v_wp, v_bp, v_layers, v_layer_stride, v_wc, v_bc;
  int n_tiles;
};

int gru_launch(const GruArgs& a, cudaStream_t st);

}  // namespace wekws
