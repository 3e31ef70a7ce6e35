// Kernel argument block of the fused FSMN kernel (fsmn.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace wekws {

struct FsmnArgs {
  const float* feats;      // (B, T, idim), stream stride feat_bstride
  const float* in_cache;   // (B, proj, pad, L) or nullptr (start of stream == zeros)
  float* out;              // (B, T, odim), stream stride out_bstride
  float* out_cache;        // (B, proj, pad, L); may alias in_cache
  const float* w;          // packed weights (model_host.cu pack_fsmn): transposed, column-padded matrices + vectors
  int B, T, S, n_tiles;
  long long feat_bstride, out_bstride;
  int idim, aff_in, lin, proj, aff_out, odim, L, lorder, rorder, act, has_cmvn, norm_var;
  int np_aff_in, np_lin, np_proj, np_aff_out, np_odim;     // output widths padded to the 128-column GEMM pass
  int sp0, sp1, spm;                                       // row strides (floats) of the shared activation buffers
  int o_mean, o_istd, o_w_in1, o_b_in1, o_w_in2, o_b_in2, o_w_out1, o_b_out1, o_w_out2, o_b_out2;
  int o_layers, layer_stride, lo_wp, lo_taps, lo_wa, lo_ba; // per-layer block: W_p^T, taps [lo+ro][proj], W_a^T, b_a
};

size_t fsmn_smem_bytes(const FsmnArgs& a);
int fsmn_tile_rows();
int fsmn_pass_cols();
int fsmn_launch(FsmnArgs a, cudaStream_t st);

}  // namespace wekws
