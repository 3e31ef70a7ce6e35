// Kernel argument block of the tensor-core MDTC kernel (mdtc_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_backbone.h"

namespace wekws {

struct TcArgs {
  const float* feats;      // (B, T, idim), stream stride feat_bstride
  const float* in_cache;   // (B, 64, P) or nullptr
  float* out;              // (B, T, odim), stream stride out_bstride
  float* out_cache;        // (B, 64, P)
  const uint8_t* wimg;     // pre-swizzled bf16 hi|lo weight images, 16 KB per slot:
                           //   [Wp atom0][Wp atom1][blk0 W1][blk0 W2][blk1 W1]...
  const float* vec;        // same per-channel vector blob as the FFMA kernel
  int B, T;
  long long feat_bstride, out_bstride;
  int idim, odim, nblocks, ktaps, P, stack_size, act, has_cmvn;
  int v_mean, v_istd, v_bp, v_blocks, v_blk_stride, v_wc, v_bc;
  int dil[kMaxBlocks];
  int coff[kMaxBlocks];
  int debug;               // WEKWS_TC_DEBUG timing-experiment flags (0 in normal use)
  int smax, spt, padr;     // streams per pass / per tile, roundup4(max pad) (set by mdtc_tc_launch)
  int tmap_idx[kMaxBlocks];               // block -> tensor map (one per distinct pad)
  alignas(64) CUtensorMap tmap[4];        // 2-D maps over in_cache viewed as [B*64][P], box [64][pad]
};

bool tc_eligible(const TcArgs& a, int padmax);
int tc_max_T();
int mdtc_tc_launch(TcArgs a, int padmax, cudaStream_t st);

}  // namespace wekws
