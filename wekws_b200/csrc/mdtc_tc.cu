// Tensor-core (tcgen05 / TMEM) fused MDTC forward for hidden_dim 64 -- the throughput path.
//
// Same math as conv_backbone.cu (KWSModel.forward, reference wekws/model/kws_model.py:65-76 with
// mdtc.py:95-121 blocks, BatchNorm folded), but every dense GEMM (first Linear 80->64 and the 34
// pointwise 64x64 convolutions) runs on the 5th-gen tensor cores:
//   * operands in shared memory, K-major SWIZZLE_128B bf16, "x3" split (tc_common.cuh) so the
//     result stays within ~2^-17 of fp32 (posterior error ~5e-6, test bar 1e-4);
//   * accumulators in TMEM (128 lanes x 64 fp32 columns per tile), read back with tcgen05.ld;
//   * weights arrive as pre-swizzled 16 KB images by cp.async.bulk (TMA engine) into a 2-slot ring,
//     cache slices arrive as per-(stream,channel) bulk copies -- all signalled on mbarriers.
//
// One CTA per SM, 16 compute warps + 1 producer/MMA-issue warp, TWO row tiles (<=128 frames each)
// in flight and software-pipelined so that the single-thread-issued MMAs of one tile run under the
// CUDA-core phase of the other:
//     DW(0) | DW(1) | EPI1(0) | EPI1(1) | EPI2(0) | EPI2(1)        (one block; '|' = CTA barrier)
//       MMA1(0) runs under DW(1), MMA1(1) under EPI1(0), MMA2(0) under EPI1(1), MMA2(1) under EPI2(0)
// Everywhere a lane owns a ROW (frame) of the tile: depthwise conv reads the time-minor residual
// stream X[c][row] and cache staging HALO[s][c][p] conflict-free, and writes whole 16-byte operand
// chunks (8 channels of its row); epilogues own the TMEM lane of their row.
#include "common.cuh"
#include "mdtc_tc.h"
#include "tc_common.cuh"

namespace wekws {

namespace {

using namespace tc;

constexpr int NCW = 16;                    // compute warps
constexpr int NT_TC = (NCW + 1) * 32;      // + producer / MMA-issue warp
constexpr int C = 64;
constexpr int RPX = 132;                   // row stride (floats) of X[c][.]
constexpr int A_BYTES = 128 * 128;         // one [128][64] bf16 operand image
constexpr int X_BYTES = 34816;             // 64 * 132 * 4 = 33792, rounded up to 1024 (holds the 2 atom-1 images)
constexpr int HALO_FLOATS = 6144;          // per tile: S * 64 * pad_l <= 6144
constexpr int W_SLOT = 16384;              // hi + lo image of one 64x64 matrix
constexpr int OFF_A = 0;                                   // 2 tiles x (hi, lo)
constexpr int OFF_X = OFF_A + 2 * 2 * A_BYTES;             // 65536
constexpr int OFF_HALO = OFF_X + 2 * X_BYTES;              // 135168
constexpr int OFF_W = OFF_HALO + 2 * HALO_FLOATS * 4;      // 184320
constexpr int SMEM_TOTAL = OFF_W + 2 * W_SLOT + 1024;      // + alignment slack = 218112

__device__ __forceinline__ float cat_tm(const float* __restrict__ xrow, const float* __restrict__ hrow,
                                        int xcol0, int pad, int p) {
  // cat(cache_slice, x) of one (stream, channel) at cat position p; xcol0 = first column of the stream in X
  return p < pad ? hrow[p] : xrow[xcol0 + p - pad];
}

__global__ void __launch_bounds__(NT_TC, 1) mdtc_tc_kernel(const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t mma_bar[2], halo_bar[2], w_bar[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_prod = warp == NCW;
  const int q = warp & 3, g = warp >> 2;          // TMEM lane quarter / 16-column group of this warp
  const int row = 32 * q + lane;                  // epilogue row of this thread
  const int T = a.T, K = a.ktaps;
  const float* vec = a.vec;

  uint8_t* Ahi[2] = {base + OFF_A, base + OFF_A + 2 * A_BYTES};
  uint8_t* Alo[2] = {Ahi[0] + A_BYTES, Ahi[1] + A_BYTES};
  float* X[2] = {reinterpret_cast<float*>(base + OFF_X), reinterpret_cast<float*>(base + OFF_X + X_BYTES)};
  float* HALO[2] = {reinterpret_cast<float*>(base + OFF_HALO), reinterpret_cast<float*>(base + OFF_HALO) + HALO_FLOATS};
  uint8_t* Wslot[2] = {base + OFF_W, base + OFF_W + W_SLOT};

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&mma_bar[i], 1); mbar_init(&halo_bar[i], 1); mbar_init(&w_bar[i], 1); }
    mbar_fence_init();
  }
  if (is_prod) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t mma_par[2] = {0, 0}, halo_par[2] = {0, 0}, w_par[2] = {0, 0};
  const uint32_t idesc = make_idesc_bf16(128, 64);
  const int natoms = (a.idim + 63) / 64;

  // balanced contiguous partition of the streams over the grid
  const int sb = (int)(((long long)a.B * blockIdx.x) / gridDim.x);
  const int se = (int)(((long long)a.B * (blockIdx.x + 1)) / gridDim.x);
  int done = sb;
  const int per_iter_max = 2 * a.smax;

  while (done < se) {
    const int remaining = se - done;
    const int iters_left = (remaining + per_iter_max - 1) / per_iter_max;
    const int take = (remaining + iters_left - 1) / iters_left;
    int S[2], b0[2], rows[2];
    S[0] = (take + 1) / 2; S[1] = take - S[0];
    b0[0] = done; b0[1] = done + S[0];
    rows[0] = S[0] * T; rows[1] = S[1] * T;
    done += take;

    // ------------------------------------------------------------------ producer helpers
    // cache slice of block `blk` for tile t -> HALO[t][(s*64 + c) * pad + p]   (bulk copies, 16..128 B each)
    auto issue_halo = [&](int t, int blk) {
      if (S[t] == 0) return;
      const int pad = a.dil[blk] * (K - 1), off = a.coff[blk];
      const int nrow = S[t] * C;
      if (a.in_cache != nullptr) {
        if (lane == 0) mbar_arrive_expect_tx(&halo_bar[t], (uint32_t)(nrow * pad * 4));
        __syncwarp();
        for (int r = lane; r < nrow; r += 32) {
          const int s = r >> 6, c = r & 63;
          bulk_g2s(HALO[t] + r * pad, a.in_cache + ((size_t)(b0[t] + s) * C + c) * a.P + off, (uint32_t)(pad * 4),
                   &halo_bar[t]);
        }
      } else {
        for (int i = lane; i < nrow * pad; i += 32) HALO[t][i] = 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&halo_bar[t]);
      }
    };
    auto load_w = [&](int slot, const uint8_t* src) {       // lane 0 only
      mbar_arrive_expect_tx(&w_bar[slot], W_SLOT);
      bulk_g2s(Wslot[slot], src, W_SLOT, &w_bar[slot]);
    };
    // 3-pass bf16x3 GEMM of tile t: D = A * W^T over `ksteps` K-steps of one operand atom
    auto issue_gemm = [&](int t, const uint8_t* ahi, const uint8_t* alo, const uint8_t* wimg, int ksteps,
                          uint32_t& acc) {
      const uint64_t da_hi = make_sdesc_sw128(smem_u32(ahi)), da_lo = make_sdesc_sw128(smem_u32(alo));
      const uint64_t dw_hi = make_sdesc_sw128(smem_u32(wimg)), dw_lo = make_sdesc_sw128(smem_u32(wimg + 8192));
      const uint32_t d = tmem + 64 * t;
      for (int k = 0; k < ksteps; ++k) { umma_bf16(d, sdesc_advance_k(da_hi, k), sdesc_advance_k(dw_hi, k), idesc, acc); acc = 1; }
      for (int k = 0; k < ksteps; ++k) umma_bf16(d, sdesc_advance_k(da_lo, k), sdesc_advance_k(dw_hi, k), idesc, 1);
      for (int k = 0; k < ksteps; ++k) umma_bf16(d, sdesc_advance_k(da_hi, k), sdesc_advance_k(dw_lo, k), idesc, 1);
    };

    // ------------------------------------------------------------------ compute helpers
    // features of tile t (+CMVN) -> bf16x3 operand images: atom 0 in A[t], atom 1 (cols 64..) in the X[t] region
    auto feat = [&](int t) {
      const int nch = ((a.idim + 15) >> 4) * 2;             // 16-byte chunks incl. zero padding to a K step
      for (int task = warp; task < 4 * nch; task += NCW) {
        const int rb = task / nch, ch = task - rb * nch;
        const int r = rb * 32 + lane;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        const int k0 = ch * 8;
        if (r < rows[t] && k0 < a.idim) {
          const int s = r / T, tt = r - s * T;
          const float* src = a.feats + (size_t)(b0[t] + s) * a.feat_bstride + (size_t)tt * a.idim + k0;
          const float4 f0 = __ldg(reinterpret_cast<const float4*>(src));
          const float4 f1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
          v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w; v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
          if (a.has_cmvn) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (v[i] - __ldg(vec + a.v_mean + k0 + i)) * __ldg(vec + a.v_istd + k0 + i);
          }
        }
        uint8_t* hi = (ch < 8) ? Ahi[t] : reinterpret_cast<uint8_t*>(X[t]);
        uint8_t* lo = (ch < 8) ? Alo[t] : reinterpret_cast<uint8_t*>(X[t]) + A_BYTES;
        split_store8(v, hi, lo, sw128_offset(r, ch & 7));
      }
    };
    auto wait_mma = [&](int t) {
      mbar_wait(&mma_bar[t], mma_par[t]);
      mma_par[t] ^= 1;
      tc_fence_after();
    };
    // x = relu(D + bp) -> X[t]                                         (subsampling.py:53-57)
    auto epi0 = [&](int t) {
      wait_mma(t);
      float d[16];
      tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + 64 * t + 16 * g, d);
      float* xp = X[t] + (16 * g) * RPX + row;
#pragma unroll
      for (int i = 0; i < 16; ++i) xp[i * RPX] = fmaxf(d[i] + __ldg(vec + a.v_bp + 16 * g + i), 0.f);
      tc_fence_before();
    };
    // new cache slice + depthwise dilated conv (+folded BN) of block blk -> operand images A[t]
    auto dw = [&](int t, int blk) {
      const int d = a.dil[blk], pad = d * (K - 1), off = a.coff[blk];
      const float* vb = vec + a.v_blocks + blk * a.v_blk_stride;
      mbar_wait(&halo_bar[t], halo_par[t]);
      halo_par[t] ^= 1;
      {   // out_cache[b][c][off + j] = cat[c][T + j]                    (mdtc.py:113)
        const int n = S[t] * C * pad;
        for (int idx = tid; idx < n; idx += NCW * 32) {
          const int j = idx % pad, cs = idx / pad, c = cs & 63, s = cs >> 6;
          const float v = cat_tm(X[t] + c * RPX, HALO[t] + cs * pad, s * T, pad, T + j);
          a.out_cache[((size_t)(b0[t] + s) * C + c) * a.P + off + j] = v;
        }
      }
      for (int task = warp; task < 32; task += NCW) {         // 4 row blocks x 8 channel groups
        const int rb = task >> 3, cg = task & 7;
        const int r = rb * 32 + lane;
        const bool valid = r < rows[t];
        const int s = valid ? r / T : 0, tt = valid ? r - s * T : 0;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = cg * 8 + i;
          float acc = __ldg(vb + K * C + c);
          if (valid) {
            const float* xrow = X[t] + c * RPX;
            const float* hrow = HALO[t] + (s * C + c) * pad;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < K) acc = fmaf(__ldg(vb + j * C + c), cat_tm(xrow, hrow, s * T, pad, tt + j * d), acc);
          }
          v[i] = acc;
        }
        split_store8(v, Ahi[t], Alo[t], sw128_offset(r, cg));
      }
      fence_proxy_async();
      tc_fence_before();
    };
    // h = relu(D + b1) -> operand images A[t]                             (mdtc.py:115)
    auto epi1 = [&](int t, int blk) {
      const float* b1 = vec + a.v_blocks + blk * a.v_blk_stride + (K + 1) * C + 16 * g;
      wait_mma(t);
      float d[16];
      tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + 64 * t + 16 * g, d);
#pragma unroll
      for (int hch = 0; hch < 2; ++hch) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(d[hch * 8 + i] + __ldg(b1 + hch * 8 + i), 0.f);
        split_store8(v, Ahi[t], Alo[t], sw128_offset(row, 2 * g + hch));
      }
      fence_proxy_async();
      tc_fence_before();
    };
    // x' = relu(D + b2 + x) -> X[t]; multi-scale sum at the end of a stack  (mdtc.py:116-118, 266-273)
    auto epi2 = [&](int t, int blk, float (&ms)[16]) {
      const float* b2 = vec + a.v_blocks + blk * a.v_blk_stride + (K + 2) * C + 16 * g;
      const bool stack_end = (blk > 0) && (blk % a.stack_size == 0);
      wait_mma(t);
      float d[16];
      tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + 64 * t + 16 * g, d);
      float* xp = X[t] + (16 * g) * RPX + row;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = fmaxf(d[i] + __ldg(b2 + i) + xp[i * RPX], 0.f);
        xp[i * RPX] = v;
        if (stack_end) ms[i] += v;
      }
      tc_fence_before();
    };

    // ============================================================== first Linear (+ReLU)
    if (is_prod) {
      if (lane == 0) {
        load_w(0, a.wimg);
        if (natoms > 1) load_w(1, a.wimg + W_SLOT);
      }
      issue_halo(0, 0);
      issue_halo(1, 0);
    } else {
      feat(0);
      fence_proxy_async();
      tc_fence_before();
    }
    __syncthreads();
    const int ks0 = (min(a.idim, 64) + 15) >> 4, ks1 = natoms > 1 ? (a.idim - 64 + 15) >> 4 : 0;
    if (is_prod) {
      if (lane == 0) {
        mbar_wait(&w_bar[0], w_par[0]);
        if (natoms > 1) mbar_wait(&w_bar[1], w_par[1]);
        tc_fence_after();
        uint32_t acc = 0;
        issue_gemm(0, Ahi[0], Alo[0], Wslot[0], ks0, acc);
        if (natoms > 1) issue_gemm(0, reinterpret_cast<uint8_t*>(X[0]), reinterpret_cast<uint8_t*>(X[0]) + A_BYTES, Wslot[1], ks1, acc);
        umma_commit(&mma_bar[0]);
      }
      w_par[0] ^= 1;
      if (natoms > 1) w_par[1] ^= 1;
    } else if (S[1]) {
      feat(1);
      fence_proxy_async();
      tc_fence_before();
    }
    __syncthreads();
    if (is_prod) {
      if (lane == 0 && S[1]) {
        tc_fence_after();
        uint32_t acc = 0;
        issue_gemm(1, Ahi[1], Alo[1], Wslot[0], ks0, acc);
        if (natoms > 1) issue_gemm(1, reinterpret_cast<uint8_t*>(X[1]), reinterpret_cast<uint8_t*>(X[1]) + A_BYTES, Wslot[1], ks1, acc);
        umma_commit(&mma_bar[1]);
      }
    } else {
      epi0(0);
    }
    __syncthreads();
    if (!is_prod && S[1]) epi0(1);
    __syncthreads();
    // both first-Linear GEMMs are complete (their epilogues waited): the weight slots are free
    if (is_prod && lane == 0) {
      load_w(0, a.wimg + 2 * W_SLOT);
      load_w(1, a.wimg + 3 * W_SLOT);
    }
    float msum[2][16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { msum[0][i] = 0.f; msum[1][i] = 0.f; }

    // ============================================================== blocks
    for (int blk = 0; blk < a.nblocks; ++blk) {
      const bool more = blk + 1 < a.nblocks;
      const uint8_t* wnext = a.wimg + (size_t)(2 + 2 * (blk + 1)) * W_SLOT;
      // ---- DW(0)
      if (!is_prod) dw(0, blk);
      __syncthreads();
      if (is_prod) {
        if (lane == 0) {
          mbar_wait(&w_bar[0], w_par[0]);
          tc_fence_after();
          uint32_t acc = 0;
          issue_gemm(0, Ahi[0], Alo[0], Wslot[0], 4, acc);
          umma_commit(&mma_bar[0]);
        }
        w_par[0] ^= 1;
        if (more) issue_halo(0, blk + 1);
      } else if (S[1]) {
        dw(1, blk);
      }
      __syncthreads();
      // ---- MMA1(1) | EPI1(0)
      if (is_prod) {
        if (lane == 0 && S[1]) {
          tc_fence_after();
          uint32_t acc = 0;
          issue_gemm(1, Ahi[1], Alo[1], Wslot[0], 4, acc);
          umma_commit(&mma_bar[1]);
        }
        if (more) issue_halo(1, blk + 1);
      } else {
        epi1(0, blk);
      }
      __syncthreads();
      // ---- MMA2(0) | EPI1(1)
      if (is_prod) {
        if (lane == 0) {
          mbar_wait(&w_bar[1], w_par[1]);
          tc_fence_after();
          uint32_t acc = 0;
          issue_gemm(0, Ahi[0], Alo[0], Wslot[1], 4, acc);
          umma_commit(&mma_bar[0]);
        }
        w_par[1] ^= 1;
      } else if (S[1]) {
        epi1(1, blk);
      }
      __syncthreads();
      // ---- MMA2(1) | EPI2(0); slot 0 is free (both MMA1 complete: EPI1 waited on them)
      if (is_prod) {
        if (lane == 0) {
          if (S[1]) {
            tc_fence_after();
            uint32_t acc = 0;
            issue_gemm(1, Ahi[1], Alo[1], Wslot[1], 4, acc);
            umma_commit(&mma_bar[1]);
          }
          if (more) load_w(0, wnext);
        }
      } else {
        epi2(0, blk, msum[0]);
      }
      __syncthreads();
      // ---- EPI2(1)
      if (!is_prod && S[1]) epi2(1, blk, msum[1]);
      __syncthreads();
      // slot 1 is free (both MMA2 complete)
      if (is_prod && lane == 0 && more) load_w(1, wnext + W_SLOT);
    }

    // ============================================================== classifier + activation
    // partial dot products over this thread's 16 columns -> scratch (A region of the tile), then reduce
    if (!is_prod) {
      const int odim = a.odim;
      for (int t = 0; t < 2; ++t) {
        if (S[t] == 0) continue;
        float* scratch = reinterpret_cast<float*>(Ahi[t]);          // [4][128][odim]
        for (int j = 0; j < odim; ++j) {
          float p = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) p = fmaf(__ldg(vec + a.v_wc + (16 * g + i) * odim + j), msum[t][i], p);
          scratch[(g * 128 + row) * odim + j] = p;
        }
      }
    }
    __syncthreads();
    if (!is_prod) {
      const int odim = a.odim;
      for (int t = 0; t < 2; ++t) {
        const float* scratch = reinterpret_cast<const float*>(Ahi[t]);
        for (int idx = tid; idx < rows[t] * odim; idx += NCW * 32) {
          const int r = idx / odim, j = idx - r * odim;
          const int s = r / T, tt = r - s * T;
          float y = __ldg(vec + a.v_bc + j);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) y += scratch[(gg * 128 + r) * odim + j];
          if (a.act == WEKWS_ACT_SIGMOID) y = sigmoidf_acc(y);
          a.out[(size_t)(b0[t] + s) * a.out_bstride + (size_t)tt * odim + j] = y;
        }
      }
    }
    __syncthreads();       // scratch / X / HALO are reused by the next iteration
  }

  tc_fence_before();
  __syncthreads();
  if (is_prod) tmem_dealloc(tmem, 128);
}

}  // namespace

bool tc_eligible(const TcArgs& a, int padmax) {
  if (a.idim % 8 != 0 || a.idim > 128 || a.odim > 8 || a.ktaps > 8) return false;
  if (padmax > 96 || a.P % 4 != 0) return false;
  for (int b = 0; b < a.nblocks; ++b)
    if ((a.dil[b] * (a.ktaps - 1)) % 4 != 0 || a.coff[b] % 4 != 0) return false;
  return true;
}

int tc_max_T() { return 128; }

int mdtc_tc_launch(TcArgs a, int padmax, cudaStream_t st) {
  WEKWS_REQUIRE(a.T >= 1 && a.T <= 128 && a.B >= 1, "mdtc_tc_launch: bad shape");
  int smax = 128 / a.T;
  const int hcap = HALO_FLOATS / (C * padmax);
  if (smax > hcap) smax = hcap;
  WEKWS_REQUIRE(smax >= 1, "mdtc_tc_launch: tile does not fit");
  a.smax = smax;
  const int sms = device_sm_count();
  const int grid = a.B < sms ? a.B : sms;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    WEKWS_CUDA_OK(cudaFuncSetAttribute(mdtc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set[dev] = true;
  }
  mdtc_tc_kernel<<<grid, NT_TC, SMEM_TOTAL, st>>>(a);
  return check_launch("mdtc_tc_kernel");
}

}  // namespace wekws
