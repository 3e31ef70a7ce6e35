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
// One CTA per SM: 16 compute warps, one MMA-issue warp (also streams the weights) and one loader
// warp (cache slices by 2-D TMA tensor copies: one instruction per stream and block).  TWO row tiles
// (<=128 frames each) are in flight.  There is no CTA-wide barrier in the steady state: compute warps
// hand finished operand tiles over through mbarriers (a_rdy[t], one arrival per warp) and keep going;
// the issue warp launches the tile's MMAs, which complete on mma_bar[t]; compute warps block only when
// they actually need an accumulator:
//
//   compute : DW(0) DW(1) EPI1(0) EPI1(1) EPI2(0) EPI2(1) |bar| DW(0) ...      (per block)
//   issuer  :      MMA1(0) MMA1(1)  MMA2(0)  MMA2(1)   + next block's weights
//   loader  :      slice(0,blk+1) slice(1,blk+1)
//
// Everywhere a lane owns a ROW (frame) of the tile: the depthwise conv reads the time-minor residual
// stream X[c][col] (cache slice and frames of a stream contiguous) conflict-free and writes whole
// 16-byte operand chunks (8 channels of its row); epilogues own the TMEM lane of their row.
#include <stdlib.h>

#include "common.cuh"
#include "mdtc_tc.h"
#include "tc_common.cuh"

namespace wekws {

namespace {

using namespace tc;

constexpr int NCW = 16;                    // compute warps
constexpr int NCT = NCW * 32;              // compute threads
constexpr int NT_TC = NCT + 96;            // + MMA-issue warp (NCW) + one loader warp per tile (NCW+1, NCW+2)
constexpr int C = 64;
// X[t] holds, per channel c, the streams of the tile back to back in "cat" form
//     X[c][s * Lw + (PADR - pad) .. s * Lw + PADR)   cache slice of the current block (copied in by the loader)
//     X[c][s * Lw + PADR .. + T)                     the residual stream x of stream s
// (Lw = PADR + roundup4(T), PADR = roundup4(max pad)) so cat(cache, x) is simply contiguous columns.
constexpr int RPX = 160;                   // row stride (floats)
constexpr int XCOLS = 156;                 // usable columns (S * Lw <= XCOLS); column XCOLS absorbs padding rows
constexpr int A_BYTES = 128 * 128;         // one [128][64] bf16 operand image
constexpr int X_BYTES = 64 * RPX * 4;      // 40960 (multiple of 1024; also hosts the 2 atom-1 images)
constexpr int STG_FLOATS = 64 * 32;        // TMA landing slot: one stream's cache slice [64][pad <= 32]
constexpr int NSLOT = 4;                   // landing slots: 2 per tile (one block of slices in flight per tile)
constexpr int W_SLOT = 16384;              // hi + lo image of one 64x64 matrix
constexpr int OFF_A = 0;                                   // 2 tiles x (hi, lo)
constexpr int OFF_X = OFF_A + 2 * 2 * A_BYTES;             // 65536
constexpr int OFF_STG = OFF_X + 2 * X_BYTES;               // 147456
constexpr int OFF_W = OFF_STG + NSLOT * STG_FLOATS * 4;    // 188416
constexpr int OFF_VEC = OFF_W + 2 * W_SLOT;                // 221184: per-block vectors (taps, biases), 2 x 2 KB
constexpr int VEC_FLOATS = 512;                            // (K + 3) * 64 floats, K <= 5
constexpr int SMEM_TOTAL = OFF_VEC + 2 * VEC_FLOATS * 4 + 1024;   // + alignment slack = 226304
static_assert(X_BYTES % 1024 == 0 && X_BYTES >= 2 * A_BYTES, "X region must host two operand images");
static_assert(SMEM_TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA may use");

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// 2-D TMA tensor copy global -> shared (box given by the tensor map), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

__global__ void __launch_bounds__(NT_TC, 1) mdtc_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t mma_bar[2], halo_bar[2], w_bar[2], a_rdy[2], w_free[2], h_free[2], stg_bar[NSLOT], vec_bar[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_prod = warp == NCW, is_loader = warp > NCW;
  const int q = warp & 3, g = (warp >> 2) & 3;    // TMEM lane quarter / 16-column group of this warp
  const int row = 32 * q + lane;                  // epilogue row of this thread
  const int T = a.T, K = a.ktaps;
  const float* vec = a.vec;

  uint8_t* Ahi[2] = {base + OFF_A, base + OFF_A + 2 * A_BYTES};
  uint8_t* Alo[2] = {Ahi[0] + A_BYTES, Ahi[1] + A_BYTES};
  float* X[2] = {reinterpret_cast<float*>(base + OFF_X), reinterpret_cast<float*>(base + OFF_X + X_BYTES)};
  float* STG = reinterpret_cast<float*>(base + OFF_STG);      // NSLOT landing slots of STG_FLOATS
  float* VEC = reinterpret_cast<float*>(base + OFF_VEC);      // [2][VEC_FLOATS]
  uint32_t sbase;                                             // shared-window address of `base`, pinned in a register
  asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(base)));
  uint8_t* Wslot[2] = {base + OFF_W, base + OFF_W + W_SLOT};

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mma_bar[i], 1); mbar_init(&halo_bar[i], 1); mbar_init(&w_bar[i], 1);
      mbar_init(&a_rdy[i], NCW); mbar_init(&w_free[i], 1); mbar_init(&h_free[i], NCW); mbar_init(&vec_bar[i], 1);
    }
    for (int i = 0; i < NSLOT; ++i) mbar_init(&stg_bar[i], 1);
    mbar_fence_init();
  }
  if (is_prod) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // phase parities (every waiter keeps its own copy; all copies advance in lock step)
  uint32_t mma_par[2] = {0, 0}, halo_par[2] = {0, 0}, w_par[2] = {0, 0}, ar_par[2] = {0, 0}, wf_par[2] = {0, 0};
  uint32_t hf_par[2] = {0, 0}, lm_par[2] = {0, 0};
  uint32_t vec_par[2] = {0, 0};
  uint32_t jobctr = 0;                           // loader: landing-slot use counter (slot = ctr % NSLOT)
  const int PADR = a.padr, Lw = a.padr + ((T + 3) & ~3);
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

    if (is_prod) {
      // ================================================================== MMA-ISSUE WARP (lane 0 works)
      if (lane == 0) {
        auto load_w = [&](int slot, const uint8_t* src) {
          mbar_arrive_expect_tx(&w_bar[slot], W_SLOT);
          bulk_g2s(Wslot[slot], src, W_SLOT, &w_bar[slot]);
        };
        // 3-pass bf16x3 GEMM of tile t: D (+)= A * W^T over `ksteps` K-steps of one operand atom
        auto issue_gemm = [&](int t, uint64_t da_hi, uint64_t da_lo, uint64_t dw_hi, uint64_t dw_lo, int ksteps,
                              uint32_t& acc) {
          const uint32_t d = tmem + 64 * t;
          for (int k = 0; k < ksteps; ++k) { umma_bf16(d, da_hi + 2 * k, dw_hi + 2 * k, idesc, acc); acc = 1; }
          for (int k = 0; k < ksteps; ++k) umma_bf16(d, da_lo + 2 * k, dw_hi + 2 * k, idesc, 1);
          for (int k = 0; k < ksteps; ++k) umma_bf16(d, da_hi + 2 * k, dw_lo + 2 * k, idesc, 1);
        };
        auto load_vec = [&](int blk) {               // taps + biases of block blk -> VEC[blk & 1]
          mbar_arrive_expect_tx(&vec_bar[blk & 1], (uint32_t)(a.v_blk_stride * 4));
          bulk_g2s(VEC + (blk & 1) * VEC_FLOATS, vec + a.v_blocks + blk * a.v_blk_stride, (uint32_t)(a.v_blk_stride * 4),
                   &vec_bar[blk & 1]);
        };
        auto wait_a = [&](int t) {                   // operand images of tile t complete
          mbar_wait_backoff(&a_rdy[t], ar_par[t]);
          ar_par[t] ^= 1;
          tc_fence_after();
        };
        uint64_t dA_hi[2], dA_lo[2], dX_hi[2], dX_lo[2], dW_hi[2], dW_lo[2];
        for (int i = 0; i < 2; ++i) {
          dA_hi[i] = make_sdesc_sw128(smem_u32(Ahi[i])); dA_lo[i] = make_sdesc_sw128(smem_u32(Alo[i]));
          dX_hi[i] = make_sdesc_sw128(smem_u32(X[i])); dX_lo[i] = make_sdesc_sw128(smem_u32(X[i]) + A_BYTES);
          dW_hi[i] = make_sdesc_sw128(smem_u32(Wslot[i])); dW_lo[i] = make_sdesc_sw128(smem_u32(Wslot[i]) + 8192);
        }
        const int ks0 = (min(a.idim, 64) + 15) >> 4, ks1 = natoms > 1 ? (a.idim - 64 + 15) >> 4 : 0;

        // ---- first Linear
        load_vec(0);                                 // VEC[0] was last read in the previous iteration (odd block count: by block nblocks-1 if even index)
        load_w(0, a.wimg);
        if (natoms > 1) load_w(1, a.wimg + W_SLOT);
        mbar_wait(&w_bar[0], w_par[0]); w_par[0] ^= 1;
        if (natoms > 1) { mbar_wait(&w_bar[1], w_par[1]); w_par[1] ^= 1; }
        for (int t = 0; t < 2; ++t) {
          if (S[t] == 0) continue;
          wait_a(t);
          uint32_t acc = 0;
          issue_gemm(t, dA_hi[t], dA_lo[t], dW_hi[0], dW_lo[0], ks0, acc);
          if (natoms > 1) issue_gemm(t, dX_hi[t], dX_lo[t], dW_hi[1], dW_lo[1], ks1, acc);
          umma_commit(&mma_bar[t]);
        }
        umma_commit(&w_free[0]);
        // both weight slots are busy until these GEMMs finish
        mbar_wait(&w_free[0], wf_par[0]); wf_par[0] ^= 1;
        load_w(0, a.wimg + 2 * W_SLOT);
        load_w(1, a.wimg + 3 * W_SLOT);

        // ---- blocks
        for (int blk = 0; blk < a.nblocks; ++blk) {
          const bool more = blk + 1 < a.nblocks;
          const uint8_t* wnext = a.wimg + (size_t)(2 + 2 * (blk + 1)) * W_SLOT;
          for (int phase = 0; phase < 2; ++phase) {        // phase 0: pointwise-1 GEMMs, phase 1: conv2 GEMMs
            mbar_wait(&w_bar[phase], w_par[phase]); w_par[phase] ^= 1;
            for (int t = 0; t < 2; ++t) {
              if (S[t] == 0) continue;
              wait_a(t);
              uint32_t acc = 0;
              if (a.debug & 1) { mbar_arrive(&mma_bar[t]); }
              else { issue_gemm(t, dA_hi[t], dA_lo[t], dW_hi[phase], dW_lo[phase], 4, acc); umma_commit(&mma_bar[t]); }
            }
            umma_commit(&w_free[phase]);
            if (phase == 1) {      // slot 0 drained long ago: refill it while the conv2 GEMMs run
              mbar_wait(&w_free[0], wf_par[0]); wf_par[0] ^= 1;
              if (more) load_w(0, wnext);
            }
          }
          mbar_wait(&w_free[1], wf_par[1]); wf_par[1] ^= 1;
          if (more) load_w(1, wnext + W_SLOT);
          // VEC[(blk+1)&1] was last read by block blk-1, which every compute warp has left (they handed over EPI1 of blk)
          if (more) load_vec(blk + 1);
        }
      }
    } else if (is_loader) {
      // ================================================================== LOADER WARPS (one per tile)
      // cache slice of (block, stream) of this warp's tile: one 2-D TMA copy [64][pad] into a landing slot, then
      // the warp scatters it into the pad columns in front of the stream's frames in X[t] (float4 copies)
      const int t = warp - NCW - 1;
      if (S[t] > 0) {
        const int St = S[t], njobs = a.nblocks * St;   // jobs ordered by (blk, stream)
        const bool have_cache = a.in_cache != nullptr;
        auto issue_tma = [&](int k) {                  // lane 0
          const int blk = k / St, sidx = k - blk * St;
          const int pad = a.dil[blk] * (K - 1);
          const uint32_t slot = 2 * t + ((jobctr + (uint32_t)k) & 1);
          mbar_arrive_expect_tx(&stg_bar[slot], (uint32_t)(C * pad * 4));
          tma_load_2d(STG + slot * STG_FLOATS, &a.tmap[a.tmap_idx[blk]], a.coff[blk], (b0[t] + sidx) * C, &stg_bar[slot]);
        };
        if (have_cache && lane == 0) issue_tma(0);
        for (int k = 0; k < njobs; ++k) {
          const int blk = k / St, sidx = k - blk * St;
          const int pad = a.dil[blk] * (K - 1);
          // the slot of job k + 1 is the one job k - 1 used: drained (program order + proxy fence below)
          if (have_cache && lane == 0 && k + 1 < njobs) issue_tma(k + 1);
          if (sidx == 0) {                             // first stream of this block: are X[t]'s pad columns free?
            if (lane == 0) {
              if (blk == 0) mbar_wait_backoff(&mma_bar[t], lm_par[t]);    // first-Linear GEMM no longer reads X[t]
              else mbar_wait_backoff(&h_free[t], hf_par[t]);              // DW(t, blk-1) done
            }
            if (blk == 0) lm_par[t] ^= 1; else hf_par[t] ^= 1;            // (mma_bar: odd number of phases per iteration)
          }
          // lane l copies float4 #(l + 32*it): channel c = c0 + it*cstep, float4 v of its pad/4
          const int v4 = pad >> 2, cstep = 32 / v4, c0 = lane / v4, v = lane - c0 * v4, iters = 2 * v4;
          float* dst = X[t] + sidx * Lw + PADR - pad + c0 * RPX + 4 * v;
          const int dstep = cstep * RPX;
          if (have_cache) {
            const uint32_t use = jobctr + (uint32_t)k, slot = 2 * t + (use & 1);
            if (lane == 0) mbar_wait_backoff(&stg_bar[slot], (use >> 1) & 1);
            __syncwarp();
            const float4* src = reinterpret_cast<const float4*>(STG + slot * STG_FLOATS) + lane;
            for (int it = 0; it < iters; it += 2) {      // iters is even: two independent copies in flight
              const float4 x0 = src[32 * it], x1 = src[32 * it + 32];
              *reinterpret_cast<float4*>(dst + it * dstep) = x0;
              *reinterpret_cast<float4*>(dst + (it + 1) * dstep) = x1;
            }
          } else {
            __syncwarp();
            for (int it = 0; it < iters; ++it) *reinterpret_cast<float4*>(dst + it * dstep) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          fence_proxy_async();                         // landing slot was read through the generic proxy; TMA rewrites it
          __syncwarp();
          if (sidx == St - 1 && lane == 0) mbar_arrive(&halo_bar[t]);     // whole tile's slice is in place
        }
        if (have_cache) jobctr += (uint32_t)njobs;
        // DW of the last block still signals h_free: consume it so the parities stay in step
        if (lane == 0) mbar_wait_backoff(&h_free[t], hf_par[t]);
        hf_par[t] ^= 1;
      }
    } else {
      // ================================================================== COMPUTE WARPS
      // X column of frame t of stream s for the rows rb*32+lane this thread may own (XCOLS = dummy for padding rows)
      int colx[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        const int r = rb * 32 + lane, s = r / T;
        colx[rb] = s * Lw + PADR + (r - s * T);
      }
      auto col_of = [&](int rb, int t) {
        const int cx = rb == 0 ? colx[0] : rb == 1 ? colx[1] : rb == 2 ? colx[2] : colx[3];
        return (rb * 32 + lane < rows[t]) ? cx : XCOLS;
      };
      auto hand_over = [&](int t) {                  // this warp's part of tile t's operand images is written
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_rdy[t]);
      };
      auto wait_mma = [&](int t) {
        mbar_wait(&mma_bar[t], mma_par[t]);
        mma_par[t] ^= 1;
        tc_fence_after();
      };
      const uint32_t tm_lane = tmem + ((uint32_t)(32 * q) << 16) + 16 * g;

      // features of tile t (+CMVN) -> bf16x3 operand images: atom 0 in A[t], atom 1 (cols 64..) in the X[t] region
      auto feat = [&](int t) {
        const int nch = ((a.idim + 15) >> 4) * 2;           // 16-byte chunks incl. zero padding to a K step
        const int nrb = (rows[t] + 31) >> 5;
        for (int task = warp; task < nrb * nch; task += NCW) {
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
        hand_over(t);
      };
      // x = relu(D + bp) -> X[t]                                         (subsampling.py:53-57)
      auto epi0 = [&](int t) {
        wait_mma(t);
        if (32 * q >= rows[t]) return;
        float d[16];
        tmem_ld16(tm_lane + 64 * t, d);
        float* xp = X[t] + (16 * g) * RPX + col_of(q, t);
        const float4* bp = reinterpret_cast<const float4*>(vec + a.v_bp + 16 * g);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 b = __ldg(bp + i4);
          xp[(4 * i4 + 0) * RPX] = fmaxf(d[4 * i4 + 0] + b.x, 0.f);
          xp[(4 * i4 + 1) * RPX] = fmaxf(d[4 * i4 + 1] + b.y, 0.f);
          xp[(4 * i4 + 2) * RPX] = fmaxf(d[4 * i4 + 2] + b.z, 0.f);
          xp[(4 * i4 + 3) * RPX] = fmaxf(d[4 * i4 + 3] + b.w, 0.f);
        }
      };
      // new cache slice + depthwise dilated conv (+folded BN) of block blk -> operand images A[t]
      auto dw = [&](int t, int blk) {
        const int d = a.dil[blk], pad = d * (K - 1), off = a.coff[blk];
        const float* vb = VEC + (blk & 1) * VEC_FLOATS;
        if (t == 0) {                                // first use of this block's vectors
          mbar_wait(&vec_bar[blk & 1], vec_par[blk & 1]);
          vec_par[blk & 1] ^= 1;
        }
        mbar_wait(&halo_bar[t], halo_par[t]);
        halo_par[t] ^= 1;
        {   // out_cache[b][c][off + j] = cat[c][T + j]                    (mdtc.py:113)
          int npw = 4;
          while (npw < pad) npw <<= 1;
          const int j = tid & (npw - 1), step = NCT / npw, nrow = S[t] * C;
          if (j < pad) {
            for (int cs = tid / npw; cs < nrow; cs += step) {
              const int s = cs >> 6, c = cs & 63;
              a.out_cache[((size_t)(b0[t] + s) * C + c) * a.P + off + j] = X[t][c * RPX + s * Lw + PADR - pad + T + j];
            }
          }
        }
        // (row block, channel group) tasks; the warp order is mirrored for tile 1 so both tiles together balance
        const int nrb = (rows[t] + 31) >> 5;
        const int w = t ? NCW - 1 - warp : warp;
        const uint32_t xs = sbase + OFF_X + t * X_BYTES;
        for (int task = w; task < ((a.debug & 2) ? 0 : nrb * 8); task += NCW) {
          const int rb = task >> 3, cg = task & 7;
          // tap j of channel c reads X[c][col - pad + j*d]: per tap one base address, channels at immediate offsets
          const uint32_t base = xs + 4u * (uint32_t)(cg * 8 * RPX + col_of(rb, t) - pad);
          const float4* wv = reinterpret_cast<const float4*>(vb + cg * 8);
          float v[8];
          {
            const float4 ba = wv[(K * C) / 4], bb = wv[(K * C) / 4 + 1];
            v[0] = ba.x; v[1] = ba.y; v[2] = ba.z; v[3] = ba.w; v[4] = bb.x; v[5] = bb.y; v[6] = bb.z; v[7] = bb.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < K) {
              const float4 wa = wv[(j * C) / 4], wb = wv[(j * C) / 4 + 1];
              const uint32_t aj = base + 4u * (uint32_t)(j * d);
              v[0] = fmaf(wa.x, lds_f32(aj + 0 * RPX * 4), v[0]);
              v[1] = fmaf(wa.y, lds_f32(aj + 1 * RPX * 4), v[1]);
              v[2] = fmaf(wa.z, lds_f32(aj + 2 * RPX * 4), v[2]);
              v[3] = fmaf(wa.w, lds_f32(aj + 3 * RPX * 4), v[3]);
              v[4] = fmaf(wb.x, lds_f32(aj + 4 * RPX * 4), v[4]);
              v[5] = fmaf(wb.y, lds_f32(aj + 5 * RPX * 4), v[5]);
              v[6] = fmaf(wb.z, lds_f32(aj + 6 * RPX * 4), v[6]);
              v[7] = fmaf(wb.w, lds_f32(aj + 7 * RPX * 4), v[7]);
            }
          }
          split_store8(v, Ahi[t], Alo[t], sw128_offset(rb * 32 + lane, cg));
        }
        hand_over(t);
        if (lane == 0) mbar_arrive(&h_free[t]);      // (after the __syncwarp in hand_over) cache columns consumed
      };
      // h = relu(D + b1) -> operand images A[t]                             (mdtc.py:115)
      auto epi1 = [&](int t, int blk) {
        const float* b1 = VEC + (blk & 1) * VEC_FLOATS + (K + 1) * C + 16 * g;
        wait_mma(t);
        if (32 * q < rows[t] && !(a.debug & 4)) {
          float d[16];
          tmem_ld16(tm_lane + 64 * t, d);
#pragma unroll
          for (int hch = 0; hch < 2; ++hch) {
            const float4 ba = reinterpret_cast<const float4*>(b1)[2 * hch];
            const float4 bb = reinterpret_cast<const float4*>(b1)[2 * hch + 1];
            float v[8];
            v[0] = fmaxf(d[hch * 8 + 0] + ba.x, 0.f); v[1] = fmaxf(d[hch * 8 + 1] + ba.y, 0.f);
            v[2] = fmaxf(d[hch * 8 + 2] + ba.z, 0.f); v[3] = fmaxf(d[hch * 8 + 3] + ba.w, 0.f);
            v[4] = fmaxf(d[hch * 8 + 4] + bb.x, 0.f); v[5] = fmaxf(d[hch * 8 + 5] + bb.y, 0.f);
            v[6] = fmaxf(d[hch * 8 + 6] + bb.z, 0.f); v[7] = fmaxf(d[hch * 8 + 7] + bb.w, 0.f);
            split_store8(v, Ahi[t], Alo[t], sw128_offset(row, 2 * g + hch));
          }
        }
        hand_over(t);
      };
      // x' = relu(D + b2 + x) -> X[t]; multi-scale sum at the end of a stack  (mdtc.py:116-118, 266-273)
      auto epi2 = [&](int t, int blk, float (&ms)[16]) {
        const float* b2 = VEC + (blk & 1) * VEC_FLOATS + (K + 2) * C + 16 * g;
        const bool stack_end = (blk > 0) && (blk % a.stack_size == 0);
        wait_mma(t);
        if (32 * q >= rows[t] || (a.debug & 4)) return;
        float d[16];
        tmem_ld16(tm_lane + 64 * t, d);
        float* xp = X[t] + (16 * g) * RPX + col_of(q, t);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 b = reinterpret_cast<const float4*>(b2)[i4];
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = 4 * i4 + u;
            const float v = fmaxf(d[i] + bb[u] + xp[i * RPX], 0.f);
            xp[i * RPX] = v;
            if (stack_end) ms[i] += v;
          }
        }
      };

      // ---- first Linear (+ReLU)
      feat(0);
      if (S[1]) feat(1);
      epi0(0);
      if (S[1]) epi0(1);
      tc_fence_before();
      compute_barrier();                     // X complete before the first depthwise conv reads across rows
      float msum[2][16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { msum[0][i] = 0.f; msum[1][i] = 0.f; }

      // ---- blocks
      for (int blk = 0; blk < a.nblocks; ++blk) {
        dw(0, blk);
        if (S[1]) dw(1, blk);
        epi1(0, blk);
        if (S[1]) epi1(1, blk);
        epi2(0, blk, msum[0]);
        if (S[1]) epi2(1, blk, msum[1]);
        tc_fence_before();
        compute_barrier();                   // x' of every row complete before the next block's conv
      }

      // ---- classifier + activation: partial dot products over this thread's 16 columns -> scratch -> reduce
      const int odim = a.odim;
      for (int t = 0; t < 2; ++t) {
        if (S[t] == 0) continue;
        float* scratch = reinterpret_cast<float*>(Ahi[t]);            // [4][128][odim] (all GEMMs have drained)
        for (int j = 0; j < odim; ++j) {
          float p = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) p = fmaf(__ldg(vec + a.v_wc + (16 * g + i) * odim + j), msum[t][i], p);
          scratch[(g * 128 + row) * odim + j] = p;
        }
      }
      compute_barrier();
      for (int t = 0; t < 2; ++t) {
        const float* scratch = reinterpret_cast<const float*>(Ahi[t]);
        for (int idx = tid; idx < rows[t] * odim; idx += NCT) {
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
    __syncthreads();       // iteration boundary: scratch / X / operand images are reused
  }

  tc_fence_before();
  __syncthreads();
  if (is_prod) tmem_dealloc(tmem, 128);
}

}  // namespace

bool tc_eligible(const TcArgs& a, int padmax) {
  if (a.idim % 8 != 0 || a.idim > 128 || a.odim > 8 || a.ktaps > 5) return false;
  if (padmax > 32 || a.P % 4 != 0) return false;
  for (int b = 0; b < a.nblocks; ++b)
    if ((a.dil[b] * (a.ktaps - 1)) % 4 != 0 || a.coff[b] % 4 != 0) return false;
  return true;
}

int tc_max_T() { return 120; }

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
}  // namespace

int mdtc_tc_launch(TcArgs a, int padmax, cudaStream_t st) {
  WEKWS_REQUIRE(a.T >= 1 && a.T <= 120 && a.B >= 1, "mdtc_tc_launch: bad shape");
  a.padr = (padmax + 3) & ~3;
  const int Lw = a.padr + ((a.T + 3) & ~3);
  int smax = 128 / a.T;
  if (smax > XCOLS / Lw) smax = XCOLS / Lw;
  WEKWS_REQUIRE(smax >= 1, "mdtc_tc_launch: tile does not fit");
  a.smax = smax;
  {
    const char* dbg = getenv("WEKWS_TC_DEBUG");    // timing experiments only (results are wrong when set)
    a.debug = dbg ? atoi(dbg) : 0;
  }
  // tensor maps over the incoming cache (B*64 rows of P floats): one per distinct slice width
  if (a.in_cache != nullptr) {
    EncodeTiledFn enc = encode_tiled_fn();
    WEKWS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    int pads[4], npads = 0;
    for (int b = 0; b < a.nblocks; ++b) {
      const int pad = a.dil[b] * (a.ktaps - 1);
      int i = 0;
      while (i < npads && pads[i] != pad) ++i;
      if (i == npads) {
        WEKWS_REQUIRE(npads < 4, "more than 4 distinct cache slice widths");
        pads[npads++] = pad;
        const cuuint64_t gdim[2] = {(cuuint64_t)a.P, (cuuint64_t)a.B * C};
        const cuuint64_t gstr[1] = {(cuuint64_t)a.P * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)pad, (cuuint32_t)C};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult rc = enc(&a.tmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.in_cache), gdim,
                                gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        WEKWS_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
      }
      a.tmap_idx[b] = i;
    }
  }
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
