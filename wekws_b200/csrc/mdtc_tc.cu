// Tensor-core (tcgen05 / TMEM) fused MDTC forward for hidden_dim 64 -- the throughput path.
//
// Same math as conv_backbone.cu (KWSModel.forward, reference wekws/model/kws_model.py:65-76 with
// mdtc.py:95-121 blocks, BatchNorm folded), but every dense GEMM (first Linear 80->64 and the 34
// pointwise 64x64 convolutions) runs on the 5th-gen tensor cores:
//   * bf16 "x3" operand split (tc_common.cuh): result within ~2^-17 of fp32 (posterior error ~5e-6, bar 1e-4);
//   * the A operand (activations) lives in TENSOR MEMORY: compute threads write their row with tcgen05.st,
//     the MMA reads it from TMEM (tcgen05.mma, A-from-TMEM form); B (weights) is a pre-swizzled K-major
//     SWIZZLE_128B image in shared memory, streamed by cp.async.bulk into a 2-slot ring;
//   * accumulators in TMEM, read back with tcgen05.ld.  Per tile: 64 columns D + 48 hi + 48 lo.
//
// One CTA per SM holds ALL of its streams (up to 7 x 40 frames) resident for the whole network: the
// residual stream X[c][col] (fp32, time-minor, every stream's cache slice directly in front of its
// frames so a dilated tap is a column offset) never leaves shared memory, and up to THREE 128-row
// tiles are in flight.  19 warps: 16 compute, 1 MMA issuer (+ weight / vector ring), 2 loaders (cache
// slices by 2-D TMA tensor copies, scattered into X).  No CTA-wide barrier in the steady state:
//
//   compute : DW(0) DW(1) DW(2) EPI1(0) EPI1(1) EPI1(2) EPI2(0) EPI2(1) EPI2(2) |bar|     (per block)
//   issuer  :      MMA1(0) MMA1(1) MMA1(2)   MMA2(0)  MMA2(1)  MMA2(2)   + next block's weights
//   loaders :      slices(blk+1) of every tile as soon as its DW is done
//
// A lane owns one ROW (frame) of a tile everywhere: row = 32 * (warp % 4) + lane is also its TMEM lane.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "mdtc_tc.h"
#include "tc_common.cuh"

namespace wekws {

namespace {

using namespace tc;

constexpr int NCW = 16;                    // compute warps
constexpr int NCT = NCW * 32;              // compute threads
constexpr int NT_TC = NCT + 96;            // + MMA-issue warp (NCW) + two loader warps (NCW+1, NCW+2)
constexpr int C = 64;
constexpr int NTILE = 3;                   // row tiles in flight
constexpr int RPX = 512;                   // row stride (floats) of X[c][.]
constexpr int XCOLS = 504;                 // usable columns (n_streams * Lw <= XCOLS); column XCOLS absorbs padding rows
constexpr int X_BYTES = 64 * RPX * 4;      // 131072
constexpr int STG_FLOATS = 64 * 32;        // TMA landing slot: one stream's cache slice [64][pad <= 32]
constexpr int NSLOT = 7;                   // loader 0: slots 0..3, loader 1: slots 4..6 (depth 3 / 2 slices in flight)
constexpr int W_SLOT = 16384;              // hi + lo image of one 64x64 matrix
constexpr int VEC_FLOATS = 512;            // per-block vectors: (K + 3) * 64 floats, K <= 5
constexpr int OFF_X = 0;
constexpr int OFF_STG = OFF_X + X_BYTES;                   // 131072
constexpr int OFF_W = OFF_STG + NSLOT * STG_FLOATS * 4;    // 188416
constexpr int OFF_VEC = OFF_W + 2 * W_SLOT;                // 221184
constexpr int SMEM_TOTAL = OFF_VEC + 2 * VEC_FLOATS * 4 + 1024;   // 226304 incl. alignment slack
static_assert(SMEM_TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
// TMEM columns of tile i: [160 i, 160 i + 64) accumulator, + 64.. A hi (<= 48 cols), + 112.. A lo
constexpr int TM_TILE = 160, TM_AHI = 64, TM_ALO = 112, TM_COLS = 512;

__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
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
// splits 8 consecutive K values of this thread's row into 4 hi + 4 lo packed columns and stores them to TMEM
__device__ __forceinline__ void split_to_tmem(const float (&v)[8], uint32_t t_hi, uint32_t t_lo) {
  uint32_t h[4], l[4];
  split2(v[0], v[1], h[0], l[0]); split2(v[2], v[3], h[1], l[1]);
  split2(v[4], v[5], h[2], l[2]); split2(v[6], v[7], h[3], l[3]);
  tmem_st4(t_hi, h);
  tmem_st4(t_lo, l);
}

__global__ void __launch_bounds__(NT_TC, 1) mdtc_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t mma_bar[NTILE], halo_bar[NTILE], a_rdy[NTILE], h_free[NTILE];
  __shared__ uint64_t w_bar[2], w_free[2], vec_bar[2], stg_bar[NSLOT];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == NCW, is_loader = warp > NCW;
  const int q = warp & 3, g = (warp >> 2) & 3;    // TMEM lane quarter / 16-column group of this (compute) warp
  const int row = 32 * q + lane;                  // the row of every tile this thread owns
  const int T = a.T, K = a.ktaps;
  const float* vec = a.vec;

  float* X = reinterpret_cast<float*>(base + OFF_X);
  float* STG = reinterpret_cast<float*>(base + OFF_STG);
  uint8_t* Wslot[2] = {base + OFF_W, base + OFF_W + W_SLOT};
  float* VEC = reinterpret_cast<float*>(base + OFF_VEC);
  uint32_t sbase;                                  // shared-window address of `base`, pinned in a register
  asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(base)));

  if (tid == 0) {
    for (int i = 0; i < NTILE; ++i) {
      mbar_init(&mma_bar[i], 1); mbar_init(&halo_bar[i], 2); mbar_init(&a_rdy[i], NCW); mbar_init(&h_free[i], NCW);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&w_bar[i], 1); mbar_init(&w_free[i], 1); mbar_init(&vec_bar[i], 1); }
    for (int i = 0; i < NSLOT; ++i) mbar_init(&stg_bar[i], 1);
    mbar_fence_init();
  }
  if (is_issuer) tmem_alloc(&tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // phase parities (every waiter keeps its own copy; all copies advance in lock step)
  uint32_t mma_par = 0, halo_par = 0, ar_par = 0, hf_par = 0;       // bit i = parity of tile i's barrier
  uint32_t w_par[2] = {0, 0}, wf_par[2] = {0, 0}, vec_par[2] = {0, 0};
  uint32_t jobctr = 0;                                               // loader: landing-slot use counter
  const uint32_t idesc = make_idesc_bf16(128, 64);
  const int natoms = (a.idim + 63) / 64;
  const int PADR = a.padr, Lw = a.padr + ((T + 3) & ~3);
  const int spt = a.spt;                                             // streams per tile

  // balanced contiguous partition of the streams over the grid
  const int sb = (int)(((long long)a.B * blockIdx.x) / gridDim.x);
  const int se = (int)(((long long)a.B * (blockIdx.x + 1)) / gridDim.x);
  int done = sb;

  while (done < se) {
    const int remaining = se - done;
    const int passes_left = (remaining + a.smax - 1) / a.smax;
    const int ns = (remaining + passes_left - 1) / passes_left;      // streams of this pass (resident in X)
    const int b0 = done;
    done += ns;
    const int ntile = (ns + spt - 1) / spt;
    auto tile_streams = [&](int i) { return min(spt, ns - i * spt); };  // streams of tile i (sequential fill)

    if (is_issuer) {
      // ================================================================== MMA-ISSUE WARP (lane 0 works)
      if (lane == 0) {
        auto load_w = [&](int slot, const uint8_t* src) {
          mbar_arrive_expect_tx(&w_bar[slot], W_SLOT);
          bulk_g2s(Wslot[slot], src, W_SLOT, &w_bar[slot]);
        };
        auto load_vec = [&](int blk) {               // taps + biases of block blk -> VEC[blk & 1]
          mbar_arrive_expect_tx(&vec_bar[blk & 1], (uint32_t)(a.v_blk_stride * 4));
          bulk_g2s(VEC + (blk & 1) * VEC_FLOATS, vec + a.v_blocks + blk * a.v_blk_stride, (uint32_t)(a.v_blk_stride * 4),
                   &vec_bar[blk & 1]);
        };
        // 3-pass bf16x3 GEMM of tile i: D (+)= A * W^T; A from TMEM (8 columns per K step), W image in smem
        auto issue_gemm = [&](int i, int a_col, uint64_t dw_hi, uint64_t dw_lo, int ksteps, uint32_t& acc) {
          const uint32_t d = tmem + TM_TILE * i, ahi = d + TM_AHI + a_col, alo = d + TM_ALO + a_col;
          for (int k = 0; k < ksteps; ++k) { umma_bf16_ts(d, ahi + 8 * k, dw_hi + 2 * k, idesc, acc); acc = 1; }
          for (int k = 0; k < ksteps; ++k) umma_bf16_ts(d, alo + 8 * k, dw_hi + 2 * k, idesc, 1);
          for (int k = 0; k < ksteps; ++k) umma_bf16_ts(d, ahi + 8 * k, dw_lo + 2 * k, idesc, 1);
        };
        auto wait_a = [&](int i) {                   // operand rows of tile i complete
          mbar_wait(&a_rdy[i], (ar_par >> i) & 1);
          ar_par ^= 1u << i;
          tc_fence_after();
        };
        uint64_t dW_hi[2], dW_lo[2];
        for (int i = 0; i < 2; ++i) {
          dW_hi[i] = make_sdesc_sw128(smem_u32(Wslot[i])); dW_lo[i] = make_sdesc_sw128(smem_u32(Wslot[i]) + 8192);
        }
        const int ks0 = (min(a.idim, 64) + 15) >> 4, ks1 = natoms > 1 ? (a.idim - 64 + 15) >> 4 : 0;

        // ---- first Linear
        load_vec(0);
        load_w(0, a.wimg);
        if (natoms > 1) load_w(1, a.wimg + W_SLOT);
        mbar_wait(&w_bar[0], w_par[0]); w_par[0] ^= 1;
        if (natoms > 1) { mbar_wait(&w_bar[1], w_par[1]); w_par[1] ^= 1; }
        for (int i = 0; i < ntile; ++i) {
          wait_a(i);
          uint32_t acc = 0;
          issue_gemm(i, 0, dW_hi[0], dW_lo[0], ks0, acc);
          if (natoms > 1) issue_gemm(i, 32, dW_hi[1], dW_lo[1], ks1, acc);
          umma_commit(&mma_bar[i]);
        }
        umma_commit(&w_free[0]);
        mbar_wait(&w_free[0], wf_par[0]); wf_par[0] ^= 1;   // both weight slots are busy until these GEMMs finish
        load_w(0, a.wimg + 2 * W_SLOT);
        load_w(1, a.wimg + 3 * W_SLOT);

        // ---- blocks
        for (int blk = 0; blk < a.nblocks; ++blk) {
          const bool more = blk + 1 < a.nblocks;
          const uint8_t* wnext = a.wimg + (size_t)(2 + 2 * (blk + 1)) * W_SLOT;
          for (int phase = 0; phase < 2; ++phase) {        // phase 0: pointwise-1 GEMMs, phase 1: conv2 GEMMs
            mbar_wait(&w_bar[phase], w_par[phase]); w_par[phase] ^= 1;
            for (int i = 0; i < ntile; ++i) {
              wait_a(i);
              uint32_t acc = 0;
              if (a.debug & 1) { mbar_arrive(&mma_bar[i]); }
              else { issue_gemm(i, 0, dW_hi[phase], dW_lo[phase], 4, acc); umma_commit(&mma_bar[i]); }
            }
            umma_commit(&w_free[phase]);
            if (phase == 1) {      // slot 0 drained long ago: refill it while the conv2 GEMMs run
              mbar_wait(&w_free[0], wf_par[0]); wf_par[0] ^= 1;
              if (more) load_w(0, wnext);
            }
          }
          mbar_wait(&w_free[1], wf_par[1]); wf_par[1] ^= 1;
          if (more) load_w(1, wnext + W_SLOT);
          // VEC[(blk+1)&1] was last read by block blk-1, which every compute warp has left
          if (more) load_vec(blk + 1);
        }
      }
    } else if (is_loader) {
      // ================================================================== LOADER WARPS
      // loader l owns the streams sg with sg % 2 == l.  Per (block, stream): one 2-D TMA copy [64][pad] into a
      // landing slot, then the warp scatters it into the pad columns in front of the stream's frames in X.
      const int l = warp - NCW - 1;
      const int nmine = (ns - l + 1) / 2;              // my streams: l, l + 2, ...
      const int njobs = a.nblocks * nmine;
      const bool have_cache = a.in_cache != nullptr && !(a.debug & 16);
      const int nsl = l == 0 ? 4 : 3, slot0 = l == 0 ? 0 : 4;   // my landing slots: a ring of nsl
      auto issue_tma = [&](int k) {                    // lane 0; job k = (blk, my m-th stream)
        const int blk = k / nmine, sg = l + 2 * (k - blk * nmine);
        const int pad = a.dil[blk] * (K - 1);
        const uint32_t slot = slot0 + (jobctr + (uint32_t)k) % nsl;
        mbar_arrive_expect_tx(&stg_bar[slot], (uint32_t)(C * pad * 4));
        tma_load_2d(STG + slot * STG_FLOATS, &a.tmap[a.tmap_idx[blk]], a.coff[blk], (b0 + sg) * C, &stg_bar[slot]);
      };
      if (have_cache && lane == 0)
        for (int k0 = 0; k0 < nsl - 1 && k0 < njobs; ++k0) issue_tma(k0);
      int k = 0;
      for (int blk = 0; blk < a.nblocks; ++blk) {
        const int pad = a.dil[blk] * (K - 1);
        for (int i = 0; i < ntile; ++i) {
          // X's pad columns of tile i are free once DW(i, blk-1) is done (at blk 0 they are free from the start)
          if (blk > 0) {
            if (lane == 0) mbar_wait_backoff(&h_free[i], (hf_par >> i) & 1);
            hf_par ^= 1u << i;
            __syncwarp();
          }
          for (int sg = i * spt; sg < i * spt + tile_streams(i); ++sg) {
            if ((sg & 1) != l) continue;
            if (have_cache && lane == 0 && k + nsl - 1 < njobs) issue_tma(k + nsl - 1);   // reuses the slot of job k-1: drained
            const int v4 = pad >> 2, sh = 31 - __clz(v4), cstep = 32 >> sh, c0 = lane >> sh, v = lane & (v4 - 1), iters = 2 * v4;
            float* dst = X + sg * Lw + PADR - pad + c0 * RPX + 4 * v;
            const int dstep = cstep * RPX;
            if (have_cache) {
              const uint32_t use = jobctr + (uint32_t)k, slot = slot0 + use % nsl;
              if (lane == 0) mbar_wait_backoff(&stg_bar[slot], (use / nsl) & 1);
              __syncwarp();
              const float4* src = reinterpret_cast<const float4*>(STG + slot * STG_FLOATS) + lane;
              for (int it = 0; it < iters; it += 2) {
                const float4 x0 = src[32 * it], x1 = src[32 * it + 32];
                *reinterpret_cast<float4*>(dst + it * dstep) = x0;
                *reinterpret_cast<float4*>(dst + (it + 1) * dstep) = x1;
              }
            } else {
              for (int it = 0; it < iters; ++it) *reinterpret_cast<float4*>(dst + it * dstep) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            fence_proxy_async();                     // the slot was read through the generic proxy; TMA rewrites it
            ++k;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&halo_bar[i]);  // my share of tile i's slices is in place (count 2: both loaders)
        }
      }
      if (have_cache) jobctr += (uint32_t)njobs;
      // DW of the last block still signals h_free: consume it so the parities stay in step
      for (int i = 0; i < ntile; ++i) {
        if (lane == 0) mbar_wait_backoff(&h_free[i], (hf_par >> i) & 1);
        hf_par ^= 1u << i;
      }
    } else {
      // ================================================================== COMPUTE WARPS
      // this thread's row of tile i -> (stream, frame) -> X column; padding rows use the dummy column
      int colx[NTILE], rows_i[NTILE];
#pragma unroll
      for (int i = 0; i < NTILE; ++i) {
        rows_i[i] = i < ntile ? tile_streams(i) * T : 0;
        const int s = row / T;
        colx[i] = row < rows_i[i] ? (i * spt + s) * Lw + PADR + (row - s * T) : XCOLS;
      }
      const bool q_live[NTILE] = {32 * q < rows_i[0], 32 * q < rows_i[1], 32 * q < rows_i[2]};
      const uint32_t tm_row = tmem + ((uint32_t)(32 * q) << 16);
      const uint32_t xs = sbase + OFF_X;
      // Cache stores go to the warps with the least depthwise work: lane quarters that are padding in some tile
      // (e.g. tiles of 120/120/40 rows: quarters 2,3 skip tile 2).  If every quarter is equally busy, all store.
      int nlive[4], maxlive = 0, minlive = NTILE;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        nlive[qq] = (32 * qq < rows_i[0]) + (32 * qq < rows_i[1]) + (32 * qq < rows_i[2]);
        maxlive = max(maxlive, nlive[qq]);
        minlive = min(minlive, nlive[qq]);
      }
      int nhq = 0, myrank = -1;                        // helper quarters and this warp's rank among them
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const bool helper = (maxlive == minlive) || (nlive[qq] == minlive);
        if (helper) { if (qq == q) myrank = nhq; ++nhq; }
      }
      const int n_store_thr = nhq * 4 * 32;            // 4 warps (g = 0..3) per quarter
      const int store_idx = myrank < 0 ? -1 : (g * nhq + myrank) * 32 + lane;
      float part[NTILE][8];                          // classifier partial sums over this thread's 16 channels
#pragma unroll
      for (int i = 0; i < NTILE; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) part[i][j] = 0.f;

      auto hand_over = [&](auto tc) {                // this warp's rows of tile i are in TMEM
        constexpr int i = decltype(tc)::value;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_rdy[i]);
      };
      auto wait_mma = [&](auto tc) {
        constexpr int i = decltype(tc)::value;
        mbar_wait(&mma_bar[i], (mma_par >> i) & 1);
        mma_par ^= 1u << i;
        tc_fence_after();
      };
      // features of tile i (+CMVN) -> bf16x3 operand rows in TMEM (8 K values = 4 packed columns per chunk)
      auto feat = [&](auto tc) {
        constexpr int i = decltype(tc)::value;
        if (q_live[i]) {
          const int nch = ((a.idim + 15) >> 4) * 2;           // 16-byte chunks incl. zero padding to a K step
          const bool valid = row < rows_i[i];
          const int s = row / T, tt = row - s * T;
          const float* src0 = a.feats + (size_t)(b0 + i * spt + s) * a.feat_bstride + (size_t)tt * a.idim;
          for (int ch = g; ch < nch; ch += 4) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = 0.f;
            const int k0 = ch * 8;
            if (valid && k0 < a.idim) {
              const float4 f0 = __ldg(reinterpret_cast<const float4*>(src0 + k0));
              const float4 f1 = __ldg(reinterpret_cast<const float4*>(src0 + k0) + 1);
              v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w; v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
              if (a.has_cmvn) {
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (v[u] - __ldg(vec + a.v_mean + k0 + u)) * __ldg(vec + a.v_istd + k0 + u);
              }
            }
            split_to_tmem(v, tm_row + TM_TILE * i + TM_AHI + 4 * ch, tm_row + TM_TILE * i + TM_ALO + 4 * ch);
          }
        }
        hand_over(tc);
      };
      // x = relu(D + bp) -> X                                            (subsampling.py:53-57)
      auto epi0 = [&](auto tc) {
        constexpr int i = decltype(tc)::value;
        wait_mma(tc);
        if (!q_live[i]) return;
        float d[16];
        tmem_ld16(tm_row + TM_TILE * i + 16 * g, d);
        float* xp = X + (16 * g) * RPX + colx[i];
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
      // new cache slice + depthwise dilated conv (+folded BN) of block blk -> operand rows of tile i in TMEM
      auto dw = [&](auto tc, int blk) {
        constexpr int i = decltype(tc)::value;
        const int d = a.dil[blk], pad = d * (K - 1), off = a.coff[blk];
        const float* vb = VEC + (blk & 1) * VEC_FLOATS;
        if (i == 0) {                                // first use of this block's vectors
          mbar_wait(&vec_bar[blk & 1], vec_par[blk & 1]);
          vec_par[blk & 1] ^= 1;
        }
        mbar_wait(&halo_bar[i], (halo_par >> i) & 1);
        halo_par ^= 1u << i;
        if (q_live[i] && !(a.debug & 2)) {
          // this warp: rows 32q.., channel groups g and g + 4; tap j of channel c reads X[c][col - pad + j*d]
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cg = g + 4 * half;
            const uint32_t bse = xs + 4u * (uint32_t)(cg * 8 * RPX + colx[i] - pad);
            const float4* wv = reinterpret_cast<const float4*>(vb + cg * 8);
            float v[8];
            {
              const float4 ba = wv[(K * C) / 4], bb = wv[(K * C) / 4 + 1];
              v[0] = ba.x; v[1] = ba.y; v[2] = ba.z; v[3] = ba.w; v[4] = bb.x; v[5] = bb.y; v[6] = bb.z; v[7] = bb.w;
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              if (j < K) {
                const float4 wa = wv[(j * C) / 4], wb = wv[(j * C) / 4 + 1];
                const uint32_t aj = bse + 4u * (uint32_t)(j * d);
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
            split_to_tmem(v, tm_row + TM_TILE * i + TM_AHI + 4 * cg, tm_row + TM_TILE * i + TM_ALO + 4 * cg);
          }
        }
        hand_over(tc);
      };
      // new cache slices of every tile: out_cache[b][c][off + j] = cat[c][T + j]   (mdtc.py:113), by the helper warps;
      // then every warp releases the tiles' cache columns to the loaders.  All of this precedes the warp's EPI1
      // hand-overs, hence the conv2 GEMMs and EPI2 (which overwrites x) cannot start before the stores are done.
      auto store_cache = [&](int blk) {
        const int d = a.dil[blk], pad = d * (K - 1), off = a.coff[blk];
        if (store_idx >= 0 && !(a.debug & 8)) {
          for (int i = 0; i < ntile; ++i) {
            const int nst = tile_streams(i), sg0 = i * spt;
            if ((T & 3) == 0) {      // rows are 16-byte aligned; pad is a power of two (tc_eligible)
              const int v4 = pad >> 2, n4 = nst * C * v4, sh = 31 - __clz(v4);
              for (int e = store_idx; e < n4; e += n_store_thr) {
                const int cs = e >> sh, v = e & (v4 - 1), s = cs >> 6, c = cs & 63;
                const float4 x4 = *reinterpret_cast<const float4*>(X + c * RPX + (sg0 + s) * Lw + PADR - pad + T + 4 * v);
                *reinterpret_cast<float4*>(a.out_cache + ((size_t)(b0 + sg0 + s) * C + c) * a.P + off + 4 * v) = x4;
              }
            } else {
              const int n = nst * C * pad, sh = 31 - __clz(pad);
              for (int e = store_idx; e < n; e += n_store_thr) {
                const int cs = e >> sh, j = e & (pad - 1), s = cs >> 6, c = cs & 63;
                a.out_cache[((size_t)(b0 + sg0 + s) * C + c) * a.P + off + j] = X[c * RPX + (sg0 + s) * Lw + PADR - pad + T + j];
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0)
          for (int i = 0; i < ntile; ++i) mbar_arrive(&h_free[i]);
      };
      // h = relu(D + b1) -> operand rows of tile i in TMEM                  (mdtc.py:115)
      auto epi1 = [&](auto tc, int blk) {
        constexpr int i = decltype(tc)::value;
        const float* b1 = VEC + (blk & 1) * VEC_FLOATS + (K + 1) * C + 16 * g;
        wait_mma(tc);
        if (q_live[i] && !(a.debug & 4)) {
          float d[16];
          tmem_ld16(tm_row + TM_TILE * i + 16 * g, d);
#pragma unroll
          for (int hch = 0; hch < 2; ++hch) {
            const float4 ba = reinterpret_cast<const float4*>(b1)[2 * hch];
            const float4 bb = reinterpret_cast<const float4*>(b1)[2 * hch + 1];
            float v[8];
            v[0] = fmaxf(d[hch * 8 + 0] + ba.x, 0.f); v[1] = fmaxf(d[hch * 8 + 1] + ba.y, 0.f);
            v[2] = fmaxf(d[hch * 8 + 2] + ba.z, 0.f); v[3] = fmaxf(d[hch * 8 + 3] + ba.w, 0.f);
            v[4] = fmaxf(d[hch * 8 + 4] + bb.x, 0.f); v[5] = fmaxf(d[hch * 8 + 5] + bb.y, 0.f);
            v[6] = fmaxf(d[hch * 8 + 6] + bb.z, 0.f); v[7] = fmaxf(d[hch * 8 + 7] + bb.w, 0.f);
            const int ch = 2 * g + hch;
            split_to_tmem(v, tm_row + TM_TILE * i + TM_AHI + 4 * ch, tm_row + TM_TILE * i + TM_ALO + 4 * ch);
          }
        }
        hand_over(tc);
      };
      // x' = relu(D + b2 + x) -> X; classifier partial sums at the end of a stack (mdtc.py:116-118, 266-273)
      auto epi2 = [&](auto tc, int blk) {
        constexpr int i = decltype(tc)::value;
        const float* b2 = VEC + (blk & 1) * VEC_FLOATS + (K + 2) * C + 16 * g;
        const bool stack_end = (blk > 0) && (blk % a.stack_size == 0);
        wait_mma(tc);
        if (!q_live[i] || (a.debug & 4)) return;
        float d[16];
        tmem_ld16(tm_row + TM_TILE * i + 16 * g, d);
        float* xp = X + (16 * g) * RPX + colx[i];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 b = reinterpret_cast<const float4*>(b2)[i4];
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = 4 * i4 + u;
            d[e] = fmaxf(d[e] + bb[u] + xp[e * RPX], 0.f);
            xp[e * RPX] = d[e];
          }
        }
        if (stack_end) {      // the classifier is linear: W_c (sum of stack outputs) = sum of W_c (stack output)
          const float* wc = vec + a.v_wc + (16 * g) * a.odim;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < a.odim) {
              float p = part[i][j];
#pragma unroll
              for (int e = 0; e < 16; ++e) p = fmaf(__ldg(wc + e * a.odim + j), d[e], p);
              part[i][j] = p;
            }
          }
        }
      };

      constexpr std::integral_constant<int, 0> T0{};
      constexpr std::integral_constant<int, 1> T1{};
      constexpr std::integral_constant<int, 2> T2{};
      // ---- first Linear (+ReLU)
      feat(T0);
      if (ntile > 1) feat(T1);
      if (ntile > 2) feat(T2);
      epi0(T0);
      if (ntile > 1) epi0(T1);
      if (ntile > 2) epi0(T2);
      tc_fence_before();
      compute_barrier();                     // X complete before the first depthwise conv reads across rows

      // ---- blocks
      for (int blk = 0; blk < a.nblocks; ++blk) {
        dw(T0, blk);
        if (ntile > 1) dw(T1, blk);
        if (ntile > 2) dw(T2, blk);
        store_cache(blk);
        epi1(T0, blk);
        if (ntile > 1) epi1(T1, blk);
        if (ntile > 2) epi1(T2, blk);
        epi2(T0, blk);
        if (ntile > 1) epi2(T1, blk);
        if (ntile > 2) epi2(T2, blk);
        tc_fence_before();
        compute_barrier();                   // x' of every row complete before the next block's conv
      }

      // ---- classifier + activation: reduce the 4 column-group partials of each row through X (now dead)
      const int odim = a.odim;
      float* scratch = X;                                             // [NTILE][4][128][odim]
#pragma unroll
      for (int i = 0; i < NTILE; ++i)
        if (i < ntile) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < odim) scratch[((i * 4 + g) * 128 + row) * odim + j] = part[i][j];
        }
      compute_barrier();
      for (int i = 0; i < ntile; ++i) {
        const int nrow = tile_streams(i) * T;
        for (int idx = tid; idx < nrow * odim; idx += NCT) {
          const int r = idx / odim, j = idx - r * odim;
          const int s = r / T, tt = r - s * T;
          float y = __ldg(vec + a.v_bc + j);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) y += scratch[((i * 4 + gg) * 128 + r) * odim + j];
          if (a.act == WEKWS_ACT_SIGMOID) y = sigmoidf_acc(y);
          a.out[(size_t)(b0 + i * spt + s) * a.out_bstride + (size_t)tt * odim + j] = y;
        }
      }
    }
    __syncthreads();       // pass boundary: X (and the classifier scratch inside it) is reused
  }

  tc_fence_before();
  __syncthreads();
  if (is_issuer) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

bool tc_eligible(const TcArgs& a, int padmax) {
  if (a.idim % 8 != 0 || a.idim > 128 || a.odim > 8 || a.ktaps > 5) return false;
  if (padmax > 32 || a.P % 4 != 0) return false;
  for (int b = 0; b < a.nblocks; ++b)
  {
    const int pad = a.dil[b] * (a.ktaps - 1);
    if (pad < 4 || (pad & (pad - 1)) != 0 || a.coff[b] % 4 != 0) return false;   // power-of-two slices, 16-byte aligned
  }
  return true;
}

int tc_max_T() { return 128; }

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
  WEKWS_REQUIRE(a.T >= 1 && a.T <= 128 && a.B >= 1, "mdtc_tc_launch: bad shape");
  a.padr = (padmax + 3) & ~3;
  const int Lw = a.padr + ((a.T + 3) & ~3);
  a.spt = 128 / a.T;                                   // streams per 128-row tile
  WEKWS_REQUIRE(a.spt >= 1 && Lw <= XCOLS, "mdtc_tc_launch: tile does not fit");
  int smax = NTILE * a.spt;                            // streams resident per pass
  if (smax > XCOLS / Lw) smax = XCOLS / Lw;
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
