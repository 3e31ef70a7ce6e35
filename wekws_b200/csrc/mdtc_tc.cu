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
// One CTA per SM: 16 compute warps + 1 producer warp that owns every asynchronous engine (bulk copies,
// MMA issue).  TWO row tiles (<=128 frames each) are in flight.  There is no CTA-wide barrier in the
// steady state: compute warps hand finished operand tiles to the producer through mbarriers
// (a_rdy[t], 512 arrivals) and keep going; the producer issues the tile's MMAs, which complete on
// mma_bar[t]; compute warps block only when they actually need an accumulator:
//
//   compute : DW(0) DW(1) EPI1(0) EPI1(1) EPI2(0) EPI2(1) |bar| DW(0) ...      (per block)
//   producer:      MMA1(0) MMA1(1)  MMA2(0)  MMA2(1)   + next block's cache slices and weights
//
// Everywhere a lane owns a ROW (frame) of the tile: the depthwise conv reads the time-minor residual
// stream X[c][col] (cache slice and frames of a stream contiguous) conflict-free and writes whole
// 16-byte operand chunks (8 channels of its row); epilogues own the TMEM lane of their row.
#include "common.cuh"
#include "mdtc_tc.h"
#include "tc_common.cuh"

namespace wekws {

namespace {

using namespace tc;

constexpr int NCW = 16;                    // compute warps
constexpr int NCT = NCW * 32;              // compute threads
constexpr int NT_TC = NCT + 32;            // + producer / MMA-issue warp
constexpr int C = 64;
// X[t] holds, per channel c, the streams of the tile back to back in "cat" form
//     X[c][s * Lw + (PADR - pad) .. s * Lw + PADR)   cache slice of the current block (bulk-copied in)
//     X[c][s * Lw + PADR .. + T)                     the residual stream x of stream s
// (Lw = PADR + roundup4(T), PADR = roundup4(max pad)) so cat(cache, x) is simply contiguous columns.
constexpr int RPX = 232;                   // row stride (floats): S * Lw <= XCOLS, + 4 spare (dummy) columns
constexpr int XCOLS = 228;                 // usable columns; column XCOLS is the write target of padding rows
constexpr int A_BYTES = 128 * 128;         // one [128][64] bf16 operand image
constexpr int X_BYTES = 64 * RPX * 4;      // 59392 (multiple of 1024; also hosts the 2 atom-1 images)
constexpr int W_SLOT = 16384;              // hi + lo image of one 64x64 matrix
constexpr int OFF_A = 0;                                   // 2 tiles x (hi, lo)
constexpr int OFF_X = OFF_A + 2 * 2 * A_BYTES;             // 65536
constexpr int OFF_W = OFF_X + 2 * X_BYTES;                 // 184320
constexpr int SMEM_TOTAL = OFF_W + 2 * W_SLOT + 1024;      // + alignment slack = 218112
static_assert(X_BYTES % 1024 == 0 && X_BYTES >= 2 * A_BYTES, "X region must host two operand images");

__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

__global__ void __launch_bounds__(NT_TC, 1) mdtc_tc_kernel(const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t mma_bar[2], halo_bar[2], w_bar[2], a_rdy[2], w_free[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_prod = warp == NCW;
  const int q = warp & 3, g = (warp >> 2) & 3;    // TMEM lane quarter / 16-column group of this warp
  const int row = 32 * q + lane;                  // epilogue row of this thread
  const int T = a.T, K = a.ktaps;
  const float* vec = a.vec;

  uint8_t* Ahi[2] = {base + OFF_A, base + OFF_A + 2 * A_BYTES};
  uint8_t* Alo[2] = {Ahi[0] + A_BYTES, Ahi[1] + A_BYTES};
  float* X[2] = {reinterpret_cast<float*>(base + OFF_X), reinterpret_cast<float*>(base + OFF_X + X_BYTES)};
  uint8_t* Wslot[2] = {base + OFF_W, base + OFF_W + W_SLOT};

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mma_bar[i], 1); mbar_init(&halo_bar[i], 1); mbar_init(&w_bar[i], 1);
      mbar_init(&a_rdy[i], NCT); mbar_init(&w_free[i], 1);
    }
    mbar_fence_init();
  }
  if (is_prod) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // phase parities (every waiter keeps its own copy; all copies advance in lock step)
  uint32_t mma_par[2] = {0, 0}, halo_par[2] = {0, 0}, w_par[2] = {0, 0}, ar_par[2] = {0, 0}, wf_par[2] = {0, 0};
  const uint32_t idesc = make_idesc_bf16(128, 64);
  const int natoms = (a.idim + 63) / 64;
  const int PADR = a.padr, Lw = a.padr + ((T + 3) & ~3);

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
      // ================================================================== PRODUCER WARP
      auto issue_halo = [&](int t, int blk) {       // cache slice of block blk -> pad columns in front of each stream
        if (S[t] == 0) return;
        const int pad = a.dil[blk] * (K - 1), off = a.coff[blk];
        const int nrow = S[t] * C;
        if (a.in_cache != nullptr) {
          if (lane == 0) mbar_arrive_expect_tx(&halo_bar[t], (uint32_t)(nrow * pad * 4));
          __syncwarp();
          for (int r = lane; r < nrow; r += 32) {
            const int s = r >> 6, c = r & 63;
            bulk_g2s(X[t] + c * RPX + s * Lw + PADR - pad, a.in_cache + ((size_t)(b0[t] + s) * C + c) * a.P + off,
                     (uint32_t)(pad * 4), &halo_bar[t]);
          }
        } else {
          for (int r = lane; r < nrow; r += 32) {
            float* dst = X[t] + (r & 63) * RPX + (r >> 6) * Lw + PADR - pad;
            for (int p = 0; p < pad; ++p) dst[p] = 0.f;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&halo_bar[t]);
        }
      };
      auto load_w = [&](int slot, const uint8_t* src) {       // lane 0 only
        mbar_arrive_expect_tx(&w_bar[slot], W_SLOT);
        bulk_g2s(Wslot[slot], src, W_SLOT, &w_bar[slot]);
      };
      // 3-pass bf16x3 GEMM of tile t: D (+)= A * W^T over `ksteps` K-steps of one operand atom
      auto issue_gemm = [&](int t, const uint8_t* ahi, const uint8_t* alo, const uint8_t* wimg, int ksteps,
                            uint32_t& acc) {
        const uint64_t da_hi = make_sdesc_sw128(smem_u32(ahi)), da_lo = make_sdesc_sw128(smem_u32(alo));
        const uint64_t dw_hi = make_sdesc_sw128(smem_u32(wimg)), dw_lo = make_sdesc_sw128(smem_u32(wimg + 8192));
        const uint32_t d = tmem + 64 * t;
        for (int k = 0; k < ksteps; ++k) { umma_bf16(d, sdesc_advance_k(da_hi, k), sdesc_advance_k(dw_hi, k), idesc, acc); acc = 1; }
        for (int k = 0; k < ksteps; ++k) umma_bf16(d, sdesc_advance_k(da_lo, k), sdesc_advance_k(dw_hi, k), idesc, 1);
        for (int k = 0; k < ksteps; ++k) umma_bf16(d, sdesc_advance_k(da_hi, k), sdesc_advance_k(dw_lo, k), idesc, 1);
      };
      auto wait_a = [&](int t) {                     // lane 0: operand images of tile t complete
        mbar_wait(&a_rdy[t], ar_par[t]);
        tc_fence_after();
      };
      const int ks0 = (min(a.idim, 64) + 15) >> 4, ks1 = natoms > 1 ? (a.idim - 64 + 15) >> 4 : 0;

      // ---- first Linear
      if (lane == 0) {
        load_w(0, a.wimg);
        if (natoms > 1) load_w(1, a.wimg + W_SLOT);
        mbar_wait(&w_bar[0], w_par[0]);
        if (natoms > 1) mbar_wait(&w_bar[1], w_par[1]);
        for (int t = 0; t < 2; ++t) {
          if (S[t] == 0) continue;
          wait_a(t);
          uint32_t acc = 0;
          issue_gemm(t, Ahi[t], Alo[t], Wslot[0], ks0, acc);
          if (natoms > 1)
            issue_gemm(t, reinterpret_cast<uint8_t*>(X[t]), reinterpret_cast<uint8_t*>(X[t]) + A_BYTES, Wslot[1], ks1, acc);
          umma_commit(&mma_bar[t]);
        }
        umma_commit(&w_free[0]);
        // the X regions double as atom-1 operand images and both weight slots are busy until these GEMMs finish
        mbar_wait(&w_free[0], wf_par[0]);
        load_w(0, a.wimg + 2 * W_SLOT);
        load_w(1, a.wimg + 3 * W_SLOT);
      }
      w_par[0] ^= 1;
      if (natoms > 1) w_par[1] ^= 1;
      wf_par[0] ^= 1;
      if (S[0]) ar_par[0] ^= 1;
      if (S[1]) ar_par[1] ^= 1;
      __syncwarp();
      issue_halo(0, 0);
      issue_halo(1, 0);

      // ---- blocks
      for (int blk = 0; blk < a.nblocks; ++blk) {
        const bool more = blk + 1 < a.nblocks;
        const uint8_t* wnext = a.wimg + (size_t)(2 + 2 * (blk + 1)) * W_SLOT;
        for (int phase = 0; phase < 2; ++phase) {          // phase 0: pointwise-1 GEMMs, phase 1: conv2 GEMMs
          if (lane == 0) mbar_wait(&w_bar[phase], w_par[phase]);
          w_par[phase] ^= 1;
          for (int t = 0; t < 2; ++t) {
            if (S[t] == 0) continue;
            if (lane == 0) {
              wait_a(t);
              uint32_t acc = 0;
              issue_gemm(t, Ahi[t], Alo[t], Wslot[phase], 4, acc);
              umma_commit(&mma_bar[t]);
            }
            ar_par[t] ^= 1;
            __syncwarp();
            // DW(t) of this block is complete once its operand tile was handed over: its cache columns are free
            if (phase == 0 && more) issue_halo(t, blk + 1);
          }
          if (lane == 0) umma_commit(&w_free[phase]);
        }
        // refill the weight ring (these waits return as soon as the GEMMs above have drained)
        for (int phase = 0; phase < 2; ++phase) {
          if (lane == 0) {
            mbar_wait(&w_free[phase], wf_par[phase]);
            if (more) load_w(phase, wnext + phase * W_SLOT);
          }
          wf_par[phase] ^= 1;
        }
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
      auto hand_over = [&](int t) {                  // operand images of tile t written -> producer may issue
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&a_rdy[t]);
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
        const float* vb = vec + a.v_blocks + blk * a.v_blk_stride;
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
        for (int task = w; task < nrb * 8; task += NCW) {
          const int rb = task >> 3, cg = task & 7;
          const int c0 = col_of(rb, t) - pad;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = cg * 8 + i;
            const float* x0 = X[t] + c * RPX + c0;
            float acc = __ldg(vb + K * C + c);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < K) acc = fmaf(__ldg(vb + j * C + c), x0[j * d], acc);
            v[i] = acc;
          }
          split_store8(v, Ahi[t], Alo[t], sw128_offset(rb * 32 + lane, cg));
        }
        hand_over(t);
      };
      // h = relu(D + b1) -> operand images A[t]                             (mdtc.py:115)
      auto epi1 = [&](int t, int blk) {
        const float* b1 = vec + a.v_blocks + blk * a.v_blk_stride + (K + 1) * C + 16 * g;
        wait_mma(t);
        if (32 * q < rows[t]) {
          float d[16];
          tmem_ld16(tm_lane + 64 * t, d);
#pragma unroll
          for (int hch = 0; hch < 2; ++hch) {
            const float4 ba = __ldg(reinterpret_cast<const float4*>(b1) + 2 * hch);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b1) + 2 * hch + 1);
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
        const float* b2 = vec + a.v_blocks + blk * a.v_blk_stride + (K + 2) * C + 16 * g;
        const bool stack_end = (blk > 0) && (blk % a.stack_size == 0);
        wait_mma(t);
        if (32 * q >= rows[t]) return;
        float d[16];
        tmem_ld16(tm_lane + 64 * t, d);
        float* xp = X[t] + (16 * g) * RPX + col_of(q, t);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(b2) + i4);
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
  if (a.idim % 8 != 0 || a.idim > 128 || a.odim > 8 || a.ktaps > 8) return false;
  if (padmax > 32 || a.P % 4 != 0) return false;
  for (int b = 0; b < a.nblocks; ++b)
    if ((a.dil[b] * (a.ktaps - 1)) % 4 != 0 || a.coff[b] % 4 != 0) return false;
  return true;
}

int tc_max_T() { return 128; }

int mdtc_tc_launch(TcArgs a, int padmax, cudaStream_t st) {
  WEKWS_REQUIRE(a.T >= 1 && a.T <= 128 && a.B >= 1, "mdtc_tc_launch: bad shape");
  a.padr = (padmax + 3) & ~3;
  a.pad_pow2 = 1;
  while (a.pad_pow2 < padmax) a.pad_pow2 <<= 1;
  const int Lw = a.padr + ((a.T + 3) & ~3);
  int smax = 128 / a.T;
  if (smax > XCOLS / Lw) smax = XCOLS / Lw;
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
