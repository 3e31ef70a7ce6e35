// Tensor-core GRU forward on 8-CTA clusters (KWSModel.forward with the GRU backbone, kws_model.py:128-133; PyTorch gate
// order r, z, n -- see gru.cu for the FP32 reference kernel of the same math).
//
// Why clusters: at the streaming shape (B = 512 streams, one frame per call) the FP32 kernel makes every SM stream all
// 786 KB of weights out of L2 for 0.2 GFLOP of math -- 116 MB of L2 egress per 20 us step.  Here a cluster of 8 CTAs
// owns a tile of 64 streams and SPLITS THE HIDDEN UNITS: CTA r keeps only the weights of units [16r, 16r+16) of both
// layers (48 gate rows x 256 inputs per layer, as pre-swizzled bf16 hi|lo images, 104 KB) resident in shared memory for
// the whole launch, computes those units for all 64 streams on tcgen05 (M = 128 rows of which 64 are streams, N = 48,
// bf16 x3 split, fp32 accumulate in TMEM), and hands the new h values to its 7 peers through distributed shared
// memory -- each value is split into bf16 hi/lo once, by its producer, and stored straight into the K-major
// SWIZZLE_128B operand image every CTA's next GEMM reads.  Per step and CTA: 99 MMAs, 96 remote 16-byte stores per
// stream row, three cluster barriers; no weight byte moves after the prologue.
//
// Roles (128 threads): warps 0-1 own the 64 stream rows (TMEM lanes 0..63): feature split, gate math, exchanges,
// classifier; warp 2 lane 0 issues the MMAs; warp 3 loads the weight images (bulk async copies) in the prologue.
#include <string.h>

#include "common.cuh"
#include "gru_tc.h"
#include "tc_common.cuh"

namespace wekws {
namespace {

using namespace tc;

constexpr int CL = 8;                 // CTAs per cluster == hidden-unit slices
constexpr int M = 64;                 // streams per cluster tile
constexpr int H = 128, UPC = H / CL;  // hidden units, units per CTA (16)
constexpr int NG = 3 * UPC;           // gate columns per CTA (48)
constexpr int NT = 128;
constexpr int SLAB = 8192;            // one operand K-slab image: 64 rows x 128 B (rows 64..127 of the M=128 MMA alias the next slab)
// shared memory map (bytes, all 1024-aligned)
constexpr int OFF_AX = 0;                         // X0 image: [slab 0 hi][slab 0 lo][slab 1 hi][slab 1 lo]
constexpr int OFF_AH = OFF_AX + 4 * SLAB;         // H images of layer l at OFF_AH + l * 4 * SLAB
constexpr int OFF_GUARD = OFF_AH + 2 * 4 * SLAB;  // 8 KB the last slab's phantom rows may read
constexpr int OFF_WP = OFF_GUARD + SLAB;          // Wp slice: [slab 0 hi][slab 0 lo][slab 1 hi][slab 1 lo], 16 rows x 128 B each
constexpr int WP_SLAB = 16 * 128;
constexpr int OFF_W = OFF_WP + 4 * WP_SLAB;       // per layer: W_ih [s0 hi][s0 lo][s1 hi][s1 lo], W_hh likewise; 48 rows x 128 B each
constexpr int W_SLABB = NG * 128;                 // 6144
constexpr int W_LAYER = 8 * W_SLABB;              // 49152
constexpr int OFF_HOWN = OFF_W + 2 * W_LAYER;     // fp32 h of this CTA's units: [L][M][UPC]
constexpr int OFF_END = OFF_HOWN + 2 * M * UPC * 4;
constexpr int SMEM_BYTES = OFF_END + 1024;
static_assert(OFF_WP % 1024 == 0 && OFF_W % 1024 == 0 && W_SLABB % 1024 == 0, "operand images must be 1024-byte aligned");
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
// TMEM columns: feature operand hi [0,48) lo [48,96); accumulators D_lin [96,112), D1 = x W_ih^T [112,160),
// D2 = h W_hh^T of layer l at [160 + 48 l, +48)
constexpr int TM_FHI = 0, TM_FLO = 48, TM_DLIN = 96, TM_D1 = 112, TM_D2 = 160, TM_COLS = 256;

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// byte offset of K element k (0..127) of row m inside a two-slab K-major SWIZZLE_128B operand image pair (hi image at
// +0, lo image at +SLAB within each slab pair)
__device__ __forceinline__ uint32_t a_off(int m, int k) {
  return (uint32_t)((k >> 6) * 2 * SLAB + m * 128 + ((((k & 63) >> 3) ^ (m & 7)) << 4) + (k & 7) * 2);
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) gru_tc_kernel(const __grid_constant__ GruTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t w_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t sbase = smem_u32(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CL, nclusters = gridDim.x / CL;
  const float* vec = a.vec;
  const int L = a.L, T = a.T;

  if (tid == 0) {
    mbar_init(&w_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // (no zero fill: rows past the tile's streams and the phantom rows 64..127 of the M = 128 MMAs only ever feed
  // accumulator rows nobody reads -- GEMM rows are independent)
  // ---- prologue: this rank's weight images -> shared memory (stay for the whole launch)
  if (warp == 3 && lane == 0) {
    const uint32_t bytes = 4 * WP_SLAB + (uint32_t)L * W_LAYER;
    const uint8_t* src = a.wimg + (size_t)rank * (4 * WP_SLAB + 2 * W_LAYER);
    mbar_arrive_expect_tx(&w_bar, bytes);
    for (uint32_t o = 0; o < bytes; o += 8192) {
      const uint32_t n = bytes - o < 8192 ? bytes - o : 8192;
      bulk_g2s(base + OFF_WP + o, src + o, n, &w_bar);
    }
  }
  fence_proxy_async();
  __syncthreads();
  cluster_barrier();                       // every CTA's barriers / images exist before anyone stores remotely

  uint32_t mma_par = 0;
  const uint32_t idesc48 = make_idesc_bf16(128, NG), idesc16 = make_idesc_bf16(128, UPC);
  const int ksf = (a.idim + 15) >> 4;      // K steps of the feature GEMM

  // 3-pass bf16x3 GEMM, both operands in shared memory: D (+)= A * W^T over `nslab` K-slabs of 64
  auto gemm_ss = [&](uint32_t d_tmem, uint32_t a_img, uint32_t w_img, uint32_t w_slabb, int nslab, uint32_t idesc) {
    uint32_t acc = 0;
    for (int s = 0; s < nslab; ++s) {
      const uint64_t ahi = make_sdesc_sw128(a_img + s * 2 * SLAB), alo = make_sdesc_sw128(a_img + s * 2 * SLAB + SLAB);
      const uint64_t whi = make_sdesc_sw128(w_img + s * 2 * w_slabb), wlo = make_sdesc_sw128(w_img + s * 2 * w_slabb + w_slabb);
      for (int k = 0; k < 4; ++k) { umma_bf16(d_tmem, ahi + 2 * k, whi + 2 * k, idesc, acc); acc = 1; }
      for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, alo + 2 * k, whi + 2 * k, idesc, 1);
      for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ahi + 2 * k, wlo + 2 * k, idesc, 1);
    }
  };
  auto wait_mma = [&]() {
    mbar_wait(&mma_bar, mma_par);
    mma_par ^= 1;
    tc_fence_after();
  };
  // the MMA issuer (warp 2 lane 0) runs `f` after the row owners' operands are visible; everybody then waits for it
  auto issue = [&](auto f) {
    tc_fence_before();
    fence_proxy_async();                   // generic-proxy writes of the images -> async proxy (tcgen05.mma reads them)
    __syncthreads();
    if (warp == 2 && lane == 0) {
      tc_fence_after();
      f();
      umma_commit(&mma_bar);
    }
    wait_mma();
  };

  for (int tile = cluster_id; tile < a.n_tiles; tile += nclusters) {
    const int b0 = tile * M;
    const int Mv = min(M, a.B - b0);       // valid streams of this tile
    const bool row_owner = tid < M;
    const int m = tid;                     // my stream row (row owners)
    const bool live = row_owner && m < Mv;
    float* hown = reinterpret_cast<float*>(base + OFF_HOWN);
    // ---- initial hidden state: every CTA splits the whole tile's h into its own images; its own units also in fp32.
    // All 128 threads: thread -> (row tid & 63, half tid >> 6 of the 128 units); the 16 float4 loads of a layer are
    // issued together (one memory latency per layer instead of one per 8 values)
    {
      const int hr = tid & (M - 1), half = tid >> 6;
      const bool hlive = hr < Mv;
      for (int l = 0; l < L; ++l) {
        float4 v4[16];
        const float4* src = (a.in_cache != nullptr && hlive)
                                ? reinterpret_cast<const float4*>(a.in_cache + ((size_t)l * a.B + b0 + hr) * H + 64 * half) : nullptr;
#pragma unroll
        for (int i = 0; i < 16; ++i) v4[i] = src ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint8_t* img = base + OFF_AH + l * 4 * SLAB;
#pragma unroll
        for (int c = 0; c < 8; ++c) {                 // 8 chunks of 8 units
          const float4 p0 = v4[2 * c], p1 = v4[2 * c + 1];
          uint4 hi, lo;
          split2(p0.x, p0.y, hi.x, lo.x); split2(p0.z, p0.w, hi.y, lo.y);
          split2(p1.x, p1.y, hi.z, lo.z); split2(p1.z, p1.w, hi.w, lo.w);
          const int k0 = 64 * half + 8 * c;
          const uint32_t off = a_off(hr, k0);
          *reinterpret_cast<uint4*>(img + off) = hi;
          *reinterpret_cast<uint4*>(img + off + SLAB) = lo;
          if ((k0 >> 4) == (int)rank) {
            float* ho = hown + (l * M + hr) * UPC + (k0 & 15);
            *reinterpret_cast<float4*>(ho) = p0;
            *reinterpret_cast<float4*>(ho + 4) = p1;
          }
        }
      }
    }
    if (tile == cluster_id) mbar_wait(&w_bar, 0);      // weights landed (first tile only)

    for (int t = 0; t < T; ++t) {
      // ================= preprocessing Linear + ReLU for my 16 output units            (subsampling.py:53-57)
      if (row_owner) {
        const float* src0 = a.feats + ((size_t)(b0 + m) * T + t) * a.idim;
        // the whole feature row in flight at once (idim <= 96: 24 float4), then CMVN + split chunk by chunk
        float4 f4[24];
        const bool vec4 = (a.idim & 3) == 0 && (reinterpret_cast<uintptr_t>(src0) & 15) == 0;
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          f4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live && 4 * i < a.idim) {
            if (vec4) f4[i] = __ldg(reinterpret_cast<const float4*>(src0) + i);
            else {
              f4[i].x = __ldg(src0 + 4 * i);
              if (4 * i + 1 < a.idim) f4[i].y = __ldg(src0 + 4 * i + 1);
              if (4 * i + 2 < a.idim) f4[i].z = __ldg(src0 + 4 * i + 2);
              if (4 * i + 3 < a.idim) f4[i].w = __ldg(src0 + 4 * i + 3);
            }
          }
        }
#pragma unroll
        for (int ch = 0; ch < 12; ++ch) {
          if (ch < 2 * ksf) {
            float v[8] = {f4[2 * ch].x, f4[2 * ch].y, f4[2 * ch].z, f4[2 * ch].w,
                          f4[2 * ch + 1].x, f4[2 * ch + 1].y, f4[2 * ch + 1].z, f4[2 * ch + 1].w};
            const int k0 = 8 * ch;
            if (a.has_cmvn && live) {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (k0 + u < a.idim) v[u] = (v[u] - __ldg(vec + a.v_mean + k0 + u)) * __ldg(vec + a.v_istd + k0 + u);
            }
            uint32_t h4[4], l4[4];
            split2(v[0], v[1], h4[0], l4[0]); split2(v[2], v[3], h4[1], l4[1]);
            split2(v[4], v[5], h4[2], l4[2]); split2(v[6], v[7], h4[3], l4[3]);
            const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
            tmem_st4(trow + TM_FHI + 4 * ch, h4);
            tmem_st4(trow + TM_FLO + 4 * ch, l4);
          }
        }
        tmem_st_wait();
      }
      // One issue for everything that does not depend on this step's exchanges: the Linear and the h-parts gh = h W_hh^T
      // of BOTH layers (they read the previous step's h images), so only the x-parts sit on the critical path later
      issue([&]() {
        const uint32_t d = tmem + TM_DLIN;
        uint32_t acc = 0;
        for (int s = 0; s * 4 < ksf; ++s) {
          const int ks = min(4, ksf - 4 * s);
          const uint64_t whi = make_sdesc_sw128(sbase + OFF_WP + s * 2 * WP_SLAB), wlo = make_sdesc_sw128(sbase + OFF_WP + s * 2 * WP_SLAB + WP_SLAB);
          for (int k = 0; k < ks; ++k) { umma_bf16_ts(d, tmem + TM_FHI + 32 * s + 8 * k, whi + 2 * k, idesc16, acc); acc = 1; }
          for (int k = 0; k < ks; ++k) umma_bf16_ts(d, tmem + TM_FLO + 32 * s + 8 * k, whi + 2 * k, idesc16, 1);
          for (int k = 0; k < ks; ++k) umma_bf16_ts(d, tmem + TM_FHI + 32 * s + 8 * k, wlo + 2 * k, idesc16, 1);
        }
        for (int l = 0; l < L; ++l)
          gemm_ss(tmem + TM_D2 + NG * l, sbase + OFF_AH + l * 4 * SLAB, sbase + OFF_W + l * W_LAYER + 4 * W_SLABB, W_SLABB, 2, idesc48);
      });
      // x0 of my units -> every CTA's X0 image (the previous step's layer-0 GEMMs are long done: five barriers ago)
      if (row_owner) {
        float d[16];
        tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + TM_DLIN, d);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float x0 = fmaxf(d[2 * u] + __ldg(vec + a.v_bp + UPC * rank + 2 * u), 0.f);
          const float x1 = fmaxf(d[2 * u + 1] + __ldg(vec + a.v_bp + UPC * rank + 2 * u + 1), 0.f);
          split2(x0, x1, hi[u], lo[u]);
        }
        const uint32_t o0 = sbase + OFF_AX + a_off(m, UPC * rank), o1 = sbase + OFF_AX + a_off(m, UPC * rank + 8);
#pragma unroll
        for (uint32_t p = 0; p < CL; ++p) {
          const uint32_t r0 = map_to_cta(o0, p), r1 = map_to_cta(o1, p);
          st_cluster_v4(r0, hi[0], hi[1], hi[2], hi[3]); st_cluster_v4(r1, hi[4], hi[5], hi[6], hi[7]);
          st_cluster_v4(r0 + SLAB, lo[0], lo[1], lo[2], lo[3]); st_cluster_v4(r1 + SLAB, lo[4], lo[5], lo[6], lo[7]);
        }
      }
      tc_fence_before();
      cluster_barrier();                   // X0 complete in every CTA; every CTA's h-part GEMMs have read the old h images

      // ================= GRU layers
      for (int l = 0; l < L; ++l) {
        const uint32_t x_img = l == 0 ? sbase + OFF_AX : sbase + OFF_AH + (l - 1) * 4 * SLAB;
        const uint32_t h_img = sbase + OFF_AH + l * 4 * SLAB;
        const uint32_t w_l = sbase + OFF_W + l * W_LAYER;
        issue([&]() { gemm_ss(tmem + TM_D1, x_img, w_l, W_SLABB, 2, idesc48); });       // gi = x W_ih^T (my 48 gate rows)
        uint32_t hi[8], lo[8];
        if (row_owner) {
          const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
          float gi[48], gh[16];
          tmem_ld16(trow + TM_D1, *reinterpret_cast<float(*)[16]>(&gi[0]));
          tmem_ld16(trow + TM_D1 + 16, *reinterpret_cast<float(*)[16]>(&gi[16]));
          tmem_ld16(trow + TM_D1 + 32, *reinterpret_cast<float(*)[16]>(&gi[32]));
          const float* bih = vec + a.v_layers + (size_t)l * a.v_layer_stride + 2 * H * 3 * H;   // b_ih (384) then b_hh (384)
          const float* bhh = bih + 3 * H;
          float rg[16];
          const uint32_t td2 = trow + TM_D2 + NG * l;
          tmem_ld16(td2, gh);                                                           // r gate
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = UPC * rank + u;
            rg[u] = sigmoidf_acc(gi[u] + __ldg(bih + j) + gh[u] + __ldg(bhh + j));
          }
          float zg[16];
          tmem_ld16(td2 + 16, gh);                                                      // z gate
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = H + UPC * rank + u;
            zg[u] = sigmoidf_acc(gi[16 + u] + __ldg(bih + j) + gh[u] + __ldg(bhh + j));
          }
          tmem_ld16(td2 + 32, gh);                                                      // n gate
          float* ho = hown + (l * M + m) * UPC;
          float hn[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = 2 * H + UPC * rank + u;
            const float n = tanhf(gi[32 + u] + __ldg(bih + j) + rg[u] * (gh[u] + __ldg(bhh + j)));
            hn[u] = (1.f - zg[u]) * n + zg[u] * ho[u];
            ho[u] = hn[u];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) split2(hn[2 * u], hn[2 * u + 1], hi[u], lo[u]);
        }
        // (the old h image was last read by the h-part GEMMs issued at the top of the step, which every CTA finished
        // before the X0 barrier; layer 1's x-part reads the NEW h0 image, after the barrier below)
        if (row_owner) {
          const uint32_t o0 = h_img + a_off(m, UPC * rank), o1 = h_img + a_off(m, UPC * rank + 8);
#pragma unroll
          for (uint32_t p = 0; p < CL; ++p) {
            const uint32_t r0 = map_to_cta(o0, p), r1 = map_to_cta(o1, p);
            st_cluster_v4(r0, hi[0], hi[1], hi[2], hi[3]); st_cluster_v4(r1, hi[4], hi[5], hi[6], hi[7]);
            st_cluster_v4(r0 + SLAB, lo[0], lo[1], lo[2], lo[3]); st_cluster_v4(r1 + SLAB, lo[4], lo[5], lo[6], lo[7]);
          }
        }
        tc_fence_before();
        cluster_barrier();                 // the new h of all 128 units is in every CTA's image
      }

      // ================= classifier on the top layer's h_t: CTA r takes streams [8r, 8r+8) of the tile
      {
        const uint8_t* img = base + OFF_AH + (L - 1) * 4 * SLAB;
        const int per = M / CL;
        for (int o = warp; o < per * a.odim; o += NT / 32) {
          const int sl = o / a.odim, j = o - sl * a.odim, ms = per * (int)rank + sl;
          if (ms >= Mv) continue;
          const float* wc = vec + a.v_wc + j;            // WcT[k][odim]
          float acc = 0.f;
#pragma unroll
          for (int u = 0; u < H / 32; ++u) {
            const int k = lane + 32 * u;
            const uint32_t off = a_off(ms, k);
            const float hv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(img + off)) +
                             __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(img + off + SLAB));
            acc = fmaf(__ldg(wc + k * a.odim), hv, acc);
          }
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
          if (lane == 0) {
            acc += __ldg(vec + a.v_bc + j);
            if (a.act == WEKWS_ACT_SIGMOID) acc = sigmoidf_acc(acc);
            a.out[((size_t)(b0 + ms) * T + t) * a.odim + j] = acc;
          }
        }
      }
      // the classifier's reads of the top image precede this CTA's next cluster barrier, hence every peer's next
      // overwrite of it (which sits behind two more barriers)
    }
    // ---- final hidden state of my units
    if (live) {
      for (int l = 0; l < L; ++l) {
        float* dst = a.out_cache + ((size_t)l * a.B + b0 + m) * H + UPC * rank;
        const float* ho = hown + (l * M + m) * UPC;
#pragma unroll
        for (int u = 0; u < UPC; u += 4) *reinterpret_cast<float4*>(dst + u) = *reinterpret_cast<const float4*>(ho + u);
      }
    }
    cluster_barrier();                     // tile boundary: images are rebuilt for the next tile
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, TM_COLS);
  cluster_barrier();                       // no CTA leaves while peers may still address its shared memory
}

}  // namespace

size_t gru_tc_image_bytes() { return (size_t)CL * (4 * WP_SLAB + 2 * W_LAYER); }

bool gru_tc_eligible(int L, int H_, int idim) { return H_ == H && (L == 1 || L == 2) && idim >= 1 && idim <= 96; }

// host: bf16 hi|lo K-major SWIZZLE_128B images of rank r's slices (tc_common.cuh layout: row n at n*128, 16-byte chunk
// c of a 64-wide K slab at chunk c ^ (n & 7))
void gru_tc_pack(uint8_t* dst, const float* wp /*[H][idim]*/, int idim, const float* const* wih /*[L] of [3H][H]*/,
                 const float* const* whh, int L, uint16_t (*bf16_rn)(float), float (*bf16_to_f)(uint16_t)) {
  const size_t per_rank = 4 * WP_SLAB + 2 * W_LAYER;
  memset(dst, 0, CL * per_rank);
  auto put = [&](uint8_t* img_hi, uint8_t* img_lo, int n, int kk, float w) {
    const uint16_t hi = bf16_rn(w), lo = bf16_rn(w - bf16_to_f(hi));
    const size_t off = (size_t)n * 128 + (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
    memcpy(img_hi + off, &hi, 2);
    memcpy(img_lo + off, &lo, 2);
  };
  for (int r = 0; r < CL; ++r) {
    uint8_t* base = dst + (size_t)r * per_rank;
    for (int n = 0; n < UPC; ++n)                       // preprocessing Linear rows 16r + n
      for (int k = 0; k < idim; ++k)
        put(base + (k >> 6) * 2 * WP_SLAB, base + (k >> 6) * 2 * WP_SLAB + WP_SLAB, n, k & 63, wp[(size_t)(UPC * r + n) * idim + k]);
    for (int l = 0; l < L; ++l)
      for (int which = 0; which < 2; ++which) {
        const float* W = which == 0 ? wih[l] : whh[l];
        uint8_t* wb = base + 4 * WP_SLAB + (size_t)l * W_LAYER + (size_t)which * 4 * W_SLABB;
        for (int n = 0; n < NG; ++n) {
          const int row = (n / UPC) * H + UPC * r + (n % UPC);     // gate (r, z, n) x unit
          for (int k = 0; k < H; ++k)
            put(wb + (k >> 6) * 2 * W_SLABB, wb + (k >> 6) * 2 * W_SLABB + W_SLABB, n, k & 63, W[(size_t)row * H + k]);
        }
      }
  }
}

int gru_tc_launch(GruTcArgs a, cudaStream_t st) {
  WEKWS_REQUIRE(a.B >= 1 && a.T >= 1, "gru_tc_launch: empty call");
  a.n_tiles = (a.B + M - 1) / M;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    WEKWS_CUDA_OK(cudaFuncSetAttribute(gru_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[dev] = true;
  }
  const int sms = device_sm_count();
  int nclusters = a.n_tiles;
  if (nclusters > sms / CL) nclusters = sms / CL;
  gru_tc_kernel<<<nclusters * CL, NT, SMEM_BYTES, st>>>(a);
  return check_launch("gru_tc_kernel");
}

}  // namespace wekws
