// sm_100a building blocks for the tensor-core path: mbarrier, bulk async copies (TMA engine),
// tcgen05 MMA / TMEM alloc / TMEM load, UMMA shared-memory + instruction descriptors, and the
// bf16 "x3" operand split (x ~ b0 + b1, products b0*w0 + b1*w0 + b0*w1, ~2^-17 relative error).
//
// Operand layout used everywhere here: K-major, SWIZZLE_128B, bf16.  A row of a 64-wide K slab
// is 128 bytes = 8 chunks of 16 B; chunk j of row r lives at position j ^ (r & 7); rows are
// packed 8 x 128 B = 1024 B per core-matrix group (SBO = 1024), so an [R][64] bf16 operand is a
// dense R*128-byte image whose base must be 1024-byte aligned.  One tcgen05.mma consumes K = 16
// (32 bytes of every row): the K step advances the descriptor start address by 32 bytes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace wekws {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifdef WEKWS_MBAR_WATCHDOG
// debug build: a wait that never completes reports which barrier it was (plus every warp's last progress mark)
// and traps instead of hanging the GPU
static __device__ int wd_marks[256][20];
__device__ __forceinline__ void wd_mark(int v) {
  if ((threadIdx.x & 31) == 0) *(volatile int*)&wd_marks[blockIdx.x & 255][(threadIdx.x >> 5) % 20] = v;
}
static __device__ __noinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (long long spin = 0; spin < (1ll << 22); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
    volatile int* m = wd_marks[0];
    printf("mbar_wait timeout: thread %d barrier@smem 0x%x parity %u marks c0 %d c1 %d c2 %d c3 %d c4 %d c8 %d c12 %d c15 %d iss %d ld0 %d ld1 %d\n",
           (int)threadIdx.x, smem_u32(bar), parity, m[0], m[1], m[2], m[3], m[4], m[8], m[12], m[15], m[16], m[17], m[18]);
  }
  for (int i = 0; i < 2000; ++i) __nanosleep(100000);       // let the other waiters report too
  __trap();
}
#else
__device__ __forceinline__ void wd_mark(int) {}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
#endif

// same, for single-lane service warps: back off between polls so the spin does not eat issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
#ifdef WEKWS_MBAR_WATCHDOG
  mbar_wait(bar, parity);
  return;
#endif
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(40);
  }
}

// -------------------------------------------------------------------- bulk async copies
// global -> shared, completion signalled on an mbarrier as transaction bytes (16 B granularity)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// ---------------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// wait that traps instead of hanging the GPU if the phase never completes (experimental kernels)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (long long spin = 0; spin < (1ll << 26); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// wait with cluster-scope acquire (the arrivals come from other CTAs); bounded like the one above
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long spin = 0; !done; ++spin) {
    if (spin > (1ll << 26)) __trap();
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// global -> the same shared-memory offset of every CTA in cta_mask; each destination's mbarrier (same offset) gets
// the bytes as complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {         // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp.  MMA issue loops should be executed by the WHOLE warp with only the instruction itself
// under `if (elect_one_sync())`: the descriptors are then computed in uniform registers and every tcgen05.mma is a single
// instruction.  With `if (lane == 0)` around a loop that derives descriptors from values the compiler cannot prove
// warp-uniform (function arguments, shared-memory loads), every MMA is wrapped in an ELECT / R2UR waterfall loop and one
// M128 N64 K16 dispatch costs ~105 cycles instead of ~55 (tests/native/mma_rate_probe.cu).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// warp-uniform copy of a value every lane holds (lets the compiler keep what is derived from it in uniform registers)
__device__ __forceinline__ uint32_t uniform32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from TMEM (row r of A = TMEM lane r; K packed two bf16 per 32-bit column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 4 / 8 consecutive 32-bit columns of this thread's TMEM lane <- registers (warp-collective)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all tcgen05 ops issued so far by this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive fp32 columns of this thread's TMEM lane; the caller issues tmem_ld_wait() before using them
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------- packed f32x2 math (FFMA2 / FADD2)
// Two fp32 values in one 64-bit register (lo = first).  fma.rn.f32x2 issues ONE instruction for two FMAs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 16-byte shared-memory accesses as two packed pairs (shared-window addresses)
__device__ __forceinline__ void lds_2x2(uint32_t addr, f32x2& a, f32x2& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_2x2(uint32_t addr, f32x2 a, f32x2 b) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}

// bf16 split of a packed pair, 5 instructions: hi = bf16x2(x) (1), its two values back as fp32 (2), the exact
// residual r = x - hi with one packed FMA (1), lo = bf16x2(r) (1).
//   split_pair_rn      : hi rounded to nearest (|r| <= 2^-9 |x|), no activation
//   split_pair_rz_relu : y = max(x, 0) folded in: hi = RZ(relu(x)) via cvt.rz.relu, so r = x - hi keeps the sign of x
//                        (r >= 0 for x >= 0, r = x < 0 for x < 0) and lo = RN(relu(r)) -- no separate max
__device__ __forceinline__ void split_pair_rn(f32x2 x, uint32_t& hi, uint32_t& lo) {
  float x0, x1, r0, r1;
  unpack2(x, x0, x1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const f32x2 hf = pack2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  const f32x2 neg1 = 0xbf800000bf800000ull;                   // {-1.0f, -1.0f}
  unpack2(fma2(hf, neg1, x), r0, r1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void split_pair_rz_relu(f32x2 x, uint32_t& hi, uint32_t& lo) {
  float x0, x1, r0, r1;
  unpack2(x, x0, x1);
  asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const f32x2 hf = pack2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  const f32x2 neg1 = 0xbf800000bf800000ull;
  unpack2(fma2(hf, neg1, x), r0, r1);
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

// UMMA instruction descriptor: bf16 A/B (K-major), fp32 accumulate, M x N  (cute mma_sm100_desc.hpp:412-434)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4)                       // c_format  = F32
         | (1u << 7)                     // a_format  = BF16
         | (1u << 10)                    // b_format  = BF16
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}
// UMMA shared-memory descriptor, K-major SWIZZLE_128B (cute mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units   [0,14)
  d |= (uint64_t)1 << 16;                                // leading byte offset (unused)   [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset = 1024 B    [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)    [46,48)
  d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B       [61,64)
  return d;
}
// advance a K-major SW128 descriptor by `ksteps` MMA K-steps (16 bf16 = 32 bytes each)
__device__ __forceinline__ uint64_t sdesc_advance_k(uint64_t desc, int ksteps) { return desc + (uint64_t)(ksteps * 2); }

// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a K-major SW128 operand image
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// ------------------------------------------------------------------------- bf16 x3 split
// x ~ b0 + b1 with b0 = bf16(x), b1 = bf16(x - b0); two values packed per 32-bit word (lo = first)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float r0 = x0 - __low2float(h), r1 = x1 - __high2float(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// splits 8 consecutive K values and stores them as one 16-byte chunk into each operand image
__device__ __forceinline__ void split_store8(const float (&v)[8], uint8_t* img_hi, uint8_t* img_lo, uint32_t off) {
  uint4 h, l;
  split2(v[0], v[1], h.x, l.x);
  split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z);
  split2(v[6], v[7], h.w, l.w);
  *reinterpret_cast<uint4*>(img_hi + off) = h;
  *reinterpret_cast<uint4*>(img_lo + off) = l;
}

}  // namespace tc
}  // namespace wekws
