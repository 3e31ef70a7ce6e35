// GPU probe for the tcgen05 building blocks in wekws_b200/csrc/tc_common.cuh (test code only):
// D[128][64] = A[128][K] * W[64][K]^T with the bf16x3 operand split, K = 64 (one SW128 atom) or
// K = 80 (second atom holds K columns 64..79).  Built by tests/test_tc_probe.py with nvcc.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../wekws_b200/csrc/tc_common.cuh"

using namespace wekws::tc;

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       float* __restrict__ D, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
  uint8_t* a_hi[2] = {base, base + 32768};                 // [atom][128 rows * 128 B]
  uint8_t* a_lo[2] = {base + 16384, base + 49152};
  uint8_t* w_hi[2] = {base + 65536, base + 65536 + 16384};
  uint8_t* w_lo[2] = {base + 65536 + 8192, base + 65536 + 24576};
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int natoms = (K + 63) / 64;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  // operands: thread = row; 16-byte chunks of 8 K values
  for (int at = 0; at < natoms; ++at) {
    for (int ch = 0; ch < 8; ++ch) {
      const int k0 = at * 64 + ch * 8;
      if (k0 >= ((K + 15) & ~15)) break;     // zero-fill up to the last 16-wide K step
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (k0 + i < K) ? A[tid * K + k0 + i] : 0.f;
      split_store8(v, a_hi[at], a_lo[at], sw128_offset(tid, ch));
      if (tid < 64) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (k0 + i < K) ? W[tid * K + k0 + i] : 0.f;
        split_store8(v, w_hi[at], w_lo[at], sw128_offset(tid, ch));
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64);
    uint32_t acc = 0;
    for (int pass = 0; pass < 3; ++pass) {
      for (int at = 0; at < natoms; ++at) {
        const uint64_t da = make_sdesc_sw128(smem_u32(pass == 1 ? a_lo[at] : a_hi[at]));
        const uint64_t db = make_sdesc_sw128(smem_u32(pass == 2 ? w_lo[at] : w_hi[at]));
        const int ksteps = (min(K - at * 64, 64) + 15) / 16;
        for (int k = 0; k < ksteps; ++k) {
          umma_bf16(tmem, sdesc_advance_k(da, k), sdesc_advance_k(db, k), idesc, acc);
          acc = 1;
        }
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int g = 0; g < 4; ++g) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + 16 * g, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[tid * 64 + 16 * g + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// Same GEMM (K = 64) with the A operand staged in TMEM by tcgen05.st (hi at columns 64.., lo at 96..).
__global__ void __launch_bounds__(128, 1) probe_ts_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                          float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
  uint8_t* w_hi = base;
  uint8_t* w_lo = base + 8192;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = 64;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
  for (int ch = 0; ch < 8; ++ch) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = A[tid * K + ch * 8 + i];
    uint32_t h[4], l[4];
    split2(v[0], v[1], h[0], l[0]); split2(v[2], v[3], h[1], l[1]);
    split2(v[4], v[5], h[2], l[2]); split2(v[6], v[7], h[3], l[3]);
    tmem_st4(lane_base + 64 + 4 * ch, h);
    tmem_st4(lane_base + 96 + 4 * ch, l);
    if (tid < 64) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = W[tid * K + ch * 8 + i];
      split_store8(v, w_hi, w_lo, sw128_offset(tid, ch));
    }
  }
  tmem_st_wait();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64);
    const uint64_t dwh = make_sdesc_sw128(smem_u32(w_hi)), dwl = make_sdesc_sw128(smem_u32(w_lo));
    uint32_t acc = 0;
    for (int k = 0; k < 4; ++k) { umma_bf16_ts(tmem, tmem + 64 + 8 * k, sdesc_advance_k(dwh, k), idesc, acc); acc = 1; }
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem, tmem + 96 + 8 * k, sdesc_advance_k(dwh, k), idesc, 1);
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem, tmem + 64 + 8 * k, sdesc_advance_k(dwl, k), idesc, 1);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int g = 0; g < 4; ++g) {
    float v[16];
    tmem_ld16(lane_base + 16 * g, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[tid * 64 + 16 * g + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

extern "C" __attribute__((visibility("default"))) int tc_probe_ts_run(const float* dA, const float* dW, float* dD) {
  const int smem = 1024 + 16384;
  probe_ts_kernel<<<1, 128, smem>>>(dA, dW, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "tc_probe_ts: %s\n", cudaGetErrorString(e)); return -2; }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tc_probe_run(const float* dA, const float* dW, float* dD, int K) {
  const int smem = 1024 + 65536 + 32768;
  cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return -1;
  probe_kernel<<<1, 128, smem>>>(dA, dW, dD, K);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "tc_probe: %s\n", cudaGetErrorString(e)); return -2; }
  return 0;
}
