// Microbenchmark: dispatch rate of tcgen05.mma.cta_group::1.kind::f16 (bf16, K = 16 per instruction) for the operand
// forms and tile shapes the kernels use: A from shared memory (SS) or TMEM (TS), M = 128, N = 64 / 128 / 256, and
// M = 64 SS.  One thread issues `reps` MMAs on fixed (zeroed) operands and commits; cycles per MMA by clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o build/mma_rate_probe tests/native/mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../wekws_b200/csrc/tc_common.cuh"

using namespace wekws::tc;

// mode 0: SS, same descriptors every time; 1: SS, K-step descriptors cycling over a 4-step slab and 4 slabs (distinct
// shared-memory lines); 2: TS (A in TMEM columns 256..), B cycling as in 1
__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 163840 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint32_t a0 = smem_u32(base), b0 = smem_u32(base) + 65536;     // A: 4 slabs x 16 KB, B: up to 3 slabs x 32 KB
    const uint32_t bslab = (uint32_t)N * 128;
    // warm-up
    for (int i = 0; i < 16; ++i) umma_bf16(tmem, make_sdesc_sw128(a0), make_sdesc_sw128(b0), idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int k = i & 3, s = (i >> 2) & 3;
      const uint64_t da = make_sdesc_sw128(a0 + (mode ? s * 16384 : 0)) + (mode ? 2 * k : 0);
      const uint64_t db = make_sdesc_sw128(b0 + (mode ? (s % 3) * bslab : 0)) + (mode ? 2 * k : 0);
      if (mode == 2) umma_bf16_ts(tmem, tmem + 256 + 8 * k + 32 * (s & 1), db, idesc, 1);
      else umma_bf16(tmem, da, db, idesc, 1);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;       // issue loop
    out[1] = t2 - t0;       // until complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// NI issuing threads (lane 0 of warps 0..NI-1), each accumulating into its own 64 TMEM columns, SS form M=128 N=64;
// unrolled by 8 with the descriptors precomputed, so the loop body is 8 x (setp + mma)
__global__ void __launch_bounds__(128, 1) rate_mt_kernel(int NI, int reps, long long* out, int N = 64, int ts = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 163840 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if ((tid & 31) == 0 && warp < NI) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(base) + warp * 16384, b0 = smem_u32(base) + 65536 + warp * 8192;
    uint64_t da[8], db[8];
    for (int i = 0; i < 8; ++i) { da[i] = make_sdesc_sw128(a0) + 2 * (i & 3); db[i] = make_sdesc_sw128(b0) + 2 * (i & 3); }
    const uint32_t d = tmem + 64 * warp;
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 8) {
      if (ts) {
#pragma unroll
        for (int u = 0; u < 8; ++u) umma_bf16_ts(d, tmem + 256 + 8 * (u & 3), db[u], idesc, 1);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) umma_bf16(d, da[u], db[u], idesc, 1);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    const long long t2 = clock64();
    out[2 * warp] = t1 - t0;
    out[2 * warp + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840 + 1024);
  const int reps = 2048;
  struct Case { int M, N, mode; const char* name; };
  const Case cases[] = {{128, 64, 0, "SS fixed"}, {128, 64, 1, "SS"}, {128, 128, 1, "SS"}, {128, 256, 1, "SS"}, {64, 64, 1, "SS"},
                        {64, 128, 1, "SS"}, {64, 256, 1, "SS"}, {128, 64, 2, "TS"}, {128, 128, 2, "TS"}, {128, 256, 2, "TS"},
                        {128, 48, 2, "TS"}, {128, 16, 2, "TS"}, {128, 16, 1, "SS"}, {128, 32, 1, "SS"}};
  for (const Case& c : cases) {
    rate_kernel<<<1, 128, 163840 + 1024>>>(c.M, c.N, c.mode, reps, d);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("{\"form\": \"%s\", \"M\": %d, \"N\": %d, \"cycles_per_mma_issue\": %.1f, \"cycles_per_mma_complete\": %.1f, \"floor_MN_over_256\": %.0f, \"err\": \"%s\"}\n",
           c.name, c.M, c.N, (double)h[0] / reps, (double)h[1] / reps, (c.M < 128 ? 128.0 : c.M) * c.N / 256.0, cudaGetErrorString(e));
  }
  long long* d8;
  cudaMalloc(&d8, 64);
  cudaFuncSetAttribute(rate_mt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840 + 1024);
  for (int ni : {1, 2, 4}) {
    rate_mt_kernel<<<1, 128, 163840 + 1024>>>(ni, reps, d8);
    long long h[8] = {0};
    cudaError_t e = cudaMemcpy(h, d8, 64, cudaMemcpyDeviceToHost);
    printf("{\"form\": \"SS M128 N64 unrolled\", \"issuers\": %d, \"cycles_per_mma_issue\": %.1f, \"cycles_per_mma_complete_thread0\": %.1f, \"aggregate_cycles_per_mma\": %.1f, \"err\": \"%s\"}\n",
           ni, (double)h[0] / reps, (double)h[1] / reps, (double)h[1] / reps / ni, cudaGetErrorString(e));
  }
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {16, 32, 64, 96, 128, 192, 256}) {
      rate_mt_kernel<<<1, 128, 163840 + 1024>>>(1, reps, d8, n, ts);
      long long h[8] = {0};
      cudaError_t e = cudaMemcpy(h, d8, 64, cudaMemcpyDeviceToHost);
      printf("{\"form\": \"%s M128 unrolled\", \"N\": %d, \"cycles_per_mma\": %.1f, \"floor_MN_over_256\": %.0f, \"err\": \"%s\"}\n", ts ? "TS" : "SS", n,
             (double)h[1] / reps, 128.0 * n / 256.0, cudaGetErrorString(e));
    }
  return 0;
}
