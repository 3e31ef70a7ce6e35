// Microbenchmark: how fast can ONE CTA pull a weight stream out of L2 with cp.async.bulk into a shared-memory ring?
// (sizing input for the weight-streaming GRU kernel: 786 KB per step and CTA).  Prints GB/s per CTA for a few grid sizes
// and chunk sizes.   nvcc -gencode arch=compute_100a,code=sm_100a -o build/l2_stream_probe tests/native/l2_stream_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t par) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int CHUNK, int SLOTS>
__global__ void probe(const uint8_t* src, int total_bytes, int reps, float* sink) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ uint64_t full[SLOTS];
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nchunk = total_bytes / CHUNK;
  if (threadIdx.x == 0) {
    // the consumer frees a slot as soon as it lands (no compute): measures the pure ingest rate with SLOTS chunks in flight
    int issued = 0, done = 0;
    const int total = nchunk * reps;
    while (done < total) {
      while (issued < total && issued - done < SLOTS) {
        const int s = issued % SLOTS;
        expect_tx(&full[s], CHUNK);
        bulk(ring + s * CHUNK, src + (size_t)(issued % nchunk) * CHUNK, CHUNK, &full[s]);
        ++issued;
      }
      wait(&full[done % SLOTS], (done / SLOTS) & 1);
      ++done;
    }
    sink[blockIdx.x] = ring[0];
  }
}

template <int CHUNK, int SLOTS>
void run(const uint8_t* d, int total, int grid, float* sink) {
  cudaFuncSetAttribute(probe<CHUNK, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHUNK * SLOTS);
  const int reps = 50;
  probe<CHUNK, SLOTS><<<grid, 32, CHUNK * SLOTS>>>(d, total, 2, sink);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<CHUNK, SLOTS><<<grid, 32, CHUNK * SLOTS>>>(d, total, reps, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per_cta = (double)total * reps / (ms * 1e-3) / 1e9;
  printf("{\"chunk\": %d, \"slots\": %d, \"grid\": %d, \"GBps_per_cta\": %.1f, \"TBps_total\": %.2f, \"us_per_786KB\": %.2f, \"err\": \"%s\"}\n", CHUNK, SLOTS,
         grid, per_cta, per_cta * grid / 1e3, 786432.0 / (per_cta * 1e3), cudaGetErrorString(cudaGetLastError()));
}

// one bulk copy at a time: issue -> complete latency in cycles (clock64), for a few sizes
__global__ void latency_probe(const uint8_t* src, long long* out) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t par = 0;
    int o = 0;
    for (int bytes : {1024, 4096, 16384, 32768, 65536}) {
      long long best = 1LL << 60;
      for (int rep = 0; rep < 20; ++rep) {
        const long long t0 = clock64();
        expect_tx(&bar, bytes);
        bulk(ring, src, bytes, &bar);
        wait(&bar, par);
        const long long t1 = clock64();
        par ^= 1;
        if (t1 - t0 < best) best = t1 - t0;
      }
      out[o++] = best;
    }
  }
}

// NTHR independent rings (one issuing thread per warp), to see whether the per-copy cost is per thread or per SM
template <int CHUNK, int SLOTS, int NTHR>
__global__ void probe_mt(const uint8_t* src, int total_bytes, int reps, float* sink) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ uint64_t full[NTHR][SLOTS];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < SLOTS; ++i) mbar_init(&full[w][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nchunk = total_bytes / CHUNK / NTHR;       // each thread streams its own share
  if ((threadIdx.x & 31) == 0) {
    uint8_t* my = ring + (size_t)w * SLOTS * CHUNK;
    const uint8_t* s0 = src + (size_t)w * nchunk * CHUNK;
    int issued = 0, done = 0, ci = 0;
    const int total = nchunk * reps;
    while (done < total) {
      while (issued < total && issued - done < SLOTS) {
        const int s = issued % SLOTS;
        expect_tx(&full[w][s], CHUNK);
        bulk(my + s * CHUNK, s0 + (size_t)ci * CHUNK, CHUNK, &full[w][s]);
        ++issued;
        ci = ci + 1 == nchunk ? 0 : ci + 1;
      }
      wait(&full[w][done % SLOTS], (done / SLOTS) & 1);
      ++done;
    }
    sink[blockIdx.x] = my[0];
  }
}

template <int CHUNK, int SLOTS, int NTHR>
void run_mt(const uint8_t* d, int total, int grid, float* sink) {
  cudaFuncSetAttribute(probe_mt<CHUNK, SLOTS, NTHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHUNK * SLOTS * NTHR);
  const int reps = 50;
  probe_mt<CHUNK, SLOTS, NTHR><<<grid, 32 * NTHR, CHUNK * SLOTS * NTHR>>>(d, total, 2, sink);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe_mt<CHUNK, SLOTS, NTHR><<<grid, 32 * NTHR, CHUNK * SLOTS * NTHR>>>(d, total, reps, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per_cta = (double)total * reps / (ms * 1e-3) / 1e9;
  printf("{\"chunk\": %d, \"slots\": %d, \"threads\": %d, \"grid\": %d, \"GBps_per_cta\": %.1f, \"us_per_786KB\": %.2f, \"err\": \"%s\"}\n", CHUNK,
         SLOTS, NTHR, grid, per_cta, 786432.0 / (per_cta * 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int total = 786432;
  uint8_t* d; float* sink;
  cudaMalloc(&d, total); cudaMemset(d, 1, total); cudaMalloc(&sink, 4096);
  {
    long long* dl; long long hl[5];
    cudaMalloc(&dl, 64);
    cudaFuncSetAttribute(latency_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    latency_probe<<<1, 32, 65536>>>(d, dl);
    cudaMemcpy(hl, dl, 40, cudaMemcpyDeviceToHost);
    printf("{\"bulk_copy_latency_cycles\": {\"1K\": %lld, \"4K\": %lld, \"16K\": %lld, \"32K\": %lld, \"64K\": %lld}}\n", hl[0], hl[1], hl[2], hl[3], hl[4]);
  }
  for (int grid : {1, 8}) {
    run_mt<16384, 2, 2>(d, total, grid, sink);
    run_mt<16384, 2, 4>(d, total, grid, sink);
    run_mt<16384, 1, 8>(d, total, grid, sink);
    run_mt<8192, 2, 8>(d, total, grid, sink);
    run_mt<4096, 4, 8>(d, total, grid, sink);
    run_mt<65536, 2, 1>(d, total, grid, sink);
  }
  for (int grid : {1, 148}) {
    run<16384, 4>(d, total, grid, sink);
    run<16384, 8>(d, total, grid, sink);
    run<32768, 4>(d, total, grid, sink);
    run<8192, 16>(d, total, grid, sink);
  }
  return 0;
}
