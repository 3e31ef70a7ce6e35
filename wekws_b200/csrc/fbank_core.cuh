// Per-frame core of the Kaldi-compatible Fbank front-end, shared by fbank.cu (stand-alone kernel) and mdtc_tc.cu (raw
// PCM -> posterior in one launch): DC removal, pre-emphasis with replicate padding, window, 512-point real FFT as a
// 256-point complex FFT (radix 8 / 8 / 4 Stockham, first radix-8 in registers, two exchanges through padded per-warp
// shared buffers), untangle + power spectrum.  One warp per frame.  torchaudio kaldi.py:183-211, 616-618.
#pragma once
#include <cuda_runtime.h>

namespace wekws {
namespace fbcore {

constexpr int WIN = 400, SHIFT = 160, NFFT = 512, NBIN = 256;
constexpr int A_SZ = 264, B_SZ = 280;       // padded exchange buffers (floats)

__device__ __forceinline__ int piA(int i) { return i + (i >> 5); }
__device__ __forceinline__ int piB(int i) { return i + 8 * (i >> 6); }

__device__ __forceinline__ void cmul(float& re, float& im, float wr, float wi) {
  const float t = re * wr - im * wi;
  im = fmaf(re, wi, im * wr);
  re = t;
}

// In-place 8-point DFT (forward, e^{-2 pi i rk/8}) of (r[], i[]).
__device__ __forceinline__ void dft8(float (&r)[8], float (&i)[8]) {
  const float b0r = r[0] + r[4], b0i = i[0] + i[4], b1r = r[0] - r[4], b1i = i[0] - i[4];
  const float b2r = r[2] + r[6], b2i = i[2] + i[6], b3r = r[2] - r[6], b3i = i[2] - i[6];
  const float b4r = r[1] + r[5], b4i = i[1] + i[5], b5r = r[1] - r[5], b5i = i[1] - i[5];
  const float b6r = r[3] + r[7], b6i = i[3] + i[7], b7r = r[3] - r[7], b7i = i[3] - i[7];
  // c1 = b1 - i b3, c3 = b1 + i b3   (-i (x+iy) = y - ix)
  const float c0r = b0r + b2r, c0i = b0i + b2i, c2r = b0r - b2r, c2i = b0i - b2i;
  const float c1r = b1r + b3i, c1i = b1i - b3r, c3r = b1r - b3i, c3i = b1i + b3r;
  const float c4r = b4r + b6r, c4i = b4i + b6i, c6r = b4r - b6r, c6i = b4i - b6i;
  const float c5r = b5r + b7i, c5i = b5i - b7r, c7r = b5r - b7i, c7i = b5i + b7r;
  const float h = 0.70710678118654752440f;
  // w1*c5, w1 = (1 - i)/sqrt2 ; w3*c7, w3 = (-1 - i)/sqrt2
  const float t5r = h * (c5r + c5i), t5i = h * (c5i - c5r);
  const float t7r = h * (c7i - c7r), t7i = -h * (c7r + c7i);
  r[0] = c0r + c4r; i[0] = c0i + c4i; r[4] = c0r - c4r; i[4] = c0i - c4i;
  r[1] = c1r + t5r; i[1] = c1i + t5i; r[5] = c1r - t5r; i[5] = c1i - t5i;
  r[2] = c2r + c6i; i[2] = c2i - c6r; r[6] = c2r - c6i; i[6] = c2i + c6r;   // -i c6
  r[3] = c3r + t7r; i[3] = c3i + t7i; r[7] = c3r - t7r; i[7] = c3i - t7i;
}


// lane-constant twiddles: pass 1 W_256^(lane k), pass 2 W_32^(p k) = W_256^(8 p k), p = lane / 8
struct LaneTwiddles {
  float t1r[8], t1i[8], t2r[8], t2i[8];
  __device__ __forceinline__ void load(const float2* __restrict__ tw256, int lane) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float2 w1 = tw256[lane * k];
      const float2 w2 = tw256[8 * (lane >> 3) * k];
      t1r[k] = w1.x; t1i[k] = w1.y; t2r[k] = w2.x; t2i[k] = w2.y;
    }
  }
};

// Power spectrum of one frame -> pw[0..255].  `s(n)`: sample n (0..399) of the frame as float (int16 scale).
// Ar/Ai: A_SZ floats each, Br/Bi: B_SZ floats each, private to the warp; pw may be Br.  s_win: window as (even, odd) pairs; s_tw512:
// W_512^k, k < 256.  All 32 lanes participate.
template <typename Sample>
__device__ __forceinline__ void frame_power_spectrum(Sample s_, const float2* __restrict__ s_win,
                                                     const float2* __restrict__ s_tw512, const LaneTwiddles& tw, float* Ar,
                                                     float* Ai, float* Br, float* Bi, float* pw, float preemph, int remove_dc, int lane) {
  const float* t1r = tw.t1r; const float* t1i = tw.t1i; const float* t2r = tw.t2r; const float* t2i = tw.t2i;
    // ---- window: lane owns packed points m = lane + 32 i (even/odd sample pair 2m, 2m+1) ----
    float xa[7], xb[7], xc[7];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int m = lane + 32 * i;
      if (m < WIN / 2) {
        xb[i] = s_(2 * m); xc[i] = s_(2 * m + 1);
        xa[i] = m > 0 ? s_(2 * m - 1) : xb[i];          // replicate pad (kaldi.py:195)
        sum += xb[i] + xc[i];
      } else {
        xa[i] = xb[i] = xc[i] = 0.f;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = remove_dc ? sum / (float)WIN : 0.f;
    float zr[8], zi[8];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int m = lane + 32 * i;
      if (m < WIN / 2) {
        const float2 w = s_win[m];
        const float pa = xa[i] - mean, pb = xb[i] - mean, pc = xc[i] - mean;
        zr[i] = (pb - preemph * pa) * w.x;
        zi[i] = (pc - preemph * pb) * w.y;
      } else {
        zr[i] = 0.f; zi[i] = 0.f;
      }
    }
    zr[7] = 0.f; zi[7] = 0.f;

    // ---- pass 1: radix 8 over r (n=256, s=1) -> A[8p + k] * W_256^(pk) ----
    dft8(zr, zi);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k) cmul(zr[k], zi[k], t1r[k], t1i[k]);
      const int idx = piA(8 * lane + k);
      Ar[idx] = zr[k]; Ai[idx] = zi[k];
    }
    __syncwarp();
    // ---- pass 2: radix 8 (n=32, s=8): j = q + 8p reads A[j + 32r], writes B[q + 64p + 8k] ----
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int idx = piA(lane + 32 * r);
      zr[r] = Ar[idx]; zi[r] = Ai[idx];
    }
    dft8(zr, zi);
    {
      const int q = lane & 7, p = lane >> 3;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k) cmul(zr[k], zi[k], t2r[k], t2i[k]);
        const int idx = piB(q + 64 * p + 8 * k);
        Br[idx] = zr[k]; Bi[idx] = zi[k];
      }
    }
    __syncwarp();
    // ---- pass 3: radix 4 (n=4, s=64): q reads B[q + 64r], writes Z[q + 64k] into A ----
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int q = lane + 32 * hh;
      float r0 = Br[piB(q)], i0 = Bi[piB(q)];
      float r1 = Br[piB(q + 64)], i1 = Bi[piB(q + 64)];
      float r2 = Br[piB(q + 128)], i2 = Bi[piB(q + 128)];
      float r3 = Br[piB(q + 192)], i3 = Bi[piB(q + 192)];
      const float s0r = r0 + r2, s0i = i0 + i2, d0r = r0 - r2, d0i = i0 - i2;
      const float s1r = r1 + r3, s1i = i1 + i3, d1r = r1 - r3, d1i = i1 - i3;
      Ar[piA(q)] = s0r + s1r;        Ai[piA(q)] = s0i + s1i;
      Ar[piA(q + 64)] = d0r + d1i;   Ai[piA(q + 64)] = d0i - d1r;     // d0 - i d1
      Ar[piA(q + 128)] = s0r - s1r;  Ai[piA(q + 128)] = s0i - s1i;
      Ar[piA(q + 192)] = d0r - d1i;  Ai[piA(q + 192)] = d0i + d1r;    // d0 + i d1
    }
    __syncwarp();
    // ---- real-FFT untangle + power spectrum -> pw[0..255] ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      const int kn = (NBIN - k) & (NBIN - 1);
      const float ar = Ar[piA(k)], ai = Ai[piA(k)];
      const float cr = Ar[piA(kn)], ci = Ai[piA(kn)];
      const float er = 0.5f * (ar + cr), ei = 0.5f * (ai - ci);
      const float dr = 0.5f * (ar - cr), di = 0.5f * (ai + ci);
      const float2 w = s_tw512[k];
      const float p = w.x * dr - w.y * di, q = w.x * di + w.y * dr;
      const float xr = er + q, xi = ei - p;
      pw[k] = xr * xr + xi * xi;
    }
    __syncwarp();
}

}  // namespace fbcore
}  // namespace wekws
