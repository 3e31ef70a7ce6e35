// Fused KWSModel.forward for the GRU backbone (reference wekws/model/kws_model.py:128-133:
// torch.nn.GRU(hdim, hdim, num_layers, batch_first=True) between LinearSubsampling1
// subsampling.py:53-57 and LinearClassifier classifier.py:63-67).
//
// PyTorch gate order r, z, n:  r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise,
// n = tanh(W_in x + b_in + r * (W_hn h + b_hn)),  h' = (1 - z) * n + z * h.
//
// A CTA owns S streams for the whole chunk of T frames; hidden state, layer input and gate
// pre-activations stay in shared memory across time steps and layers, so per step only the
// idim input features are read and odim posteriors written; the (L,B,H) cache is read and
// written once per call.  One thread per gate row (3H = 384 threads): the transposed weights
// [k][3H] are streamed from L2 with coalesced 128-byte warp loads and each weight is reused
// for the S streams from registers.
#include "common.cuh"
#include "gru.h"

namespace wekws {

namespace {

template <int S>
__global__ void __launch_bounds__(384, 1) gru_kernel(const GruArgs a) {
  constexpr int H = 128, G = 3 * H;
  extern __shared__ __align__(16) float sm[];
  float* xin = sm;                    // [S][H]   current layer input
  float* hst = xin + S * H;           // [L][S][H]
  float* gi = hst + a.L * S * H;      // [S][G]
  float* gh = gi + S * G;             // [S][G]
  float* fin = gh + S * G;            // [S][idimP]
  const int tid = threadIdx.x;
  const int idim = a.idim, idimP = (idim + 3) & ~3;
  const float* vec = a.vec;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int b0 = tile * S;
    const int Sv = min(S, a.B - b0);
    __syncthreads();
    for (int i = tid; i < a.L * S * H; i += G) {
      const int l = i / (S * H), rem = i - l * S * H, s = rem / H, j = rem - s * H;
      float v = 0.f;
      if (a.in_cache != nullptr && s < Sv) v = a.in_cache[((size_t)l * a.B + b0 + s) * H + j];
      hst[i] = v;
    }
    for (int t = 0; t < a.T; ++t) {
      // features of this step (+CMVN)
      for (int i = tid; i < S * idimP; i += G) {
        const int s = i / idimP, k = i - s * idimP;
        float v = 0.f;
        if (s < Sv && k < idim) {
          v = __ldg(a.feats + ((size_t)(b0 + s) * a.T + t) * idim + k);
          if (a.has_cmvn) v = (v - __ldg(vec + a.v_mean + k)) * __ldg(vec + a.v_istd + k);
        }
        fin[i] = v;
      }
      __syncthreads();
      // preprocessing Linear + ReLU: threads j < H
      if (tid < H) {
        float acc[S];
        const float bp = __ldg(vec + a.v_bp + tid);
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = bp;
        const float* wp = vec + a.v_wp + tid;          // WpT[k][H]
        for (int k = 0; k < idim; ++k) {
          const float w = __ldg(wp + k * H);
#pragma unroll
          for (int s = 0; s < S; ++s) acc[s] = fmaf(w, fin[s * idimP + k], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) xin[s * H + tid] = fmaxf(acc[s], 0.f);
      }
      __syncthreads();
      for (int l = 0; l < a.L; ++l) {
        const float* wih = vec + a.v_layers + l * a.v_layer_stride + tid;   // WihT[k][G]
        const float* whh = wih + H * G;                                      // WhhT[k][G]
        const float* hl = hst + l * S * H;
        float ai[S], ah[S];
        {
          const float bi = __ldg(vec + a.v_layers + l * a.v_layer_stride + 2 * H * G + tid);
          const float bh = __ldg(vec + a.v_layers + l * a.v_layer_stride + 2 * H * G + G + tid);
#pragma unroll
          for (int s = 0; s < S; ++s) { ai[s] = bi; ah[s] = bh; }
        }
#pragma unroll 2
        for (int k = 0; k < H; k += 4) {
          float wi[4], wh[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { wi[u] = __ldg(wih + (k + u) * G); wh[u] = __ldg(whh + (k + u) * G); }
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float4 x4 = *reinterpret_cast<const float4*>(xin + s * H + k);
            const float4 h4 = *reinterpret_cast<const float4*>(hl + s * H + k);
            ai[s] = fmaf(wi[0], x4.x, ai[s]); ai[s] = fmaf(wi[1], x4.y, ai[s]);
            ai[s] = fmaf(wi[2], x4.z, ai[s]); ai[s] = fmaf(wi[3], x4.w, ai[s]);
            ah[s] = fmaf(wh[0], h4.x, ah[s]); ah[s] = fmaf(wh[1], h4.y, ah[s]);
            ah[s] = fmaf(wh[2], h4.z, ah[s]); ah[s] = fmaf(wh[3], h4.w, ah[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) { gi[s * G + tid] = ai[s]; gh[s * G + tid] = ah[s]; }
        __syncthreads();
        for (int i = tid; i < S * H; i += G) {
          const int s = i / H, j = i - s * H;
          const float r = sigmoidf_acc(gi[s * G + j] + gh[s * G + j]);
          const float z = sigmoidf_acc(gi[s * G + H + j] + gh[s * G + H + j]);
          const float n = tanhf(gi[s * G + 2 * H + j] + r * gh[s * G + 2 * H + j]);
          const float hp = hst[(l * S + s) * H + j];
          const float hn = (1.f - z) * n + z * hp;
          hst[(l * S + s) * H + j] = hn;
          xin[s * H + j] = hn;
        }
        __syncthreads();
      }
      // classifier on the top layer's h_t
      for (int i = tid; i < Sv * a.odim; i += G) {
        const int s = i / a.odim, j = i - s * a.odim;
        float acc = __ldg(vec + a.v_bc + j);
        const float* wc = vec + a.v_wc + j;            // WcT[k][odim]
        for (int k = 0; k < H; ++k) acc = fmaf(__ldg(wc + k * a.odim), xin[s * H + k], acc);
        if (a.act == WEKWS_ACT_SIGMOID) acc = sigmoidf_acc(acc);
        a.out[((size_t)(b0 + s) * a.T + t) * a.odim + j] = acc;
      }
      // (next step's feature load only touches `fin`; the barrier after it orders xin reuse)
    }
    __syncthreads();
    for (int i = tid; i < a.L * Sv * H; i += G) {
      const int l = i / (Sv * H), rem = i - l * Sv * H, s = rem / H, j = rem - s * H;
      a.out_cache[((size_t)l * a.B + b0 + s) * H + j] = hst[(l * S + s) * H + j];
    }
  }
}

template <int S>
int launch_s(const GruArgs& a, cudaStream_t st) {
  GruArgs b = a;
  b.n_tiles = (a.B + S - 1) / S;
  const int idimP = (a.idim + 3) & ~3;
  const size_t smem = (size_t)(S * 128 + a.L * S * 128 + 2 * S * 384 + S * idimP) * sizeof(float);
  WEKWS_CUDA_OK(cudaFuncSetAttribute(gru_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = device_sm_count();
  const int grid = b.n_tiles < sms ? b.n_tiles : sms;
  gru_kernel<S><<<grid, 384, smem, st>>>(b);
  return check_launch("gru_kernel");
}

}  // namespace

int gru_launch(const GruArgs& a, cudaStream_t st) {
  WEKWS_REQUIRE(a.H == 128, "GRU hidden_dim %d unsupported (128 only)", a.H);
  WEKWS_REQUIRE(a.L >= 1 && a.L <= 4, "GRU num_layers %d unsupported (1..4)", a.L);
  const int sms = device_sm_count();
  if (a.B <= sms) return launch_s<1>(a, st);
  if (a.B <= 2 * sms) return launch_s<2>(a, st);
  if (a.B <= 4 * sms) return launch_s<4>(a, st);
  return launch_s<8>(a, st);
}

}  // namespace wekws
