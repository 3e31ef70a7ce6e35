// See keyword_spotting_b200.h.  Host C++ over the C ABI; links libwekws_b200.so and the CUDA runtime.
#include "keyword_spotting_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>

#include "wekws_b200.h"

namespace wekws {

namespace {

[[noreturn]] void Fatal(const char* what, const char* detail) {
  fprintf(stderr, "wekws_b200 KeywordSpotting: %s: %s\n", what, detail ? detail : "");
  exit(-1);                                            // the reference's CHECK/LOG(FATAL) behaviour (utils/log.h:51-79)
}
void CudaOk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) Fatal(what, cudaGetErrorString(e));
}
void AbiOk(int rc, const char* what) {
  if (rc != WEKWS_OK) Fatal(what, wekws_last_error());
}
template <typename T>
T ReadPod(std::ifstream& is, const char* what) {
  T v;
  is.read(reinterpret_cast<char*>(&v), sizeof(T));
  if (!is) Fatal("truncated model file", what);
  return v;
}

}  // namespace

// .wkb layout (little endian): "WKB1", int32 version = 2, int32 nconfig, wekws_model_config (nconfig x int32; missing
// trailing fields of an older, shorter config block are zero), int32 ntensors, then per tensor: int32 name_len, name
// bytes, int64 numel, numel x float32.
KeywordSpotting::KeywordSpotting(const std::string& model_path) {
  std::ifstream is(model_path, std::ios::binary);
  if (!is) Fatal("cannot open model file", model_path.c_str());
  char magic[4];
  is.read(magic, 4);
  if (!is || memcmp(magic, "WKB1", 4) != 0) Fatal("not a wekws_b200 native model file (.wkb)", model_path.c_str());
  if (ReadPod<int32_t>(is, "version") != 2) Fatal("unsupported .wkb version", model_path.c_str());
  const int32_t nconfig = ReadPod<int32_t>(is, "config length");
  if (nconfig < 10 || nconfig > 64) Fatal("corrupt config block", model_path.c_str());
  wekws_model_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  for (int32_t i = 0; i < nconfig; ++i) {
    const int32_t v = ReadPod<int32_t>(is, "config");
    if (i < (int32_t)(sizeof(cfg) / sizeof(int32_t))) reinterpret_cast<int32_t*>(&cfg)[i] = v;
  }
  AbiOk(wekws_model_create(&cfg, &model_), "wekws_model_create");
  const int32_t ntensors = ReadPod<int32_t>(is, "tensor count");
  std::vector<float> buf;
  for (int32_t i = 0; i < ntensors; ++i) {
    const int32_t len = ReadPod<int32_t>(is, "name length");
    if (len <= 0 || len > 4096) Fatal("corrupt tensor name", model_path.c_str());
    std::string name(len, '\0');
    is.read(&name[0], len);
    const int64_t numel = ReadPod<int64_t>(is, "numel");
    if (!is || numel < 0 || numel > (1ll << 31)) Fatal("corrupt tensor header", name.c_str());
    buf.resize(numel);
    is.read(reinterpret_cast<char*>(buf.data()), numel * sizeof(float));
    if (!is) Fatal("truncated tensor data", name.c_str());
    AbiOk(wekws_model_set_tensor(model_, name.c_str(), buf.data(), numel), name.c_str());
  }
  AbiOk(wekws_model_finalize(model_), "wekws_model_finalize");      // folds BatchNorm, packs, uploads
  idim_ = cfg.idim;
  odim_ = cfg.odim;
  cache_dim_ = cfg.backbone == WEKWS_BACKBONE_FSMN ? cfg.fsmn_proj_dim : cfg.hdim;
  cache_len_ = wekws_model_padding(model_);
  // conv backbones: (1, hidden_dim, padding); GRU: (num_layers, 1, hidden_dim); FSMN: (1, proj_dim, padding, num_layers)
  cache_floats_ = cfg.backbone == WEKWS_BACKBONE_GRU    ? (long long)cfg.num_layers * cfg.hdim
                  : cfg.backbone == WEKWS_BACKBONE_FSMN ? (long long)cfg.fsmn_proj_dim * cache_len_ * cfg.num_layers
                                                        : (long long)cfg.hdim * cache_len_;
  cudaStream_t st;
  CudaOk(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
  stream_ = st;
  CudaOk(cudaMalloc(reinterpret_cast<void**>(&d_cache_), (cache_floats_ > 0 ? cache_floats_ : 1) * sizeof(float)), "cudaMalloc cache");
  printf("Kws Model Info:\n\tcache_dim: %d\n\tcache_len: %d\n", cache_dim_, cache_len_);
  Reset();
}

KeywordSpotting::~KeywordSpotting() {
  if (stream_) cudaStreamSynchronize(static_cast<cudaStream_t>(stream_));
  cudaFree(d_in_); cudaFree(d_out_); cudaFree(d_cache_);
  cudaFreeHost(h_in_); cudaFreeHost(h_out_);
  if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
  wekws_model_destroy(model_);
}

void KeywordSpotting::Reset() { started_ = false; }

void KeywordSpotting::Reserve(int num_frames) {
  if (num_frames <= capacity_) return;
  cudaStreamSynchronize(static_cast<cudaStream_t>(stream_));
  cudaFree(d_in_); cudaFree(d_out_); cudaFreeHost(h_in_); cudaFreeHost(h_out_);
  capacity_ = num_frames < 64 ? 64 : 2 * num_frames;
  CudaOk(cudaMalloc(reinterpret_cast<void**>(&d_in_), (size_t)capacity_ * idim_ * sizeof(float)), "cudaMalloc input");
  CudaOk(cudaMalloc(reinterpret_cast<void**>(&d_out_), (size_t)capacity_ * odim_ * sizeof(float)), "cudaMalloc output");
  CudaOk(cudaMallocHost(reinterpret_cast<void**>(&h_in_), (size_t)capacity_ * idim_ * sizeof(float)), "cudaMallocHost input");
  CudaOk(cudaMallocHost(reinterpret_cast<void**>(&h_out_), (size_t)capacity_ * odim_ * sizeof(float)), "cudaMallocHost output");
}

void KeywordSpotting::Forward(const std::vector<std::vector<float>>& feats, std::vector<std::vector<float>>* prob) {
  prob->clear();
  if (feats.empty()) return;
  const int num_frames = static_cast<int>(feats.size());
  Reserve(num_frames);
  for (int i = 0; i < num_frames; ++i) {
    if (static_cast<int>(feats[i].size()) != idim_) Fatal("feature dimension does not match the model", "Forward");
    memcpy(h_in_ + (size_t)i * idim_, feats[i].data(), idim_ * sizeof(float));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CudaOk(cudaMemcpyAsync(d_in_, h_in_, (size_t)num_frames * idim_ * sizeof(float), cudaMemcpyHostToDevice, st), "H2D");
  // the cache is updated in place: every slice is read before it is overwritten (include/wekws_b200.h)
  AbiOk(wekws_model_forward(model_, d_in_, started_ ? d_cache_ : nullptr, d_out_, d_cache_, 1, num_frames, 0, st),
        "wekws_model_forward");
  started_ = true;
  CudaOk(cudaMemcpyAsync(h_out_, d_out_, (size_t)num_frames * odim_ * sizeof(float), cudaMemcpyDeviceToHost, st), "D2H");
  CudaOk(cudaStreamSynchronize(st), "cudaStreamSynchronize");
  prob->resize(num_frames);
  for (int i = 0; i < num_frames; ++i) (*prob)[i].assign(h_out_ + (size_t)i * odim_, h_out_ + (size_t)(i + 1) * odim_);
}

}  // namespace wekws
