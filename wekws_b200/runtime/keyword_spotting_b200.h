// Native (no Python, no ONNX Runtime) stand-in for the reference runtime's keyword spotter,
// runtime/core/kws/keyword_spotting.h:26-55: same class name, namespace and public surface
//   KeywordSpotting(model_path) / Reset() / InitEngineThreads(n) / Forward(feats, &prob)
// so runtime/core/bin/kws_main.cc:43-61 and stream_kws_main.cc compile against it unchanged.  Instead of an
// ONNX file, model_path names a ".wkb" weight file written by wekws_b200.export_native(model, path): the model
// configuration plus every tensor of the reference state_dict under its reference key.  The network itself is
// the fused sm_100a kernels behind the C ABI (include/wekws_b200.h); the ONNX contract of export_onnx.py:55-77
// (inputs input,cache -> outputs output,r_cache, B = 1, cache_dim = hidden_dim, cache_len = backbone.padding)
// maps one to one onto wekws_model_forward(d_feats, d_cache -> d_out, d_cache).
#ifndef WEKWS_B200_RUNTIME_KEYWORD_SPOTTING_B200_H_
#define WEKWS_B200_RUNTIME_KEYWORD_SPOTTING_B200_H_

#include <string>
#include <vector>

struct wekws_model;

namespace wekws {

class KeywordSpotting {
 public:
  explicit KeywordSpotting(const std::string& model_path);
  ~KeywordSpotting();
  KeywordSpotting(const KeywordSpotting&) = delete;
  KeywordSpotting& operator=(const KeywordSpotting&) = delete;

  // Call reset if keyword is detected: the next Forward starts a new stream (zero cache,
  // keyword_spotting.cc:47-54)
  void Reset();

  // Kept for source compatibility (the reference sets ONNX Runtime thread counts here); nothing to do on a GPU.
  static void InitEngineThreads(int num_threads) { (void)num_threads; }

  // feats: num_frames x feature_dim.  prob: num_frames x output_dim posteriors of this chunk; the streaming
  // cache is carried between calls (keyword_spotting.cc:56-95).  Exits with a message on error, as the
  // reference does through LOG(FATAL).
  void Forward(const std::vector<std::vector<float>>& feats, std::vector<std::vector<float>>* prob);

  int feature_dim() const { return idim_; }
  int output_dim() const { return odim_; }
  int cache_dim() const { return cache_dim_; }   // ONNX metadata "cache_dim" (export_onnx.py:74)
  int cache_len() const { return cache_len_; }   // ONNX metadata "cache_len" (export_onnx.py:76)

 private:
  void Reserve(int num_frames);

  wekws_model* model_ = nullptr;
  void* stream_ = nullptr;
  int idim_ = 0, odim_ = 0, cache_dim_ = 0, cache_len_ = 0;
  long long cache_floats_ = 0;
  bool started_ = false;                 // false: next Forward passes a null cache (== zeros)
  float* d_in_ = nullptr;
  float* d_out_ = nullptr;
  float* d_cache_ = nullptr;
  float* h_in_ = nullptr;                // pinned staging buffers
  float* h_out_ = nullptr;
  int capacity_ = 0;                     // frames the buffers hold
};

}  // namespace wekws

#endif  // WEKWS_B200_RUNTIME_KEYWORD_SPOTTING_B200_H_
