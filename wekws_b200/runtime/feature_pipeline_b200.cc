// See feature_pipeline_b200.h.
#include "feature_pipeline_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <limits>

#include "wekws_b200.h"

namespace wenet {

namespace {

[[noreturn]] void Fatal(const char* what, const char* detail) {
  fprintf(stderr, "wekws_b200 FeaturePipeline: %s: %s\n", what, detail ? detail : "");
  exit(-1);
}
void CudaOk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) Fatal(what, cudaGetErrorString(e));
}
void AbiOk(int rc, const char* what) {
  if (rc != WEKWS_OK) Fatal(what, wekws_last_error());
}
int UpperPowerOfTwo(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}
double MelScale(double freq) { return 1127.0 * log(1.0 + freq / 700.0); }

}  // namespace

void FeaturePipelineConfig::Info() const {
  printf("feature pipeline config num_bins %d frame_length %d frame_shift %d\n", num_bins, frame_length, frame_shift);
}

FeaturePipeline::FeaturePipeline(const FeaturePipelineConfig& config) : config_(config), feature_dim_(config.num_bins) {
  const int win = config_.frame_length, n_fft = UpperPowerOfTwo(win), nbin = n_fft / 2;
  // Hamming window (frontend/fbank.h:90-96) and Kaldi's triangular mel banks between 20 Hz and Nyquist
  // (fbank.h:44-88; bins weighted by their position between the neighbouring centre frequencies in the mel domain)
  std::vector<float> window(win), mel((size_t)config_.num_bins * nbin, 0.f);
  const double kPi = 3.14159265358979323846;
  for (int i = 0; i < win; ++i) window[i] = (float)(0.54 - 0.46 * cos(2.0 * kPi * i / (win - 1)));
  const double low = MelScale(20.0), high = MelScale(0.5 * config_.sample_rate);
  const double delta = (high - low) / (config_.num_bins + 1), bin_width = (double)config_.sample_rate / n_fft;
  for (int b = 0; b < config_.num_bins; ++b) {
    const double left = low + b * delta, center = left + delta, right = center + delta;
    for (int k = 0; k < nbin; ++k) {
      const double m = MelScale(bin_width * k);
      if (m > left && m < right) mel[(size_t)b * nbin + k] = (float)(m <= center ? (m - left) / (center - left) : (right - m) / (right - center));
    }
  }
  wekws_fbank_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.sample_rate = config_.sample_rate; cfg.frame_length = win; cfg.frame_shift = config_.frame_shift; cfg.n_fft = n_fft;
  cfg.num_mel_bins = config_.num_bins; cfg.preemphasis = 0.97f; cfg.remove_dc = 1;
  cfg.log_floor = std::numeric_limits<float>::epsilon();
  AbiOk(wekws_fbank_create(&cfg, window.data(), mel.data(), &fbank_), "wekws_fbank_create");
  cudaStream_t st;
  CudaOk(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
  stream_ = st;
}

FeaturePipeline::~FeaturePipeline() {
  if (stream_) cudaStreamSynchronize(static_cast<cudaStream_t>(stream_));
  cudaFree(d_wav_); cudaFree(d_feat_); cudaFreeHost(h_feat_);
  if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
  wekws_fbank_destroy(fbank_);
}

void FeaturePipeline::AcceptWaveform(const std::vector<float>& wav) {
  // remainder of the previous call + new samples (feature_pipeline.cc:30-34)
  std::vector<float> waves;
  waves.reserve(remained_wav_.size() + wav.size());
  waves.insert(waves.end(), remained_wav_.begin(), remained_wav_.end());
  waves.insert(waves.end(), wav.begin(), wav.end());
  const int num_frames = (int)wekws_fbank_num_frames(fbank_, (int64_t)waves.size());
  if (num_frames > 0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (waves.size() > wav_capacity_) {
      cudaStreamSynchronize(st);
      cudaFree(d_wav_);
      wav_capacity_ = 2 * waves.size();
      CudaOk(cudaMalloc(reinterpret_cast<void**>(&d_wav_), wav_capacity_ * sizeof(float)), "cudaMalloc wav");
    }
    const size_t nfeat = (size_t)num_frames * feature_dim_;
    if (nfeat > feat_capacity_) {
      cudaStreamSynchronize(st);
      cudaFree(d_feat_); cudaFreeHost(h_feat_);
      feat_capacity_ = 2 * nfeat;
      CudaOk(cudaMalloc(reinterpret_cast<void**>(&d_feat_), feat_capacity_ * sizeof(float)), "cudaMalloc feats");
      CudaOk(cudaMallocHost(reinterpret_cast<void**>(&h_feat_), feat_capacity_ * sizeof(float)), "cudaMallocHost feats");
    }
    CudaOk(cudaMemcpyAsync(d_wav_, waves.data(), waves.size() * sizeof(float), cudaMemcpyHostToDevice, st), "H2D");
    AbiOk(wekws_fbank_forward(fbank_, d_wav_, WEKWS_PCM_F32, 1, (int64_t)waves.size(), (int64_t)waves.size(), nullptr,
                              nullptr, nullptr, d_feat_, num_frames, st), "wekws_fbank_forward");
    CudaOk(cudaMemcpyAsync(h_feat_, d_feat_, nfeat * sizeof(float), cudaMemcpyDeviceToHost, st), "D2H");
    CudaOk(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    std::lock_guard<std::mutex> lock(mutex_);
    for (int i = 0; i < num_frames; ++i)
      feature_queue_.emplace_back(h_feat_ + (size_t)i * feature_dim_, h_feat_ + (size_t)(i + 1) * feature_dim_);
    num_frames_ += num_frames;
  }
  // keep what the next frame still needs (feature_pipeline.cc:41-44)
  const size_t consumed = (size_t)config_.frame_shift * num_frames;
  remained_wav_.assign(waves.begin() + consumed, waves.end());
  finish_condition_.notify_one();
}

void FeaturePipeline::AcceptWaveform(const std::vector<int16_t>& wav) {
  std::vector<float> float_wav(wav.size());
  for (size_t i = 0; i < wav.size(); i++) float_wav[i] = static_cast<float>(wav[i]);
  this->AcceptWaveform(float_wav);
}

void FeaturePipeline::set_input_finished() {
  {
    std::lock_guard<std::mutex> lock(mutex_);
    if (input_finished_) Fatal("set_input_finished", "called twice");
    input_finished_ = true;
  }
  finish_condition_.notify_one();
}

bool FeaturePipeline::ReadOne(std::vector<float>* feat) {
  std::unique_lock<std::mutex> lock(mutex_);
  finish_condition_.wait(lock, [this] { return !feature_queue_.empty() || input_finished_; });
  if (feature_queue_.empty()) return false;                 // input finished and nothing left
  *feat = std::move(feature_queue_.front());
  feature_queue_.pop_front();
  return true;
}

bool FeaturePipeline::Read(int num_frames, std::vector<std::vector<float>>* feats) {
  feats->clear();
  std::vector<float> feat;
  while ((int)feats->size() < num_frames) {
    if (!ReadOne(&feat)) return false;
    feats->push_back(std::move(feat));
  }
  return true;
}

void FeaturePipeline::Reset() {
  std::lock_guard<std::mutex> lock(mutex_);
  input_finished_ = false;
  num_frames_ = 0;
  remained_wav_.clear();
  feature_queue_.clear();
}

int FeaturePipeline::NumQueuedFrames() const {
  std::lock_guard<std::mutex> lock(mutex_);
  return (int)feature_queue_.size();
}

}  // namespace wenet
