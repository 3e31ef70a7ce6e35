// GPU twin of the reference runtime's streaming front-end, runtime/core/frontend/feature_pipeline.h:28-117:
// same namespace, class names and public surface (FeaturePipelineConfig, FeaturePipeline::AcceptWaveform /
// set_input_finished / ReadOne / Read / Reset / num_frames / feature_dim / IsLastFrame / NumQueuedFrames), same
// two-thread contract (producer AcceptWaveform, blocking consumer Read).  The feature extraction done inside
// AcceptWaveform -- wenet::Fbank::Compute, frontend/fbank.h:138-198: Hamming window, DC removal, pre-emphasis
// 0.97, 512-point FFT power spectrum, Kaldi mel banks from 20 Hz, log -- runs as wekws_fbank_forward on the GPU.
#ifndef WEKWS_B200_RUNTIME_FEATURE_PIPELINE_B200_H_
#define WEKWS_B200_RUNTIME_FEATURE_PIPELINE_B200_H_

#include <stdint.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

struct wekws_fbank;

namespace wenet {

struct FeaturePipelineConfig {
  int num_bins;
  int sample_rate;
  int frame_length;
  int frame_shift;
  FeaturePipelineConfig(int num_bins, int sample_rate) : num_bins(num_bins), sample_rate(sample_rate) {
    frame_length = sample_rate / 1000 * 25;  // 25 ms
    frame_shift = sample_rate / 1000 * 10;   // 10 ms
  }
  void Info() const;
};

class FeaturePipeline {
 public:
  explicit FeaturePipeline(const FeaturePipelineConfig& config);
  ~FeaturePipeline();
  FeaturePipeline(const FeaturePipeline&) = delete;
  FeaturePipeline& operator=(const FeaturePipeline&) = delete;

  // The feature extraction is done in AcceptWaveform(); samples in int16 range.
  void AcceptWaveform(const std::vector<float>& wav);
  void AcceptWaveform(const std::vector<int16_t>& wav);

  int num_frames() const { return num_frames_; }
  int feature_dim() const { return feature_dim_; }
  const FeaturePipelineConfig& config() const { return config_; }

  // Call when the speech input ends; never call AcceptWaveform() afterwards.
  void set_input_finished();
  bool input_finished() const { return input_finished_; }

  // Blocking: false once the input is finished and the queue is empty.
  bool ReadOne(std::vector<float>* feat);
  // Blocking: false if fewer than num_frames could be read before the end of input.
  bool Read(int num_frames, std::vector<std::vector<float>>* feats);

  void Reset();
  bool IsLastFrame(int frame) const { return input_finished_ && (frame == num_frames_ - 1); }
  int NumQueuedFrames() const;

 private:
  const FeaturePipelineConfig config_;
  int feature_dim_;
  wekws_fbank* fbank_ = nullptr;
  void* stream_ = nullptr;
  float* d_wav_ = nullptr;
  float* d_feat_ = nullptr;
  float* h_feat_ = nullptr;      // pinned
  size_t wav_capacity_ = 0, feat_capacity_ = 0;

  std::deque<std::vector<float>> feature_queue_;
  int num_frames_ = 0;
  bool input_finished_ = false;
  std::vector<float> remained_wav_;   // samples after the last complete frame shift, kept for the next call
  mutable std::mutex mutex_;
  std::condition_variable finish_condition_;
};

}  // namespace wenet

#endif  // WEKWS_B200_RUNTIME_FEATURE_PIPELINE_B200_H_
