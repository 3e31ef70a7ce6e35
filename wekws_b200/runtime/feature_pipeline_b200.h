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

// 25 ms windows every 10 ms, as the reference derives them from the sample rate.
struct FeaturePipelineConfig {
  int num_bins, sample_rate, frame_length, frame_shift;
  FeaturePipelineConfig(int bins, int rate)
      : num_bins(bins), sample_rate(rate), frame_length(rate / 1000 * 25), frame_shift(rate / 1000 * 10) {}
  void Info() const;
};

class FeaturePipeline {
 public:
  explicit FeaturePipeline(const FeaturePipelineConfig& config);
  ~FeaturePipeline();
  FeaturePipeline(const FeaturePipeline&) = delete;
  FeaturePipeline& operator=(const FeaturePipeline&) = delete;

  // Producer side.  Samples are in int16 range; every complete frame of (carry + wav) is turned into a feature
  // row on the GPU before the call returns, the tail that the next frame still needs is carried over.
  void AcceptWaveform(const std::vector<float>& wav);
  void AcceptWaveform(const std::vector<int16_t>& wav);
  void set_input_finished();                       // end of audio: wakes a blocked reader; no AcceptWaveform after it

  // Consumer side (both block while the queue is empty and the input is still open).
  bool ReadOne(std::vector<float>* feat);          // false <=> input finished and queue drained
  bool Read(int num_frames, std::vector<std::vector<float>>* feats);   // false <=> stream ended before num_frames rows

  void Reset();                                    // new utterance: drops queue, carry and counters
  int num_frames() const { return num_frames_; }   // rows produced so far
  int feature_dim() const { return feature_dim_; }
  int NumQueuedFrames() const;
  bool input_finished() const { return input_finished_; }
  bool IsLastFrame(int frame) const { return input_finished_ && (frame == num_frames_ - 1); }
  const FeaturePipelineConfig& config() const { return config_; }

 private:
  const FeaturePipelineConfig config_;
  int feature_dim_;
  // device side: front-end handle, stream, staging buffers (grown on demand)
  wekws_fbank* fbank_ = nullptr;
  void* stream_ = nullptr;
  float* d_wav_ = nullptr;
  float* d_feat_ = nullptr;
  float* h_feat_ = nullptr;                        // pinned
  size_t wav_capacity_ = 0, feat_capacity_ = 0;
  // host side: produced rows, carry-over samples, end-of-input flag
  std::deque<std::vector<float>> feature_queue_;
  std::vector<float> remained_wav_;
  int num_frames_ = 0;
  bool input_finished_ = false;
  mutable std::mutex mutex_;
  std::condition_variable finish_condition_;
};

}  // namespace wenet

#endif  // WEKWS_B200_RUNTIME_FEATURE_PIPELINE_B200_H_
