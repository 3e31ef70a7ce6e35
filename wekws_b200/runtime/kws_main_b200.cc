// Feature-level twin of runtime/core/bin/kws_main.cc:43-61: streams a feature matrix through
// wekws::KeywordSpotting batch by batch and prints "frame N prob p..." lines (same format).
//   kws_main_b200 model.wkb feats.f32 feature_dim batch_size [reset_every_n_batches]
// feats.f32: raw little-endian float32, num_frames x feature_dim.
//   kws_main_b200 --pcm model.wkb audio.s16 fbank_dim batch_size [chunk_samples]
// the full kws_main.cc flow: raw 16 kHz int16 PCM -> wenet::FeaturePipeline (GPU Fbank) fed chunk by chunk from a
// producer thread (stream_kws_main.cc:36-43 pattern) -> wekws::KeywordSpotting batch by batch.
#include <stdio.h>
#include <stdlib.h>

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <stdint.h>
#include <string.h>

#include <thread>

#include "feature_pipeline_b200.h"
#include "keyword_spotting_b200.h"

static int RunPcm(int argc, char* argv[]) {
  if (argc < 6) {
    fprintf(stderr, "Usage: kws_main_b200 --pcm kws_model.wkb audio.s16 fbank_dim(int) batch_size(int) [chunk_samples]\n");
    return 2;
  }
  const std::string model_path = argv[2], pcm_path = argv[3];
  const int num_bins = std::stoi(argv[4]), batch_size = std::stoi(argv[5]);
  const size_t chunk = argc > 6 ? std::stoul(argv[6]) : 4800;       // 0.3 s
  std::ifstream is(pcm_path, std::ios::binary | std::ios::ate);
  if (!is) { fprintf(stderr, "cannot open %s\n", pcm_path.c_str()); return 2; }
  const std::streamsize bytes = is.tellg();
  is.seekg(0);
  std::vector<int16_t> pcm(bytes / sizeof(int16_t));
  is.read(reinterpret_cast<char*>(pcm.data()), bytes);

  wenet::FeaturePipelineConfig feature_config(num_bins, 16000);
  wenet::FeaturePipeline feature_pipeline(feature_config);
  wekws::KeywordSpotting spotter(model_path);
  std::thread producer([&] {
    for (size_t s = 0; s < pcm.size(); s += chunk) {
      const size_t e = std::min(pcm.size(), s + chunk);
      feature_pipeline.AcceptWaveform(std::vector<int16_t>(pcm.begin() + s, pcm.begin() + e));
    }
    feature_pipeline.set_input_finished();
  });
  std::cout.precision(9);
  int offset = 0;
  while (true) {                                                     // kws_main.cc:45-61
    std::vector<std::vector<float>> feats, prob;
    const bool ok = feature_pipeline.Read(batch_size, &feats);
    spotter.Forward(feats, &prob);
    for (size_t i = 0; i < prob.size(); i++) {
      std::cout << "frame " << offset + i << " prob";
      for (size_t j = 0; j < prob[i].size(); j++) std::cout << " " << prob[i][j];
      std::cout << std::endl;
    }
    if (!ok) break;
    offset += prob.size();
  }
  producer.join();
  return 0;
}

int main(int argc, char* argv[]) {
  if (argc > 1 && strcmp(argv[1], "--pcm") == 0) return RunPcm(argc, argv);
  if (argc < 5) {
    fprintf(stderr, "Usage: kws_main_b200 kws_model.wkb feats.f32 feature_dim(int) batch_size(int) [reset_every]\n");
    return 2;
  }
  const std::string model_path = argv[1], feats_path = argv[2];
  const int dim = std::stoi(argv[3]), batch_size = std::stoi(argv[4]);
  const int reset_every = argc > 5 ? std::stoi(argv[5]) : 0;
  std::ifstream is(feats_path, std::ios::binary | std::ios::ate);
  if (!is) { fprintf(stderr, "cannot open %s\n", feats_path.c_str()); return 2; }
  const std::streamsize bytes = is.tellg();
  is.seekg(0);
  std::vector<float> flat(bytes / sizeof(float));
  is.read(reinterpret_cast<char*>(flat.data()), bytes);
  const int num_frames = static_cast<int>(flat.size() / dim);

  wekws::KeywordSpotting::InitEngineThreads(1);
  wekws::KeywordSpotting spotter(model_path);
  std::cout.precision(9);
  int offset = 0, batches = 0;
  while (offset < num_frames) {
    const int n = std::min(batch_size, num_frames - offset);
    std::vector<std::vector<float>> feats(n), prob;
    for (int i = 0; i < n; ++i) feats[i].assign(flat.begin() + (size_t)(offset + i) * dim, flat.begin() + (size_t)(offset + i + 1) * dim);
    spotter.Forward(feats, &prob);
    for (size_t i = 0; i < prob.size(); i++) {
      std::cout << "frame " << offset + i << " prob";
      for (size_t j = 0; j < prob[i].size(); j++) std::cout << " " << prob[i][j];
      std::cout << std::endl;
    }
    offset += n;
    if (reset_every > 0 && ++batches % reset_every == 0) spotter.Reset();
  }
  return 0;
}
