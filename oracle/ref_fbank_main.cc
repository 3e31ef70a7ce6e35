// Driver around the REFERENCE's own C++ front-end (runtime/core/frontend/fbank.h + fft.cc), compiled where
// the sources lie under /root/reference by oracle/Makefile into oracle/_ref/ref_fbank.  Test infrastructure
// only: it pins the Hamming-window mode of the GPU Fbank kernel and of oracle/kws_oracle.py::fbank to what
// wenet::Fbank::Compute (fbank.h:138-198) really produces.  No reference source is copied into this repo.
//
//   ref_fbank <num_bins> <in.f32> <out.f32>     raw little-endian float32 samples in, frames x bins out
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "frontend/fbank.h"

int main(int argc, char** argv) {
  if (argc != 4) {
    fprintf(stderr, "usage: %s num_bins in.f32 out.f32\n", argv[0]);
    return 2;
  }
  const int num_bins = atoi(argv[1]);
  FILE* fi = fopen(argv[2], "rb");
  if (!fi) return 3;
  fseek(fi, 0, SEEK_END);
  const long n = ftell(fi) / 4;
  fseek(fi, 0, SEEK_SET);
  std::vector<float> wave(n);
  if (fread(wave.data(), 4, n, fi) != (size_t)n) return 4;
  fclose(fi);
  wenet::Fbank fbank(num_bins, 16000, 400, 160);   // same arguments as FeaturePipelineConfig (feature_pipeline.h)
  std::vector<std::vector<float>> feats;
  const int frames = fbank.Compute(wave, &feats);
  FILE* fo = fopen(argv[3], "wb");
  if (!fo) return 5;
  for (int i = 0; i < frames; ++i) fwrite(feats[i].data(), 4, num_bins, fo);
  fclose(fo);
  return 0;
}
