// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// extern "C" handle wrapper around the UNMODIFIED reference classes
//   fftconvolver::FFTConvolver          (libs/FFTConvolver/FFTConvolver.h:62-80)
//   fftconvolver::TwoStageFFTConvolver  (libs/FFTConvolver/TwoStageFFTConvolver.h:65-83)
// compiled from the sources where they lie under /root/reference (see oracle/Makefile).
// No reference source is copied into this repository; this file only forwards calls.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>
#include "AudioFFT.h"
#include "FFTConvolver.h"
#include "TwoStageFFTConvolver.h"

using fftconvolver::FFTConvolver;
using fftconvolver::TwoStageFFTConvolver;

extern "C" {

void* ref_uniform_create() { return new FFTConvolver(); }
void ref_uniform_destroy(void* h) { delete static_cast<FFTConvolver*>(h); }
int ref_uniform_init(void* h, size_t block, const float* ir, size_t len) {
  return static_cast<FFTConvolver*>(h)->init(block, ir, len) ? 1 : 0;
}
void ref_uniform_process(void* h, const float* in, float* out, size_t len) {
  static_cast<FFTConvolver*>(h)->process(in, out, len);
}
void ref_uniform_clear(void* h) { static_cast<FFTConvolver*>(h)->clear(); }
void ref_uniform_reset(void* h) { static_cast<FFTConvolver*>(h)->reset(); }

void* ref_twostage_create() { return new TwoStageFFTConvolver(); }
void ref_twostage_destroy(void* h) { delete static_cast<TwoStageFFTConvolver*>(h); }
int ref_twostage_init(void* h, size_t head, size_t tail, const float* ir, size_t len) {
  return static_cast<TwoStageFFTConvolver*>(h)->init(head, tail, ir, len) ? 1 : 0;
}
void ref_twostage_process(void* h, const float* in, float* out, size_t len) {
  static_cast<TwoStageFFTConvolver*>(h)->process(in, out, len);
}
void ref_twostage_clear(void* h) { static_cast<TwoStageFFTConvolver*>(h)->clear(); }
void ref_twostage_reset(void* h) { static_cast<TwoStageFFTConvolver*>(h)->reset(); }

// Drives `calls` successive process() calls of `chunk` samples each over contiguous
// in/out arrays (used by the CPU-baseline timers so the Python loop is not in the timed path).
void ref_uniform_run(void* h, const float* in, float* out, size_t chunk, size_t calls) {
  FFTConvolver* c = static_cast<FFTConvolver*>(h);
  for (size_t i = 0; i < calls; ++i) c->process(in + i * chunk, out + i * chunk, chunk);
}
void ref_twostage_run(void* h, const float* in, float* out, size_t chunk, size_t calls) {
  TwoStageFFTConvolver* c = static_cast<TwoStageFFTConvolver*>(h);
  for (size_t i = 0; i < calls; ++i) c->process(in + i * chunk, out + i * chunk, chunk);
}

// The STFT loop of Impulse::applyDecay (src/dsp/Impulse.cpp:602-648) driven through the reference's OWN compiled
// audiofft::AudioFFT (libs/FFTConvolver/AudioFFT.cpp, the FFT Impulse uses — src/dsp/Impulse.h:8,33).  Impulse.cpp
// itself needs the JUCE audio-format classes and cannot be built here, so the loop around the two FFT calls is
// restated (same statements, same order); the transform arithmetic — where the precision lives — is the reference's.
// Pins oc_apply_decay of partconv_oracle.c.  window: the 4096-point window of Impulse.cpp:65-69.
void ref_stft_decay(float* buf, size_t n, const double* lut, double srate, const float* window) {
  const size_t FFT_SIZE = 4096, HOP_SIZE = FFT_SIZE / 4;                        // Impulse.h:21-22
  const size_t numBlocks = (n + HOP_SIZE - 1) / HOP_SIZE;
  if (numBlocks < 1) return;
  audiofft::AudioFFT fft;
  fft.init(FFT_SIZE);
  std::vector<float> output(n, 0.f), norm(n, 0.f), block(FFT_SIZE, 0.0f), re(FFT_SIZE), im(FFT_SIZE);
  std::vector<double> decayACC(FFT_SIZE / 2 + 1, 1.0);
  const int skipBlocks = (int)std::ceil(100 * srate / (1000.0 * FFT_SIZE));   // EARLY_REFLECTIONS_MS, Globals.h:34
  for (size_t b = 0; b < numBlocks; ++b) {
    std::fill(block.begin(), block.end(), 0.0f);
    const size_t start = b * HOP_SIZE;
    const size_t blockSize = std::min(FFT_SIZE, n - start);
    for (size_t i = 0; i < blockSize; ++i) block[i] = buf[start + i] * window[i];
    fft.fft(block.data(), re.data(), im.data());
    if ((long long)b > (long long)skipBlocks) {
      for (size_t k = 1; k < FFT_SIZE / 2 + 1; ++k) {
        const double dec = decayACC[k] * lut[k];
        decayACC[k] = dec;
        re[k] *= (float)dec;
        im[k] *= (float)dec;
      }
    }
    fft.ifft(block.data(), re.data(), im.data());
    for (size_t i = 0; i < blockSize; ++i) {
      output[start + i] += block[i];
      norm[start + i] += window[i];
    }
  }
  for (size_t i = 0; i < n; ++i) buf[i] = norm[i] > 0.0f ? output[i] / norm[i] : 0.f;
}

}  // extern "C"
