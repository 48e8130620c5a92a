// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// extern "C" handle wrapper around the UNMODIFIED reference classes
//   fftconvolver::FFTConvolver          (libs/FFTConvolver/FFTConvolver.h:62-80)
//   fftconvolver::TwoStageFFTConvolver  (libs/FFTConvolver/TwoStageFFTConvolver.h:65-83)
// compiled from the sources where they lie under /root/reference (see oracle/Makefile).
// No reference source is copied into this repository; this file only forwards calls.
#include <cstddef>
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

}  // extern "C"
