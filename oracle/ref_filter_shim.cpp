// TEST INFRASTRUCTURE ONLY — extern "C" wrapper around the UNMODIFIED reference class Filter (src/dsp/Filter.{h,cpp}),
// compiled from /root/reference where it lies (oracle/Makefile, target `ref`) against oracle/juce_min/JuceHeader.h.
// Pins the restatement oc_filter_* of chain_oracle.c.
#include <cstddef>
#include "Filter.h"

extern "C" {
void* ref_filter_create(int slope, int mode) { return new Filter((FilterSlope)slope, (FilterMode)mode); }
void ref_filter_destroy(void* f) { delete static_cast<Filter*>(f); }
void ref_filter_init(void* f, float srate, float freq, float q) { static_cast<Filter*>(f)->init(srate, freq, q); }
void ref_filter_reset(void* f, float v) { static_cast<Filter*>(f)->reset(v); }
void ref_filter_run(void* f, const float* in, float* out, size_t n) {
  Filter* p = static_cast<Filter*>(f);
  for (size_t i = 0; i < n; ++i) out[i] = p->eval(in[i]);
}
float ref_filter_coeff(float freq, float srate) { return Filter::getCoeff(freq, srate); }
}
