/* TEST INFRASTRUCTURE ONLY — CPU restatement of the per-sample send / wet chain around the convolver in
 * REEVRAudioProcessor::processBlock (SURVEY 8f-1 remainder and 8f-4), the oracle of b200conv_chain_process:
 *
 *   send   lin = dry * ysend ; low cut (HP) if > 20 Hz ; high cut (LP) if < 20 kHz   src/PluginProcessor.cpp:1639-1653
 *          Filter::init / eval / reset                                                 src/dsp/Filter.cpp:3-75
 *          Filter::getCoeff + LookupTable::cubic (tan LUT, 2048 points)                src/dsp/Filter.h:28-44, src/dsp/Utils.h:44-112
 *   delay  predelay ring                                                               src/PluginProcessor.cpp:1766-1790
 *   wet    L = LL (+ RL), R = RR (+ LR) ; * yrev ; mid/side width ; normalisation      src/PluginProcessor.cpp:1832-1856
 *   mix    out = drygain * dry + wetgain * wet                                         src/PluginProcessor.cpp:1859-1876
 *
 * float32 arithmetic in the reference's evaluation order.  The filter part is pinned against the reference's own
 * Filter.cpp compiled into oracle/_ref/libreffilter.so (tests/test_chain.py); the convolver in the middle is the
 * oracle of partconv_oracle.c.  Nothing in the product may link this file.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* ---- LookupTable (src/dsp/Utils.h:50-112) with the tan() table of Filter::coeffLUT (src/dsp/Filter.h:28-36) ---- */
#define OC_LUT_N 2048
static float g_lut[OC_LUT_N];
static int g_lut_ready = 0;
static const float kLutMin = 0.0f, kLutMax = 0.5f;

static void lut_init(void) {
  if (g_lut_ready) return;
  const float pi = 3.14159265358979323846f;
  for (size_t i = 0; i < OC_LUT_N; ++i) {
    float x = (float)i / (float)(OC_LUT_N - 1);                 /* Utils.h:69 */
    float mapped = kLutMin + x * (kLutMax - kLutMin);           /* :70 */
    if (mapped < kLutMin) mapped = kLutMin;
    if (mapped > kLutMax) mapped = kLutMax;
    const float kMaxRads = 0.499f * pi;                         /* Filter.h:30 */
    float scaled = mapped * pi;                                 /* :31 */
    g_lut[i] = tanf(scaled < kMaxRads ? scaled : kMaxRads);     /* :32 (std::tan(float)) */
  }
  g_lut_ready = 1;
}

static float lut_cubic(float input) {                           /* Utils.h:89-112 */
  lut_init();
  const float scaler = (float)(OC_LUT_N - 1) / (kLutMax - kLutMin);   /* :65 */
  const float offset = -kLutMin * scaler;                       /* :66 */
  if (input < kLutMin) input = kLutMin;
  if (input > kLutMax) input = kLutMax;
  float index = input * scaler + offset;
  int i = (int)index;
  float t = index - (float)i;
  int i0 = i - 1 > 0 ? i - 1 : 0;
  int i1 = i;
  int i2 = i + 1 < OC_LUT_N - 1 ? i + 1 : OC_LUT_N - 1;
  int i3 = i + 2 < OC_LUT_N - 1 ? i + 2 : OC_LUT_N - 1;
  float y0 = g_lut[i0], y1 = g_lut[i1], y2 = g_lut[i2], y3 = g_lut[i3];
  float a0 = y3 - y2 - y0 + y1;
  float a1 = y0 - y1 - a0;
  float a2 = y2 - y0;
  float a3 = y1;
  return (a0 * t * t * t) + (a1 * t * t) + (a2 * t) + a3;
}

float oc_filter_coeff(float freq, float srate) {                /* Filter::getCoeff, Filter.h:40-44 */
  const float lo = 20.0f, hi = srate * 0.48f;
  if (freq < lo) freq = lo;
  if (freq > hi) freq = hi;
  float ratio = freq / srate;
  if (ratio < 0.0f) ratio = 0.0f;
  if (ratio > 0.5f) ratio = 0.5f;
  return lut_cubic(ratio);
}

/* ---- Filter (src/dsp/Filter.cpp) ---- */
typedef struct oc_filter {
  int slope, mode;                       /* slope 0/1/2 = 6/12/24 dB, mode 0/1/2 = LP/BP/HP (Filter.h:7-18) */
  float g, k, k2, a1, a2, a3, a12, a22, a32;
  float ic1, ic2, ic3, ic4, state;
} oc_filter;

void oc_filter_init(oc_filter* f, int slope, int mode, float srate, float freq, float q) {   /* Filter.cpp:3-21 */
  const float q2 = 0.6173f;
  memset(f, 0, sizeof(*f));
  f->slope = slope; f->mode = mode;
  f->g = oc_filter_coeff(freq, srate);
  f->k = 2 - 2 * q;
  f->k2 = 2 - 2 * q2;
  if (slope == 0) {
    f->g = f->g / (1.0f + f->g);
  } else {
    f->a1 = 1.0f / (1.0f + f->g * (f->g + f->k));
    f->a2 = f->g * f->a1;
    f->a3 = f->g * f->a2;
    f->a12 = 1.0f / (1.0f + f->g * (f->g + f->k2));
    f->a22 = f->g * f->a12;
    f->a32 = f->g * f->a22;
  }
}

void oc_filter_reset(oc_filter* f, float v) { f->ic1 = f->ic2 = f->ic3 = f->ic4 = v; f->state = v; }   /* :71-75 */

float oc_filter_eval(oc_filter* f, float sample) {              /* Filter.cpp:23-68 */
  if (f->slope == 0) {
    float delta = f->g * (sample - f->state);
    f->state += delta;
    float low = f->state;
    return f->mode == 0 ? f->state : sample - low;
  }
  float v3 = sample - f->ic2;
  float v1 = f->a1 * f->ic1 + f->a2 * v3;
  float v2 = f->ic2 + f->a2 * f->ic1 + f->a3 * v3;
  f->ic1 = 2.0f * v1 - f->ic1;
  f->ic2 = 2.0f * v2 - f->ic2;
  float output;
  if (f->mode == 0) output = v2;
  else if (f->mode == 1) output = v1;
  else output = sample - f->k * v1 - v2;
  if (f->slope == 1) return output;
  v3 = output - f->ic4;
  v1 = f->a12 * f->ic3 + f->a22 * v3;
  v2 = f->ic4 + f->a22 * f->ic3 + f->a32 * v3;
  f->ic3 = 2.0f * v1 - f->ic3;
  f->ic4 = 2.0f * v2 - f->ic4;
  if (f->mode == 0) output = v2;
  else if (f->mode == 1) output = v1;
  else output = output - f->k2 * v1 - v2;
  return output;
}

void* oc_filter_create(int slope, int mode, float srate, float freq, float q) {
  oc_filter* f = (oc_filter*)malloc(sizeof(oc_filter));
  oc_filter_init(f, slope, mode, srate, freq, q);
  return f;
}
void oc_filter_destroy(void* f) { free(f); }
void oc_filter_run(void* f, const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = oc_filter_eval((oc_filter*)f, in[i]);
}

/* ---- the chain state ---- */
typedef struct oc_chain {
  float srate;
  int lowcut_on, highcut_on;
  oc_filter lc[2], hc[2];
  int predelay, delay_size, delaypos;
  float* delay[2];
  float width, drygain, wetgain;
} oc_chain;

/* q as PluginProcessor.cpp:845-848 chooses it: 0.0765 for 24 dB, 0.2929 otherwise */
static float q_for(int slope) { return slope == 2 ? 0.0765f : 0.2929f; }

void* oc_chain_create(float srate, float lowcut_hz, int lowcut_slope, float highcut_hz, int highcut_slope,
                      int predelay, int delay_size, float width, float drygain, float wetgain) {
  oc_chain* c = (oc_chain*)calloc(1, sizeof(oc_chain));
  c->srate = srate;
  c->lowcut_on = lowcut_hz > 20.0f;                              /* PluginProcessor.cpp:1643 */
  c->highcut_on = highcut_hz < 20000.0f;                         /* :1647 */
  for (int ch = 0; ch < 2; ++ch) {
    oc_filter_init(&c->lc[ch], lowcut_slope, 2, srate, lowcut_hz, q_for(lowcut_slope));     /* HP, PluginProcessor.h:250 */
    oc_filter_init(&c->hc[ch], highcut_slope, 0, srate, highcut_hz, q_for(highcut_slope));  /* LP, :248 */
    oc_filter_reset(&c->lc[ch], 0.0f);
    oc_filter_reset(&c->hc[ch], 0.0f);
  }
  c->predelay = predelay;
  c->delay_size = delay_size > predelay ? delay_size : predelay + 1;
  for (int ch = 0; ch < 2; ++ch) c->delay[ch] = (float*)calloc((size_t)c->delay_size, sizeof(float));
  c->width = width; c->drygain = drygain; c->wetgain = wetgain;
  return c;
}
void oc_chain_destroy(void* p) {
  oc_chain* c = (oc_chain*)p;
  free(c->delay[0]); free(c->delay[1]); free(c);
}

/* send side of one block: dry L/R (n samples) -> convolver input L/R  (PluginProcessor.cpp:1639-1653, 1766-1790) */
void oc_chain_send(void* p, const float* dryL, const float* dryR, const float* ysend, float* convL, float* convR, size_t n) {
  oc_chain* c = (oc_chain*)p;
  const float* dry[2] = {dryL, dryR};
  float* conv[2] = {convL, convR};
  for (int ch = 0; ch < 2; ++ch) {
    for (size_t i = 0; i < n; ++i) {
      float v = dry[ch][i] * ysend[i];
      if (c->lowcut_on) v = oc_filter_eval(&c->lc[ch], v);
      if (c->highcut_on) v = oc_filter_eval(&c->hc[ch], v);
      c->delay[ch][(c->delaypos + (int)i) % c->delay_size] = v;                      /* :1773-1776 */
    }
    const int readpos = (c->delaypos + c->delay_size - c->predelay) % c->delay_size; /* :1783 */
    for (size_t i = 0; i < n; ++i) conv[ch][i] = c->delay[ch][(readpos + (int)i) % c->delay_size];
  }
  c->delaypos = (c->delaypos + (int)n) % c->delay_size;                               /* :1790 */
}

/* wet side: convolver outputs LL, RR (+ RL, LR when quad & true stereo) -> plugin output (:1832-1876) */
void oc_chain_wet(void* p, const float* dryL, const float* dryR, const float* LL, const float* RR, const float* LR,
                  const float* RL, const float* yrev, float* outL, float* outR, size_t n) {
  oc_chain* c = (oc_chain*)p;
  const float normalization = 1.0f / (1.0f + c->width);
  for (size_t i = 0; i < n; ++i) {
    float wl = LL[i], wr = RR[i];
    if (RL) wl += RL[i];
    if (LR) wr += LR[i];
    float lin = wl * yrev[i], rin = wr * yrev[i];
    float mid = (lin + rin) * 0.5f, side = (lin - rin) * 0.5f;
    float lout = (mid + side * c->width) * normalization;
    float rout = (mid - side * c->width) * normalization;
    outL[i] = dryL[i] * c->drygain + lout * c->wetgain;
    outR[i] = dryR[i] * c->drygain + rout * c->wetgain;
  }
}

/* ---- IR shaping pipeline: the device-resident subset of Impulse::recalcImpulse
 *      (src/dsp/Impulse.cpp:297-360) in the reference's order:
 *        calculateAutoGain :703-720 -> scale :313-320 ; reverse :322-330 ; applyTrim :437-470 ; applyGain :472-486 ;
 *        applyDecayEQ -> applyDecay :602-648 (oc_apply_decay, pinned by the reference's AudioFFT) ; applyClip :488-501 ;
 *        applyEnvelope :651-680.
 *      Not covered (stay on the host, JUCE interpolators / other filter class): resampleIRToProjectRate, applyStretch,
 *      applyParamEQ.  ch[c]: C channels of n taps, shaped in place; returns the new length (trim shortens). ---- */
void oc_apply_decay(float* buf, size_t n, const double* lut, double srate);

size_t oc_ir_shape(float** ch, int C, size_t n, int autogain, int reverse, float trim_left, float trim_right, float gain,
                   const double* lut, double srate, int clip, float attack, float decay) {
  if (n == 0 || C < 2) return n;
  if (autogain) {                                               /* calculateAutoGain(bufferLL, bufferRR) */
    double energy = 0.0;
    for (size_t i = 0; i < n; ++i) {
      double l = (double)ch[0][i], r = (double)ch[1][i];
      energy += l * l + r * r;
    }
    float ag = 1.0f;
    if (energy > 0.0) {
      double a = 1.0 / sqrt(energy);
      if (a > 1.0) a = 1.0;
      ag = (float)a;
    }
    for (int c = 0; c < C; ++c) for (size_t i = 0; i < n; ++i) ch[c][i] *= ag;
  }
  if (reverse)
    for (int c = 0; c < C; ++c)
      for (size_t i = 0; i < n / 2; ++i) { float t = ch[c][i]; ch[c][i] = ch[c][n - 1 - i]; ch[c][n - 1 - i] = t; }
  {                                                             /* applyTrim */
    size_t start = (size_t)(trim_left * (float)n);
    size_t end = n - (size_t)(trim_right * (float)n);
    if (start >= end || start >= n || end > n) return 0;
    if (start > 0) for (int c = 0; c < C; ++c) memmove(ch[c], ch[c] + start, (end - start) * sizeof(float));
    n = end - start;
  }
  for (int c = 0; c < C; ++c) for (size_t i = 0; i < n; ++i) ch[c][i] *= gain;       /* applyGain */
  if (lut) for (int c = 0; c < C; ++c) oc_apply_decay(ch[c], n, lut, srate);        /* applyDecayEQ */
  if (clip)
    for (int c = 0; c < C; ++c)
      for (size_t i = 0; i < n; ++i) { float v = ch[c][i]; ch[c][i] = v < -1.f ? -1.f : (v > 1.f ? 1.f : v); }
  {                                                             /* applyEnvelope */
    int size = (int)n;
    int attackSize = (int)(attack * (float)size);
    int decaySize = (int)(decay * (float)size);
    for (int i = 0; i < attackSize; ++i) {
      float g = (float)i / (float)attackSize;
      for (int c = 0; c < C; ++c) ch[c][i] *= g;
    }
    for (int i = 0; i < decaySize; ++i) {
      float t = (float)i / (float)decaySize;
      float g = 1.0f - (float)pow((double)t, 0.5);
      int idx = size - decaySize + i;
      for (int c = 0; c < C; ++c) ch[c][idx] *= g;
    }
  }
  return n;
}
