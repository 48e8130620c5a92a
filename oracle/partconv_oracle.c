/*
 * partconv_oracle.c — TEST INFRASTRUCTURE ONLY (never linked, imported or executed by the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it).
 *
 * Plain-C CPU restatement of the reference's partitioned-convolution hot path
 * (tiagolr/reevr @ 989b8dce, libs/FFTConvolver).  Each function cites the reference
 * file:line it follows.  Parity status: PINNED — tests/test_oracle.py checks this file
 * against (i) the reference's own known-answer self-test (58 cases of
 * libs/FFTConvolver/test/Test.cpp:253-338, naive-convolution truth, the reference's own
 * tolerance), (ii) outputs of the unmodified reference compiled here (oracle/_ref) and
 * (iii) the committed fixtures under tests/golden/ generated from oracle/_ref.
 *
 * Arithmetic contract restated from the reference:
 *   - all interfaces float32 (Utilities.h:180 `typedef float Sample`);
 *   - forward/inverse real FFT computed in float64 and rounded to float32
 *     (AudioFFT.cpp:114-159: float->double copy, Ooura rdft in double, double->float);
 *     the transform itself is the textbook DFT (forward unscaled, e^{-2*pi*i*k*n/N};
 *     inverse scaled so ifft(fft(x)) == x), so any exact-to-double FFT reproduces it up
 *     to the final float32 rounding;
 *   - complex multiply-accumulate in float32 without FMA contraction, in the evaluation
 *     order of the SSE build (Utilities.cpp:70-91), which is what x86-64 builds use
 *     (Utilities.h:27-31).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------
 * Real FFT pair, float32 in/out, float64 inside.
 * Follows the CONTRACT of AudioFFT::fft / ::ifft (AudioFFT.cpp:1001-1010) as realised by
 * OouraFFT::fft :114-137 / ::ifft :139-159: re/im split arrays of N/2+1 bins, im[0] =
 * im[N/2] = 0, forward unscaled with the standard negative-exponent sign, inverse scaled by
 * 1/N overall.  The butterfly network is an ordinary iterative radix-2 (not Ooura's radix-4
 * code): the result is the same DFT to double precision.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  size_t n;        /* transform size (power of two, >= 1) */
  double* wr;      /* cos(2*pi*j/n), j < n/2 */
  double* wi;      /* sin(2*pi*j/n) */
  double* ar;      /* work arrays, n each */
  double* ai;
} oc_fft;

static void oc_fft_free(oc_fft* f) {
  free(f->wr); free(f->wi); free(f->ar); free(f->ai);
  memset(f, 0, sizeof(*f));
}

static void oc_fft_init(oc_fft* f, size_t n) {
  oc_fft_free(f);
  f->n = n;
  if (n == 0) return;            /* AudioFFT::init(0) is legal (AudioFFT.cpp:51-54, used by reset) */
  size_t h = n / 2 ? n / 2 : 1;
  f->wr = (double*)malloc(h * sizeof(double));
  f->wi = (double*)malloc(h * sizeof(double));
  f->ar = (double*)malloc(n * sizeof(double));
  f->ai = (double*)malloc(n * sizeof(double));
  for (size_t j = 0; j < h; ++j) {
    double a = 2.0 * M_PI * (double)j / (double)n;
    f->wr[j] = cos(a);
    f->wi[j] = sin(a);
  }
}

/* in-place complex DFT on f->ar/ai; sign = -1 forward, +1 inverse (unscaled) */
static void oc_fft_complex(oc_fft* f, int sign) {
  const size_t n = f->n;
  double* ar = f->ar; double* ai = f->ai;
  /* bit reversal */
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = ar[i]; ar[i] = ar[j]; ar[j] = t;
      t = ai[i]; ai[i] = ai[j]; ai[j] = t;
    }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t half = len >> 1, step = n / len;
    for (size_t base = 0; base < n; base += len) {
      for (size_t k = 0; k < half; ++k) {
        const double c = f->wr[k * step];
        const double s = (double)sign * f->wi[k * step];
        const size_t u = base + k, v = u + half;
        const double tr = ar[v] * c - ai[v] * s;
        const double ti = ar[v] * s + ai[v] * c;
        ar[v] = ar[u] - tr; ai[v] = ai[u] - ti;
        ar[u] += tr;        ai[u] += ti;
      }
    }
  }
}

/* AudioFFT::fft (AudioFFT.cpp:1001-1004 -> OouraFFT::fft :114-137) */
static void oc_rfft(oc_fft* f, const float* x, float* re, float* im) {
  const size_t n = f->n;
  for (size_t i = 0; i < n; ++i) { f->ar[i] = (double)x[i]; f->ai[i] = 0.0; }
  oc_fft_complex(f, -1);
  for (size_t k = 0; k <= n / 2; ++k) { re[k] = (float)f->ar[k]; im[k] = (float)f->ai[k]; }
  im[0] = 0.0f;            /* :131 */
  im[n / 2] = 0.0f;        /* :136 */
}

/* AudioFFT::ifft (AudioFFT.cpp:1007-1010 -> OouraFFT::ifft :139-159) */
static void oc_irfft(oc_fft* f, float* x, const float* re, const float* im) {
  const size_t n = f->n, h = n / 2;
  /* Hermitian extension; DC and Nyquist taken as purely real (the reference packs only
     re[0] and re[N/2] into Ooura's a[0], a[1] — :143, :155). */
  f->ar[0] = (double)re[0]; f->ai[0] = 0.0;
  if (h > 0) { f->ar[h] = (double)re[h]; f->ai[h] = 0.0; }
  for (size_t k = 1; k < h; ++k) {
    f->ar[k] = (double)re[k];     f->ai[k] = (double)im[k];
    f->ar[n - k] = (double)re[k]; f->ai[n - k] = -(double)im[k];
  }
  oc_fft_complex(f, +1);
  const double scale = 1.0 / (double)n;   /* Ooura gives n/2 * x, scaled by 2/n at :158 */
  for (size_t i = 0; i < n; ++i) x[i] = (float)(f->ar[i] * scale);
}

/* ------------------------------------------------------------------------------------------
 * Vector helpers (Utilities.cpp)
 * ---------------------------------------------------------------------------------------- */

/* ComplexMultiplyAccumulate, SSE evaluation order (Utilities.cpp:70-91):
 * first 4*(len/4) lanes: re = (re + ra*rb) - ia*ib ; im = (im + ra*ib) + ia*rb
 * scalar tail (:87-91):  re += ra*rb - ia*ib       ; im += ra*ib + ia*rb            */
static void oc_cmac(float* re, float* im, const float* ra, const float* ia,
                    const float* rb, const float* ib, size_t len) {
  const size_t end4 = 4 * (len / 4);
  for (size_t i = 0; i < end4; ++i) {
    float r = re[i] + ra[i] * rb[i];
    r = r - ia[i] * ib[i];
    re[i] = r;
    float m = im[i] + ra[i] * ib[i];
    m = m + ia[i] * rb[i];
    im[i] = m;
  }
  for (size_t i = end4; i < len; ++i) {
    re[i] += ra[i] * rb[i] - ia[i] * ib[i];
    im[i] += ra[i] * ib[i] + ia[i] * rb[i];
  }
}

/* NextPowerOf2 (Utilities.h:280-289) */
static size_t oc_next_pow2(size_t v) {
  size_t p = 1;
  while (p < v) p *= 2;
  return p;
}

/* trailing-tap trim, |h| < 1e-6 absolute (FFTConvolver.cpp:103-106, TwoStageFFTConvolver.cpp:107-110) */
static size_t oc_trim(const float* ir, size_t len) {
  while (len > 0 && fabs((double)ir[len - 1]) < 0.000001f) --len;
  return len;
}

/* ------------------------------------------------------------------------------------------
 * Uniform partitioned convolver (FFTConvolver.{h,cpp})
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  size_t block;      /* _blockSize  B  */
  size_t seg;        /* _segSize    2B */
  size_t count;      /* _segCount   P  */
  size_t bins;       /* _fftComplexSize K = B+1 */
  float* fdl_re;     /* _segments   [P][K] */
  float* fdl_im;
  float* ir_re;      /* _segmentsIR [P][K] */
  float* ir_im;
  float* fftbuf;     /* _fftBuffer  [2B] */
  float* pre_re;     /* _preMultiplied [K] */
  float* pre_im;
  float* cv_re;      /* _conv [K] */
  float* cv_im;
  float* overlap;    /* _overlap [B] */
  float* inbuf;      /* _inputBuffer [B] */
  size_t fill;       /* _inputBufferFill */
  size_t current;    /* _current */
  oc_fft fft;
} oc_uniform;

oc_uniform* oc_uniform_create(void) { return (oc_uniform*)calloc(1, sizeof(oc_uniform)); }

/* FFTConvolver::reset (FFTConvolver.cpp:56-78) */
void oc_uniform_reset(oc_uniform* c) {
  free(c->fdl_re); free(c->fdl_im); free(c->ir_re); free(c->ir_im); free(c->fftbuf);
  free(c->pre_re); free(c->pre_im); free(c->cv_re); free(c->cv_im);
  free(c->overlap); free(c->inbuf);
  oc_fft_free(&c->fft);
  memset(c, 0, sizeof(*c));
}

void oc_uniform_destroy(oc_uniform* c) { if (c) { oc_uniform_reset(c); free(c); } }

/* FFTConvolver::clear (FFTConvolver.cpp:80-90): zero overlap, input buffer and every FDL
 * segment; keeps the IR spectra.  NOTE (faithfully restated quirk): the fill
 * counter and the pre-multiplied sum are NOT reset. */
void oc_uniform_clear(oc_uniform* c) {
  if (c->count == 0) { c->current = 0; return; }
  memset(c->overlap, 0, c->block * sizeof(float));
  memset(c->inbuf, 0, c->block * sizeof(float));
  memset(c->fdl_re, 0, c->count * c->bins * sizeof(float));
  memset(c->fdl_im, 0, c->count * c->bins * sizeof(float));
  c->current = 0;
}

/* FFTConvolver::init (FFTConvolver.cpp:93-152). Returns 1 on success, 0 on failure. */
int oc_uniform_init(oc_uniform* c, size_t blockSize, const float* ir, size_t irLen) {
  oc_uniform_reset(c);
  if (blockSize == 0) return 0;                                   /* :97-100 */
  irLen = oc_trim(ir, irLen);                                      /* :103-106 */
  if (irLen == 0) return 1;                                        /* :108-111 */
  c->block = oc_next_pow2(blockSize);                              /* :113 */
  c->seg = 2 * c->block;
  c->count = (size_t)ceil((double)((float)irLen / (float)c->block));  /* :115, float division */
  c->bins = c->seg / 2 + 1;                                        /* AudioFFT::ComplexSize :1013-1016 */
  oc_fft_init(&c->fft, c->seg);
  const size_t P = c->count, K = c->bins, B = c->block;
  c->fftbuf = (float*)calloc(c->seg, sizeof(float));
  c->fdl_re = (float*)calloc(P * K, sizeof(float));
  c->fdl_im = (float*)calloc(P * K, sizeof(float));
  c->ir_re = (float*)calloc(P * K, sizeof(float));
  c->ir_im = (float*)calloc(P * K, sizeof(float));
  for (size_t p = 0; p < P; ++p) {                                 /* :129-137 */
    const size_t remaining = irLen - p * B;
    const size_t n = remaining >= B ? B : remaining;
    memcpy(c->fftbuf, ir + p * B, n * sizeof(float));              /* CopyAndPad, Utilities.h:311-317 */
    memset(c->fftbuf + n, 0, (c->seg - n) * sizeof(float));
    oc_rfft(&c->fft, c->fftbuf, c->ir_re + p * K, c->ir_im + p * K);
  }
  c->pre_re = (float*)calloc(K, sizeof(float));
  c->pre_im = (float*)calloc(K, sizeof(float));
  c->cv_re = (float*)calloc(K, sizeof(float));
  c->cv_im = (float*)calloc(K, sizeof(float));
  c->overlap = (float*)calloc(B, sizeof(float));
  c->inbuf = (float*)calloc(B, sizeof(float));
  c->fill = 0;
  c->current = 0;
  return 1;
}

/* FFTConvolver::process (FFTConvolver.cpp:155-212) */
void oc_uniform_process(oc_uniform* c, const float* input, float* output, size_t len) {
  if (c->count == 0) {                                             /* :157-161 */
    memset(output, 0, len * sizeof(float));
    return;
  }
  const size_t B = c->block, P = c->count, K = c->bins;
  size_t done = 0;
  while (done < len) {
    const int was_empty = (c->fill == 0);                          /* :166 */
    size_t n = len - done;
    if (n > B - c->fill) n = B - c->fill;                          /* :167 */
    const size_t pos = c->fill;
    memcpy(c->inbuf + pos, input + done, n * sizeof(float));       /* :169 */

    /* forward FFT of [inbuf ; 0] into FDL slot `current`            :172-173 */
    memcpy(c->fftbuf, c->inbuf, B * sizeof(float));
    memset(c->fftbuf + B, 0, B * sizeof(float));
    float* cur_re = c->fdl_re + c->current * K;
    float* cur_im = c->fdl_im + c->current * K;
    oc_rfft(&c->fft, c->fftbuf, cur_re, cur_im);

    if (was_empty) {                                               /* :176-185 */
      memset(c->pre_re, 0, K * sizeof(float));
      memset(c->pre_im, 0, K * sizeof(float));
      for (size_t i = 1; i < P; ++i) {
        const size_t a = (c->current + i) % P;
        oc_cmac(c->pre_re, c->pre_im, c->ir_re + i * K, c->ir_im + i * K,
                c->fdl_re + a * K, c->fdl_im + a * K, K);
      }
    }
    memcpy(c->cv_re, c->pre_re, K * sizeof(float));                /* :186 */
    memcpy(c->cv_im, c->pre_im, K * sizeof(float));
    oc_cmac(c->cv_re, c->cv_im, cur_re, cur_im, c->ir_re, c->ir_im, K);   /* :187 */

    oc_irfft(&c->fft, c->fftbuf, c->cv_re, c->cv_im);              /* :190 */

    for (size_t i = 0; i < n; ++i)                                 /* Sum, :193 / Utilities.cpp:34-51 */
      output[done + i] = c->fftbuf[pos + i] + c->overlap[pos + i];

    c->fill += n;
    if (c->fill == B) {                                            /* :197-208 */
      memset(c->inbuf, 0, B * sizeof(float));
      c->fill = 0;
      memcpy(c->overlap, c->fftbuf + B, B * sizeof(float));
      c->current = c->current > 0 ? c->current - 1 : P - 1;
    }
    done += n;
  }
}

/* ------------------------------------------------------------------------------------------
 * Two-stage convolver (TwoStageFFTConvolver.{h,cpp}), background processing run
 * synchronously as the base class does (TwoStageFFTConvolver.cpp:236-250).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  size_t head_block, tail_block;
  oc_uniform head, tail0, tail;
  float *out0, *pre0;          /* _tailOutput0 / _tailPrecalculated0 [T] (NULL if unused) */
  float *out1, *pre1;          /* _tailOutput  / _tailPrecalculated  [T] */
  float* tail_in;              /* _tailInput [T] */
  float* bg_in;                /* _backgroundProcessingInput [T] */
  size_t tail_fill;            /* _tailInputFill */
  size_t pre_pos;              /* _precalculatedPos */
} oc_twostage;

oc_twostage* oc_twostage_create(void) { return (oc_twostage*)calloc(1, sizeof(oc_twostage)); }

/* TwoStageFFTConvolver::reset (:51-67) */
void oc_twostage_reset(oc_twostage* c) {
  oc_uniform_reset(&c->head); oc_uniform_reset(&c->tail0); oc_uniform_reset(&c->tail);
  free(c->out0); free(c->pre0); free(c->out1); free(c->pre1); free(c->tail_in); free(c->bg_in);
  memset(c, 0, sizeof(*c));
}

void oc_twostage_destroy(oc_twostage* c) { if (c) { oc_twostage_reset(c); free(c); } }

/* TwoStageFFTConvolver::clear (:69-84) */
void oc_twostage_clear(oc_twostage* c) {
  const size_t T = c->tail_block;
  if (c->out1) memset(c->out1, 0, T * sizeof(float));
  if (c->out0) memset(c->out0, 0, T * sizeof(float));
  if (c->pre1) memset(c->pre1, 0, T * sizeof(float));
  if (c->pre0) memset(c->pre0, 0, T * sizeof(float));
  if (c->tail_in) memset(c->tail_in, 0, T * sizeof(float));
  if (c->bg_in) memset(c->bg_in, 0, T * sizeof(float));
  c->tail_fill = 0;
  c->pre_pos = 0;
  oc_uniform_clear(&c->head); oc_uniform_clear(&c->tail0); oc_uniform_clear(&c->tail);
}

/* TwoStageFFTConvolver::init (:87-148) */
int oc_twostage_init(oc_twostage* c, size_t headBlock, size_t tailBlock, const float* ir, size_t irLen) {
  oc_twostage_reset(c);
  if (headBlock == 0 || tailBlock == 0) return 0;                  /* :94-97 */
  if (headBlock > tailBlock) { size_t t = headBlock; headBlock = tailBlock; tailBlock = t; }  /* :100-104 */
  irLen = oc_trim(ir, irLen);                                      /* :107-110 */
  if (irLen == 0) return 1;                                        /* :112-115 */
  c->head_block = oc_next_pow2(headBlock);                         /* :117-118 */
  c->tail_block = oc_next_pow2(tailBlock);
  const size_t T = c->tail_block;
  oc_uniform_init(&c->head, c->head_block, ir, irLen < T ? irLen : T);        /* :120-121 */
  if (irLen > T) {                                                 /* :123-129 */
    const size_t n = (irLen - T) < T ? (irLen - T) : T;
    oc_uniform_init(&c->tail0, c->head_block, ir + T, n);
    c->out0 = (float*)calloc(T, sizeof(float));
    c->pre0 = (float*)calloc(T, sizeof(float));
  }
  if (irLen > 2 * T) {                                             /* :131-138 */
    oc_uniform_init(&c->tail, T, ir + 2 * T, irLen - 2 * T);
    c->out1 = (float*)calloc(T, sizeof(float));
    c->pre1 = (float*)calloc(T, sizeof(float));
    c->bg_in = (float*)calloc(T, sizeof(float));
  }
  if (c->pre0 || c->pre1) c->tail_in = (float*)calloc(T, sizeof(float));     /* :140-143 */
  c->tail_fill = 0;
  c->pre_pos = 0;
  return 1;
}

/* TwoStageFFTConvolver::process (:151-233) */
void oc_twostage_process(oc_twostage* c, const float* input, float* output, size_t len) {
  oc_uniform_process(&c->head, input, output, len);                /* :154 */
  if (!c->tail_in) return;                                         /* :157 */
  const size_t H = c->head_block, T = c->tail_block;
  size_t done = 0;
  while (done < len) {
    size_t n = len - done;
    const size_t room = H - (c->tail_fill % H);                    /* :163 */
    if (n > room) n = room;
    /* add what the tails computed earlier for these output positions   :166-193 */
    if (c->pre0) for (size_t i = 0; i < n; ++i) output[done + i] += c->pre0[c->pre_pos + i];
    if (c->pre1) for (size_t i = 0; i < n; ++i) output[done + i] += c->pre1[c->pre_pos + i];
    c->pre_pos += n;
    memcpy(c->tail_in + c->tail_fill, input + done, n * sizeof(float));      /* :196-197 */
    c->tail_fill += n;
    /* first tail block: head-sized partitions, runs on every completed head block   :201-210 */
    if (c->pre0 && c->tail_fill % H == 0) {
      const size_t off = c->tail_fill - H;
      oc_uniform_process(&c->tail0, c->tail_in + off, c->out0 + off, H);
      if (c->tail_fill == T) { float* t = c->pre0; c->pre0 = c->out0; c->out0 = t; }
    }
    /* remaining tail: one T-sized block per T input samples ("background")   :213-222 */
    if (c->pre1 && c->tail_fill == T) {
      float* t = c->pre1; c->pre1 = c->out1; c->out1 = t;
      memcpy(c->bg_in, c->tail_in, T * sizeof(float));
      oc_uniform_process(&c->tail, c->bg_in, c->out1, T);          /* doBackgroundProcessing :247-250 */
    }
    if (c->tail_fill == T) { c->tail_fill = 0; c->pre_pos = 0; }   /* :224-228 */
    done += n;
  }
}

/* ------------------------------------------------------------------------------------------
 * Helpers for tests / CPU-baseline timing
 * ---------------------------------------------------------------------------------------- */
void oc_uniform_run(oc_uniform* c, const float* in, float* out, size_t chunk, size_t calls) {
  for (size_t i = 0; i < calls; ++i) oc_uniform_process(c, in + i * chunk, out + i * chunk, chunk);
}
void oc_twostage_run(oc_twostage* c, const float* in, float* out, size_t chunk, size_t calls) {
  for (size_t i = 0; i < calls; ++i) oc_twostage_process(c, in + i * chunk, out + i * chunk, chunk);
}

/* Naive direct convolution in the reference self-test's accumulation type (float), the
 * known-answer truth of libs/FFTConvolver/test/Test.cpp:32-66 (same sums, same order). */
void oc_naive_convolve(const float* in, size_t inLen, const float* ir, size_t irLen, float* out) {
  if (irLen > inLen) { oc_naive_convolve(ir, irLen, in, inLen, out); return; }
  memset(out, 0, (inLen + irLen - 1) * sizeof(float));
  for (size_t n = 0; n < irLen; ++n)
    for (size_t m = 0; m <= n; ++m) out[n] += ir[m] * in[n - m];
  for (size_t n = irLen; n < inLen; ++n)
    for (size_t m = 0; m < irLen; ++m) out[n] += ir[m] * in[n - m];
  for (size_t n = inLen; n < inLen + irLen - 1; ++n)
    for (size_t m = n - inLen + 1; m < irLen; ++m) out[n] += ir[m] * in[n - m];
}

/* Introspection used by tests (post-trim partition count etc.) */
size_t oc_uniform_partitions(const oc_uniform* c) { return c->count; }
size_t oc_uniform_block(const oc_uniform* c) { return c->block; }


/* ------------------------------------------------------------------------------------------
 * SURVEY 8f-3 groundwork (NOT on the round-1 hot path): the FFT-heavy core of the IR shaping,
 * Impulse::applyDecay (src/dsp/Impulse.cpp:602-648) — a 4096-point STFT with hop 1024
 * (Impulse.h:21-22), the Blackman-type analysis window of Impulse.cpp:65-69, a per-bin decay that
 * compounds once per block after the early-reflection blocks (:612, :626-633) and an overlap-add
 * normalised by the summed window (:637-647).  Parity status: UNPINNED — Impulse.cpp needs JUCE and
 * cannot be compiled here, so this restatement is only checked through properties
 * (tests/test_oracle.py: unit LUT reproduces the input, independent float64 numpy model).
 * ---------------------------------------------------------------------------------------- */
#define OC_STFT_N 4096
#define OC_STFT_HOP (OC_STFT_N / 4)

void oc_decay_window(float* w) {                         /* Impulse.cpp:65-69 */
  const float step = 2.0f * 3.14159265358979323846f / (float)OC_STFT_N;
  for (int i = 0; i < OC_STFT_N / 2; ++i)
    w[i] = 0.42f - 0.50f * cosf((float)i * step) + 0.08f * cosf(2.0f * (float)i * step);
  for (int i = OC_STFT_N / 2; i < OC_STFT_N; ++i) w[i] = w[OC_STFT_N - 1 - i];
}

/* buf[n] in place; lut[OC_STFT_N/2 + 1] per-bin decay per block; srate as in the reference */
void oc_apply_decay(float* buf, size_t n, const double* lut, double srate) {
  if (n == 0) return;
  const size_t nblocks = (n + OC_STFT_HOP - 1) / OC_STFT_HOP;                       /* :604 */
  const int K = OC_STFT_N / 2 + 1;
  float* window = (float*)malloc(OC_STFT_N * sizeof(float));
  float* block = (float*)malloc(OC_STFT_N * sizeof(float));
  float* re = (float*)malloc(K * sizeof(float));
  float* im = (float*)malloc(K * sizeof(float));
  float* out = (float*)calloc(n, sizeof(float));
  float* norm = (float*)calloc(n, sizeof(float));
  double* acc = (double*)malloc(K * sizeof(double));
  for (int k = 0; k < K; ++k) acc[k] = 1.0;
  oc_decay_window(window);
  oc_fft f; memset(&f, 0, sizeof(f));
  oc_fft_init(&f, OC_STFT_N);
  const int skip = (int)ceil(100.0 * srate / (1000.0 * (double)OC_STFT_N));         /* :612, EARLY_REFLECTIONS_MS = 100 (Globals.h:34) */
  for (size_t b = 0; b < nblocks; ++b) {
    const size_t start = b * OC_STFT_HOP;
    size_t bs = n - start < (size_t)OC_STFT_N ? n - start : (size_t)OC_STFT_N;      /* :618 */
    memset(block, 0, OC_STFT_N * sizeof(float));
    for (size_t i = 0; i < bs; ++i) block[i] = buf[start + i] * window[i];          /* :620-621 */
    oc_rfft(&f, block, re, im);                                                     /* :623 */
    if ((long long)b > (long long)skip)                                             /* :626 */
      for (int k = 1; k < K; ++k) {
        const double d = acc[k] * lut[k];
        acc[k] = d;
        re[k] *= (float)d;
        im[k] *= (float)d;
      }
    oc_irfft(&f, block, re, im);                                                    /* :635 */
    for (size_t i = 0; i < bs; ++i) {                                               /* :637-644 */
      out[start + i] += block[i];
      norm[start + i] += window[i];
    }
  }
  for (size_t i = 0; i < n; ++i) buf[i] = norm[i] > 0.0f ? out[i] / norm[i] : 0.0f; /* :646-648 */
  oc_fft_free(&f);
  free(window); free(block); free(re); free(im); free(out); free(norm); free(acc);
}
