// irshape.cu — SURVEY 8f-3 ("next" row): the FFT-heavy core of REEV-R's IR shaping on the device.
//
// Impulse::applyDecay (src/dsp/Impulse.cpp:602-648): 4096-point STFT, hop 1024 (Impulse.h:21-22),
// analysis window of Impulse.cpp:65-69, per-bin decay that compounds once per block after the
// early-reflection blocks (:612, :626-633), inverse transform, overlap-add normalised by the summed
// window (:637-648).  On the device every STFT block is independent: the compounded decay of block b
// is lut[k]^(b - skip) in closed form, so all blocks are transformed, scaled and inverse-transformed
// in three batched launches and a fourth kernel gathers the (up to 4) overlapping blocks per sample.
// Reuses the Stockham passes / split functions of kernels.cuh (M = 2048 complex points per 4096 real).
//
// STATUS: written after the round-1 GPU budget was spent — verified on the CPU emulation against the C
// restatement (oracle/partconv_oracle.c::oc_apply_decay, itself unpinned); not yet run on a GPU.
#if defined(PC_EMULATE)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/b200conv.h"
#include "kernels.cuh"

namespace {

constexpr int kN = 4096;          // STFT size (Impulse.h:21)
constexpr int kM = kN / 2;        // complex points of the real transform
constexpr int kHop = kN / 4;      // Impulse.h:22

// analysis-windowed block straight from global memory: z[n] = (x[2n] w[2n], x[2n+1] w[2n+1])
struct StftIn {
  const float* src; const float* win; int nv;
  PC_HD int prep(int base) const { return base; }
  PC_HD float2 at(int tok, int off) const {
    const int i0 = 2 * (tok + off), i1 = i0 + 1;
    return make_float2(i0 < nv ? src[i0] * win[i0] : 0.0f, i1 < nv ? src[i1] * win[i1] : 0.0f);
  }
};
// last inverse pass: all 2M samples of the block, scaled, to a dense scratch row
struct StftOut {
  float* dst; float scale;
  PC_HD int prep(int base) const { return base; }
  PC_HD void put(int tok, int off, float2 v) const {
    const int n = tok + off;
    dst[2 * n] = v.x * scale;
    dst[2 * n + 1] = v.y * scale;
  }
};

// spectrum row of block b scaled by lut^m (m = b - skip > 0); entry 0 packs (DC, Nyquist): DC is left alone
// (the reference loop starts at k = 1, Impulse.cpp:627) and the Nyquist bin uses lut[M]
PC_HD void decay_scale(float2* row, const double* lut, int m, int k) {
  if (m <= 0) return;
  if (k == 0) { row[0].y *= (float)std::pow(lut[kM], (double)m); return; }
  const float g = (float)std::pow(lut[k], (double)m);
  row[k].x *= g;
  row[k].y *= g;
}

// out[i] = sum over the blocks covering i of scratch[b][i - b*hop], divided by the summed window (:637-648)
PC_HD float stft_gather(const float* scratch, const float* win, long long n, long long nblocks, long long i) {
  float acc = 0.0f, norm = 0.0f;
  const long long b_hi = i / kHop;
  for (long long b = b_hi; b >= 0 && b > b_hi - kN / kHop; --b) {
    if (b >= nblocks) continue;
    const long long off = i - b * kHop;
    if (off < kN) { acc += scratch[b * kN + off]; norm += win[off]; }
  }
  (void)n;
  return norm > 0.0f ? acc / norm : 0.0f;
}

#if !defined(PC_EMULATE)
__global__ void __launch_bounds__(512) k_stft_fwd(const float* x, long long n, const float* win, const float2* tw, float2* spec, long long nblocks) {
  extern __shared__ float2 sm[];
  const long long b = blockIdx.x;
  if (b >= nblocks) return;
  constexpr int NT = pc::fft_threads(kM);
  const int tx = threadIdx.x;
  float2* bufA = sm; float2* bufB = sm + kM;
  const long long start = b * kHop;
  const long long rem = n - start;
  const int nv = rem > kN ? kN : (int)rem;
  constexpr int R0 = pc::pass_radix(kM, 1);
  for (int i = tx; i < kM / R0; i += NT)
    pc::stockham_butterfly<false>(StftIn{x + start, win, nv}, pc::SmemOut{bufA}, tw + pc::tw_pass_offset(kM, 1), kM, 1, R0, i);
  __syncthreads();
  float2* res = pc::fft_mid_passes<false, kM, R0>(bufA, bufB, tw, tx, true);
  float2* row = spec + b * kM;
  for (int k = tx; k <= kM / 2; k += NT) pc::fwd_split(res, row, tw, kM, k);
}

__global__ void k_stft_decay(float2* spec, const double* lut, long long nblocks, int skip) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  if (k < kM && b < nblocks) decay_scale(spec + b * kM, lut, (int)(b - skip), k);
}

__global__ void __launch_bounds__(512) k_stft_inv(const float2* spec, const float2* zero_row, const float2* tw, float* scratch, long long nblocks) {
  extern __shared__ float2 sm[];
  const long long b = blockIdx.x;
  if (b >= nblocks) return;
  constexpr int NT = pc::fft_threads(kM);
  const int tx = threadIdx.x;
  float2* bufA = sm; float2* bufB = sm + kM;
  const float2* row = spec + b * kM;
  for (int k = tx; k <= kM / 2; k += NT) pc::inv_pre(row, zero_row, bufA, tw, kM, k, 1, 0);
  __syncthreads();
  float2* in = pc::fft_mid_passes<true, kM, 1>(bufA, bufB, tw, tx, true);
  // last pass (the one that reaches length M) writes all 2M samples of the block
  int p = 1;
  while (p * pc::pass_radix(kM, p) != kM) p *= pc::pass_radix(kM, p);
  const int R = pc::pass_radix(kM, p);
  for (int i = tx; i < kM / R; i += NT)
    pc::stockham_butterfly<true>(pc::SmemIn{in}, StftOut{scratch + b * kN, 1.0f / (float)kM}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
}

__global__ void k_stft_gather(const float* scratch, const float* win, float* out, long long n, long long nblocks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = stft_gather(scratch, win, n, nblocks, i);
}
#endif

// host reference of the per-block pipeline for the emulation build (same phase functions, loops for threads)
#if defined(PC_EMULATE)
void emu_stft(const float* x, long long n, const float* win, const float2* tw, const double* lut, int skip, float* out) {
  const long long nblocks = (n + kHop - 1) / kHop;
  std::vector<float2> spec((size_t)nblocks * kM), bufA(kM), bufB(kM), zero(kM, make_float2(0.f, 0.f));
  std::vector<float> scratch((size_t)nblocks * kN, 0.0f);
  for (long long b = 0; b < nblocks; ++b) {
    const long long start = b * kHop, rem = n - start;
    const int nv = rem > kN ? kN : (int)rem;
    const int R0 = pc::pass_radix(kM, 1);
    for (int i = 0; i < kM / R0; ++i)
      pc::stockham_butterfly<false>(StftIn{x + start, win, nv}, pc::SmemOut{bufA.data()}, tw + pc::tw_pass_offset(kM, 1), kM, 1, R0, i);
    float2* in = bufA.data(); float2* o = bufB.data();
    for (int p = R0; p < kM;) {
      const int R = pc::pass_radix(kM, p);
      for (int i = 0; i < kM / R; ++i) pc::stockham_butterfly<false>(pc::SmemIn{in}, pc::SmemOut{o}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
      std::swap(in, o);
      p *= R;
    }
    float2* row = spec.data() + b * kM;
    for (int k = 0; k <= kM / 2; ++k) pc::fwd_split(in, row, tw, kM, k);
    for (int k = 0; k < kM; ++k) decay_scale(row, lut, (int)(b - skip), k);
    for (int k = 0; k <= kM / 2; ++k) pc::inv_pre(row, zero.data(), bufA.data(), tw, kM, k, 1, 0);
    in = bufA.data(); o = bufB.data();
    for (int p = 1; p < kM;) {
      const int R = pc::pass_radix(kM, p);
      const bool last = p * R == kM;
      for (int i = 0; i < kM / R; ++i) {
        if (!last) pc::stockham_butterfly<true>(pc::SmemIn{in}, pc::SmemOut{o}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
        else pc::stockham_butterfly<true>(pc::SmemIn{in}, StftOut{scratch.data() + b * kN, 1.0f / (float)kM}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
      }
      std::swap(in, o);
      p *= R;
    }
  }
  for (long long i = 0; i < n; ++i) out[i] = stft_gather(scratch.data(), win, n, nblocks, i);
}
#endif

}  // namespace

extern "C" int b200conv_ir_decay_eq(int device, float* ir, size_t n, const double* lut, double srate) {
  if (!ir || !lut) return B200CONV_EINVAL;
  if (n == 0) return B200CONV_OK;
  // window (Impulse.cpp:65-69) and twiddles on the host, in the layout of kernels.cuh
  std::vector<float> win(kN);
  const float step = 2.0f * 3.14159265358979323846f / (float)kN;
  for (int i = 0; i < kN / 2; ++i) win[i] = 0.42f - 0.50f * std::cos((float)i * step) + 0.08f * std::cos(2.0f * (float)i * step);
  for (int i = kN / 2; i < kN; ++i) win[i] = win[kN - 1 - i];
  std::vector<float2> tw(pc::tw_table_len(kM));
  for (int k = 0; k <= kM / 2; ++k) {
    const double a = -2.0 * M_PI * (double)k / (2.0 * (double)kM);
    tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  for (int p = 1; p < kM;) {
    const int R = pc::pass_radix(kM, p);
    const int off = pc::tw_pass_offset(kM, p);
    for (int r = 1; r < R; ++r)
      for (int k = 0; k < p; ++k) {
        const double a = -2.0 * M_PI * (double)r * (double)k / ((double)p * (double)R);
        tw[off + (r - 1) * p + k] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
    p *= R;
  }
  const int skip = (int)std::ceil(100.0 * srate / (1000.0 * (double)kN));       // EARLY_REFLECTIONS_MS = 100, :612
  const long long nblocks = ((long long)n + kHop - 1) / kHop;
#if defined(PC_EMULATE)
  (void)device;
  std::vector<float> out(n);
  emu_stft(ir, (long long)n, win.data(), tw.data(), lut, skip, out.data());
  std::memcpy(ir, out.data(), n * sizeof(float));
  return B200CONV_OK;
#else
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return B200CONV_ECUDA; }
  float *dx = nullptr, *dwin = nullptr, *dscratch = nullptr, *dout = nullptr;
  float2 *dtw = nullptr, *dspec = nullptr, *dzero = nullptr;
  double* dlut = nullptr;
  cudaStream_t st = nullptr;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&dx, n * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(&dout, n * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(&dwin, kN * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(&dtw, tw.size() * sizeof(float2)) == cudaSuccess;
  ok = ok && cudaMalloc(&dlut, (kM + 1) * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&dspec, (size_t)nblocks * kM * sizeof(float2)) == cudaSuccess;
  ok = ok && cudaMalloc(&dzero, kM * sizeof(float2)) == cudaSuccess;
  ok = ok && cudaMalloc(&dscratch, (size_t)nblocks * kN * sizeof(float)) == cudaSuccess;
  if (ok) {
    const size_t smem = 2 * kM * sizeof(float2);
    cudaFuncSetAttribute(k_stft_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_stft_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemcpyAsync(dx, ir, n * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dwin, win.data(), kN * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dlut, lut, (kM + 1) * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(dzero, 0, kM * sizeof(float2), st);
    const int NT = pc::fft_threads(kM);
    k_stft_fwd<<<(unsigned)nblocks, NT, smem, st>>>(dx, (long long)n, dwin, dtw, dspec, nblocks);
    k_stft_decay<<<dim3((kM + 255) / 256, (unsigned)nblocks), 256, 0, st>>>(dspec, dlut, nblocks, skip);
    k_stft_inv<<<(unsigned)nblocks, NT, smem, st>>>(dspec, dzero, dtw, dscratch, nblocks);
    k_stft_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dscratch, dwin, dout, (long long)n, nblocks);
    cudaMemcpyAsync(ir, dout, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    ok = cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  }
  cudaFree(dx); cudaFree(dout); cudaFree(dwin); cudaFree(dtw); cudaFree(dlut); cudaFree(dspec); cudaFree(dzero); cudaFree(dscratch);
  if (st) cudaStreamDestroy(st);
  if (!ok) { cudaGetLastError(); return B200CONV_ECUDA; }
  return B200CONV_OK;
#endif
}
