// Standalone check + timing of the tensor-core sweep (reevr_b200/csrc/kernels_tc.cuh) against a float64 CPU sum on
// sampled outputs.  No torch, no library: build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o tools/bin/tc_sweep_test tools/tc_sweep_test.cu
// usage: tc_sweep_test [C B P nb [grid]]      (default: a small ragged case, then the metric shape)
#include "../reevr_b200/csrc/kernels_tc.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline float frand() {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (float)((rng_state >> 40) & 0xFFFFFF) / 16777216.0f - 0.5f;
}

static int run_case(int C, int B, int P, int nb, int grid, int reps) {
  using namespace pc::tc;
  const Geom g = make_geom(P, nb);
  std::printf("case C=%d B=%d P=%d nb=%d: Q=%d nchunk=%d ntile=%d rows=%d  (A %.1f MB, Xt %.1f MB, Yt %.1f MB)\n", C, B, P, nb, g.Q, g.nchunk, g.ntile,
              g.rows, (double)C * B * g.nchunk * 2 * 16384 / 1e6, (double)C * B * 4 * g.Lt * 4 / 1e6, (double)C * B * 4 * g.Lty * 4 / 1e6);
  if (!geom_ok(g, B)) { std::printf("  geometry not supported\n"); return 1; }
  const long long xrow0 = g.Q + 3;                       // a few unused rows in front, filled with NaN
  const long long R = xrow0 + nb;
  const long long row_lo = xrow0 - (P - 1);
  std::vector<float2> H((size_t)C * P * B), X((size_t)C * R * B);
  for (auto& v : H) v = make_float2(frand(), frand());
  for (long long c = 0; c < C; ++c)
    for (long long r = 0; r < R; ++r)
      for (int k = 0; k < B; ++k)
        X[((size_t)c * R + r) * B + k] = r < row_lo ? make_float2(NAN, NAN) : make_float2(frand(), frand());
  float2 *dH, *dX, *dY; float *dA, *dXt, *dYt; int* derr;
  const size_t lines = (size_t)C * B;
  CK(cudaMalloc(&dH, H.size() * 8)); CK(cudaMalloc(&dX, X.size() * 8));
  CK(cudaMalloc(&dY, (size_t)(nb + 1) * C * B * 8));
  CK(cudaMalloc(&dA, lines * g.nchunk * 2 * 16384)); CK(cudaMalloc(&dXt, lines * 4 * g.Lt * 4)); CK(cudaMalloc(&dYt, lines * 4 * g.Lty * 4));
  CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
  CK(cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dY, 0xFF, (size_t)(nb + 1) * C * B * 8));
  CK(cudaMemset(dYt, 0xFF, lines * 4 * g.Lty * 4));
  CK(cudaFuncSetAttribute(k_tc_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaEvent_t ev[5];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  BuildAParams bp{dH, (long long)P * B, B, P, g.Q, g.nchunk, dA};
  SplitXParams sp{dX, R * B, xrow0 - g.Q, row_lo, R, B, g.rows, dXt};
  SweepParams wp{dA, dXt, dYt, (int)lines, g.ntile, g.nchunk, g.rows, g.Lty, 0, derr};
  MergeYParams mp{dYt, g.Lty, B, nb, dY, (long long)B, (long long)C * B, 1};
  float best[4] = {1e30f, 1e30f, 1e30f, 1e30f};
  for (int rep = 0; rep < reps; ++rep) {
    CK(cudaEventRecord(ev[0]));
    k_tc_build_a<<<dim3(g.nchunk, B, C), 256>>>(bp);
    CK(cudaEventRecord(ev[1]));
    k_tc_split_x<<<dim3((unsigned)(g.rows * 2), B / 32, C), dim3(32, 8)>>>(sp);
    CK(cudaEventRecord(ev[2]));
    k_tc_sweep<<<grid, kThreads, kSmemBytes>>>(wp);
    CK(cudaEventRecord(ev[3]));
    k_tc_merge_y<<<dim3((nb + 31) / 32, B / 32, C), dim3(32, 8)>>>(mp);
    CK(cudaEventRecord(ev[4]));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::printf("  kernel error: %s\n", cudaGetErrorString(e)); return 3; }
    for (int i = 0; i < 4; ++i) { float ms; CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); best[i] = std::fmin(best[i], ms); }
  }
  int err = 0;
  CK(cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost));
  const double macs = 4.0 * P * (double)nb * B * C;     // real FMAs of the direct form
  std::printf("  build_a %.3f ms | split_x %.3f ms | sweep %.3f ms (%.1f TFLOP/s direct-form equivalent) | merge_y %.3f ms | barrier status %d\n", best[0], best[1],
              best[2], 2.0 * macs / (best[2] * 1e-3) / 1e12, best[3], err);
  if (std::getenv("TC_EXPERIMENTS")) {                  // where does the time go?  (results of these runs are not valid outputs)
    for (int dbg : {1, 2, 4, 6, 7}) {
      SweepParams we = wp;
      we.dbg = dbg;
      float bestd = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(ev[0]));
        k_tc_sweep<<<grid, kThreads, kSmemBytes>>>(we);
        CK(cudaEventRecord(ev[1]));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, ev[0], ev[1])); bestd = std::fmin(bestd, ms);
      }
      std::printf("  experiment dbg=%d (%s%s%s): sweep %.3f ms\n", dbg, (dbg & 1) ? "no Yt stores " : "", (dbg & 2) ? "strips loaded once " : "",
                  (dbg & 4) ? "A ring loaded once" : "", bestd);
    }
    CK(cudaMemset(derr, 0, 4));
    k_tc_sweep<<<grid, kThreads, kSmemBytes>>>(wp);   // restore valid planes for the check below
    k_tc_merge_y<<<dim3((nb + 31) / 32, B / 32, C), dim3(32, 8)>>>(mp);
    CK(cudaDeviceSynchronize());
  }
  std::vector<float2> Y((size_t)(nb + 1) * C * B);
  CK(cudaMemcpy(Y.data(), dY, Y.size() * 8, cudaMemcpyDeviceToHost));
  double worst = 0, peak = 0;
  int bad = 0;
  const int nsamp = 4000;
  for (int sidx = 0; sidx < nsamp; ++sidx) {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    const int c = (int)((rng_state >> 33) % C);
    int k = (int)((rng_state >> 13) % B);
    long long t = (long long)((rng_state >> 20) % (unsigned long long)nb);
    if (sidx < 8) k = 0;
    if (sidx % 7 == 1) t = nb - 1 - (sidx % 5);
    if (sidx % 7 == 2) t = sidx % 130;
    double re = 0, im = 0;
    for (int p = 0; p < P; ++p) {
      const float2 h = H[((size_t)c * P + p) * B + k];
      const float2 x = X[((size_t)c * R + (xrow0 + t - p)) * B + k];
      if (k == 0) { re += (double)h.x * x.x; im += (double)h.y * x.y; }
      else { re += (double)h.x * x.x - (double)h.y * x.y; im += (double)h.x * x.y + (double)h.y * x.x; }
    }
    const float2 y = Y[((size_t)(1 + t) * C + c) * B + k];
    const double e = std::fmax(std::fabs(re - y.x), std::fabs(im - y.y));
    if (!(e == e)) ++bad;
    worst = std::fmax(worst, e);
    peak = std::fmax(peak, std::fmax(std::fabs(re), std::fabs(im)));
  }
  std::printf("  %d sampled outputs: max |err| %.3e, peak %.3f -> %.2e of peak, NaN %d  %s\n", nsamp, worst, peak, worst / peak, bad,
              (bad == 0 && err == 0 && worst < 1e-5 * peak) ? "PARITY OK" : "PARITY FAILED");
  cudaFree(dH); cudaFree(dX); cudaFree(dY); cudaFree(dA); cudaFree(dXt); cudaFree(dYt); cudaFree(derr);
  return (bad == 0 && err == 0 && worst < 1e-5 * peak) ? 0 : 1;
}

int main(int argc, char** argv) {
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  if (argc >= 5) return run_case(std::atoi(argv[1]), std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]), argc > 5 ? std::atoi(argv[5]) : nsm, 3);
  int rc = run_case(2, 64, 100, 300, nsm, 1);
  if (rc > 1) return rc;
  rc |= run_case(1, 32, 938, 20000, nsm, 2);
  if (rc > 1) return rc;
  rc |= run_case(2, 512, 938, 112608, nsm, 3);
  return rc;
}
