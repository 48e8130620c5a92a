// tcgen05 probe for the tensor-core form of the batched sweep.
// Answers, on a B200, the three questions the design hangs on before any product code is written:
//   1. do hand-built SWIZZLE_128B K-major shared-memory descriptors + kind::tf32 give the expected product,
//   2. can the B operand be a ROW-SHIFTED window of one resident strip (start address += q * 128 B), and does the
//      descriptor's base_offset field have to carry q % 8 for that,
//   3. what a 128 x N x 8 tf32 MMA costs in cycles with both operands in shared memory (N = 128 / 256).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o tc_probe tools/tc_probe.cu ; run: ./tc_probe
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);              // start address, 16-byte units
  d |= (uint64_t)1 << 16;                                // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // stride between 8-row groups
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(unsigned long long* bar, unsigned parity, long long max_cycles) {
  const long long t0 = clock64();
  unsigned ok = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > max_cycles) return false;
  }
}

struct ProbeParams {
  const float* A;      // 128 x 32 (row-major, tf32-exact values)
  const float* X;      // strip_rows x 32
  float* D;            // 128 x N
  int N, strip_rows, q, base_off_mode;   // base_off_mode: 0 -> base_offset 0, 1 -> q % 8
  int iters;           // > 0: timing run, the 4 k-steps are issued `iters` times
  long long* cycles;
  int* status;
};

// swizzled (SWIZZLE_128B, K-major) byte offset of element (row r, float e < 32) in an image whose base is 1024-aligned
__host__ __device__ inline uint32_t sw128(uint32_t r, uint32_t e) { return r * 128u + ((((e >> 2) ^ (r & 7u)) & 7u) << 4) + (e & 3u) * 4u; }

__global__ void __launch_bounds__(128) k_probe(ProbeParams P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;                       // 128 rows x 128 B
  unsigned char* sX = base + 16384;               // strip_rows x 128 B
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_holder)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 128 * 32; i += 128) { const int r = i >> 5, e = i & 31; *reinterpret_cast<float*>(sA + sw128(r, e)) = P.A[i]; }
  for (int i = tid; i < P.strip_rows * 32; i += 128) { const int r = i >> 5, e = i & 31; *reinterpret_cast<float*>(sX + sw128(r, e)) = P.X[i]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_holder;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.N >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sX) + (uint32_t)P.q * 128u;
    const uint32_t bo = P.base_off_mode ? (uint32_t)(P.q & 7) : 0u;
    t0 = clock64();
    const int reps = P.iters > 0 ? P.iters : 1;
    for (int it = 0; it < reps; ++it)
      for (int kk = 0; kk < 4; ++kk)
        mma_tf32(tmem, make_desc(a0 + kk * 32, 1024, 0), make_desc(b0 + kk * 32, 1024, bo), idesc, (it | kk) ? 1u : 0u);
    mma_commit(&bar);
  }
  const bool ok = mbar_wait_bounded(&bar, 0, 4000000000LL);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) { if (P.cycles) *P.cycles = clock64() - t0; if (!ok) *P.status = 1; }
  if (ok && P.iters <= 0) {
    for (int c0 = 0; c0 < P.N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                     "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                     "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) P.D[(size_t)tid * P.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

static float tf32_round(float x) { uint32_t u; std::memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float y; std::memcpy(&y, &u, 4); return y; }

int main() {
  const int strip_rows = 288;
  std::vector<float> A(128 * 32), X((size_t)strip_rows * 32);
  srand(7);
  for (auto& v : A) v = tf32_round((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : X) v = tf32_round((float)rand() / RAND_MAX - 0.5f);
  float *dA, *dX, *dD; long long* dcyc; int* dst;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaMalloc(&dcyc, 8)); CK(cudaMalloc(&dst, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 1024 + 16384 + (size_t)strip_rows * 128;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int failures = 0;
  const int qs[] = {0, 1, 3, 7, 8, 9, 15};
  for (int N : {128, 256})
    for (int mode = 0; mode < 2; ++mode)
      for (int q : qs) {
        if (mode == 1 && (q & 7) == 0) continue;
        CK(cudaMemset(dD, 0, 128 * 256 * 4)); CK(cudaMemset(dst, 0, 4));
        ProbeParams P{dA, dX, dD, N, strip_rows, q, mode, 0, dcyc, dst};
        k_probe<<<1, 128, smem>>>(P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { std::printf("N=%d q=%d base_off_mode=%d: kernel error %s\n", N, q, mode, cudaGetErrorString(e)); return 3; }
        std::vector<float> D((size_t)128 * N); int st = 0;
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
        double worst = 0, ref_max = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int e2 = 0; e2 < 32; ++e2) acc += (double)A[m * 32 + e2] * (double)X[(size_t)(n + q) * 32 + e2];
            worst = std::fmax(worst, std::fabs(acc - (double)D[(size_t)m * N + n]));
            ref_max = std::fmax(ref_max, std::fabs(acc));
          }
        const bool good = st == 0 && worst < 1e-5 * ref_max;
        std::printf("N=%3d q=%2d base_offset=%s : status %d  max|err| %.3e (ref max %.3f)  %s\n", N, q, mode ? "q%8" : "0  ", st, worst, ref_max,
                    good ? "MATCH" : "differs");
        if (!good && mode == 0 && q == 0) ++failures;
      }
  for (int N : {128, 256}) {
    for (int q : {0, 1, 4, 8, 13})
    for (int iters : {2048}) {
      CK(cudaMemset(dst, 0, 4));
      ProbeParams P{dA, dX, dD, N, strip_rows, q, 0, iters, dcyc, dst};
      k_probe<<<1, 128, smem>>>(P);
      CK(cudaDeviceSynchronize());
      long long cyc = 0; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
      std::printf("timing N=%3d window shift q=%2d: %d x 4 MMAs (128 x %d x 8 tf32, SS) in %lld cycles = %.1f cycles per MMA\n", N, q, iters, N, cyc, (double)cyc / (iters * 4.0));
    }
  }
  std::printf(failures ? "PROBE FAILED\n" : "PROBE DONE\n");
  return failures ? 1 : 0;
}
