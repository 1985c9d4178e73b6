// tc_probe.cu — measurement tool (not part of the product library): checks the tcgen05 building blocks the
// tensor-core FIR (gsdr_b200/csrc/fir_tc_kernel.cuh) relies on, on one CTA, against a CPU computation:
//   * instruction descriptor for kind::tf32, M = 128, N = 16..64, FP32 accumulate in TMEM;
//   * the B operand as a K-major, un-swizzled shared-memory descriptor whose rows sit at a 16-byte pitch
//     (SBO = 128 B) and whose start address moves in 16-byte steps — the banded-Toeplitz tap panels;
//   * the A operand from tensor memory (written with tcgen05.st, lane = row, column = k) and from shared memory;
//   * what the hardware does with the 13 low mantissa bits of an FP32 operand (TF32), and the error of the
//     3-pass split (hi*hi + hi*lo + lo*hi) against double;
//   * MMA issue rate for the small-N shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t st_ = (x);                                                                     \
    if (st_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d: %s\n", cudaGetErrorName(st_), __FILE__, __LINE__, #x);   \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ unsigned smemU32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smemU32(bar)),
      "r"(parity)
      : "memory");
}

// kind::tf32, FP32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t idescTf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// un-swizzled K-major shared-memory operand: rows of a core matrix 16 B apart, core matrices along M/N `sbo`
// bytes apart, along K `lbo` bytes apart
__device__ __forceinline__ uint64_t smemDesc(unsigned addr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mmaTS(uint32_t dTmem, uint32_t aTmem, uint64_t bDesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(dTmem),
      "r"(aTmem), "l"(bDesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mmaSS(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(dTmem),
      "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mmaCommit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemU32(bar))
               : "memory");
}
__device__ __forceinline__ void tmemSt8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmemLd16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// D[128 x N] = A[128 x K] * B[K x N] (+ the split passes).  mode bit 0: A from shared memory instead of TMEM;
// passes: 1 = A*B as given, 3 = hi/lo split of both operands (Ahi*Bhi + Ahi*Blo + Alo*Bhi).
// B is stored as k-panels: panel j (4 consecutive k) is an array over n of 16-byte entries, panels `lboB` bytes
// apart, and the array starts `shift` entries into its allocation (start address not 128-byte aligned).
template <int N>
__global__ void __launch_bounds__(128) probeKernel(const float* __restrict__ A, const float* __restrict__ B,
                                                   float* __restrict__ D, int K, int mode, int passes, int shift) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmemBase;
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  const int nk = K / 8;            // MMA steps
  const unsigned lboB = (N + 8) * 16;  // deliberately not a multiple of 128
  float* bHi = reinterpret_cast<float*>(smem);                        // [2*nk panels][N + 8 entries][4]
  float* bLo = bHi + (size_t)2 * nk * (N + 8) * 4;
  float* aHi = bLo + (size_t)2 * nk * (N + 8) * 4;                    // SS mode: [2*nk panels][128 rows][4]
  float* aLo = aHi + (size_t)2 * nk * 128 * 4;
  if (tid == 0) {
    mbarInit(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBase)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B panels (hi / lo parts)
  for (int i = tid; i < 2 * nk * N * 4; i += 128) {
    const int e = i & 3, n = (i >> 2) % N, j = (i >> 2) / N;
    const float v = B[(size_t)(4 * j + e) * N + n];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    bHi[((size_t)j * (N + 8) + shift + n) * 4 + e] = passes == 3 ? hi : v;
    bLo[((size_t)j * (N + 8) + shift + n) * 4 + e] = v - hi;
  }
  if (mode & 1) {
    for (int i = tid; i < 2 * nk * 128 * 4; i += 128) {
      const int e = i & 3, m = (i >> 2) % 128, j = (i >> 2) / 128;
      const float v = A[(size_t)m * K + 4 * j + e];
      const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      aHi[((size_t)j * 128 + m) * 4 + e] = passes == 3 ? hi : v;
      aLo[((size_t)j * 128 + m) * 4 + e] = v - hi;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmemBase;
  const uint32_t colD = 0, colAhi = 64, colAlo = 64 + 8 * nk;  // nk <= 8
  if (!(mode & 1)) {
    // A -> TMEM: thread m owns row m (= TMEM lane m); 8 columns per MMA step
    const uint32_t lane = (warp * 32u) << 16;
    for (int s = 0; s < nk; s++) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const float v = A[(size_t)tid * K + 8 * s + e];
        const uint32_t h = __float_as_uint(v) & 0xFFFFE000u;
        hi[e] = passes == 3 ? h : __float_as_uint(v);
        lo[e] = __float_as_uint(v - __uint_as_float(h));
      }
      tmemSt8(tb + lane + colAhi + 8 * s, hi);
      tmemSt8(tb + lane + colAlo + 8 * s, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = idescTf32(128, N);
    uint32_t acc = 0;
    for (int pass = 0; pass < passes; pass++) {
      const bool aIsLo = pass == 2, bIsLo = pass == 1;
      for (int s = 0; s < nk; s++) {
        const float* bp = (bIsLo ? bLo : bHi) + ((size_t)(2 * s) * (N + 8) + shift) * 4;
        const uint64_t bDesc = smemDesc(smemU32(bp), lboB, 128);
        if (mode & 1) {
          const float* ap = (aIsLo ? aLo : aHi) + (size_t)(2 * s) * 128 * 4;
          mmaSS(tb + colD, smemDesc(smemU32(ap), 128 * 16, 128), bDesc, idesc, acc);
        } else {
          mmaTS(tb + colD, tb + (aIsLo ? colAlo : colAhi) + 8 * s, bDesc, idesc, acc);
        }
        acc = 1;
      }
    }
    mmaCommit(&bar);
  }
  mbarWait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmemLd16(tb + ((warp * 32u) << 16) + colD + c, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 16; e++) D[(size_t)tid * N + c + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256) : "memory");
  }
}

// Issue-rate probe: `iters` x (64 MMAs + one commit/wait) per CTA, operands never change.  The loop is warp-uniform
// with one elected lane issuing (as CUTLASS does), descriptors are precomputed, and the MMAs rotate over `nacc`
// accumulators: nacc = 1 makes every MMA depend on the previous one (same TMEM tile).
template <int N>
__global__ void __launch_bounds__(128) rateKernel(int iters, int nacc, int mode, unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmemBase;
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  for (unsigned i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) {
    mbarInit(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBase)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmemBase;
  if (warp == 0) {
    unsigned leader;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
    const uint32_t idesc = idescTf32(128, N);
    const uint64_t bDesc = smemDesc(smemU32(smem), 1024, 128);
    const uint64_t aDesc = smemDesc(smemU32(smem) + 16384, 2048, 128);
    const uint32_t aT = tb + 448;  // A operand columns (TS mode)
    const long long t0 = clock64();
    unsigned parity = 0;
    for (int it = 0; it < iters; it++) {
      if (leader) {
#pragma unroll
        for (int s = 0; s < 64; s++) {
          const uint32_t d = tb + (uint32_t)N * (uint32_t)(s % nacc);
          if (mode & 1) {
            mmaSS(d, aDesc, bDesc, idesc, 1);
          } else {
            mmaTS(d, aT, bDesc, idesc, 1);
          }
        }
        mmaCommit(&bar);
      }
      __syncwarp();
      mbarWait(&bar, parity);
      parity ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && leader) cycles[0] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

static double tf32Trunc(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return (double)r;
}

template <int N>
static int runProbe(int K, int mode, int passes, int shift, bool exactData) {
  std::vector<float> A(128 * K), B((size_t)K * N), D(128 * N, -777.0f);
  uint32_t rng = 12345u + K * 7 + N;
  auto rnd = [&]() {
    rng = rng * 1664525u + 1013904223u;
    return (float)((rng >> 8) & 0xFFFF) / 65536.0f - 0.5f;
  };
  for (int m = 0; m < 128; m++)
    for (int k = 0; k < K; k++) A[m * K + k] = exactData ? (float)((m % 7) - 3) + 0.25f * (float)(k % 5) : rnd();
  for (int k = 0; k < K; k++)
    for (int n = 0; n < N; n++) B[(size_t)k * N + n] = exactData ? 0.5f * (float)(((k * 3 + n) % 11) - 5) : rnd();
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  const int nk = K / 8;
  const size_t smemBytes = (size_t)4 * nk * (N + 8) * 16 + (size_t)4 * nk * 128 * 16 + 1024;
  CK(cudaFuncSetAttribute(probeKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  probeKernel<N><<<1, 128, smemBytes>>>(dA, dB, dD, K, mode, passes, shift);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double errExact = 0, errTrunc = 0, ref2 = 0;
  int bad = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++) {
      double full = 0, tr = 0;
      for (int k = 0; k < K; k++) {
        full += (double)A[m * K + k] * (double)B[(size_t)k * N + n];
        tr += tf32Trunc(A[m * K + k]) * tf32Trunc(B[(size_t)k * N + n]);
      }
      const double got = D[m * N + n];
      errExact = fmax(errExact, fabs(got - full));
      errTrunc = fmax(errTrunc, fabs(got - tr));
      ref2 = fmax(ref2, fabs(full));
      if (exactData && fabs(got - full) > 1e-4 && bad < 6) {
        printf("    mismatch D[%d][%d] = %g, expected %g\n", m, n, got, full);
        bad++;
      }
    }
  printf("N=%d K=%d A-from-%s passes=%d shift=%d %s: max|D - exact| = %.3e, max|D - trunc-tf32 product| = %.3e (max|D| %.2f)\n",
         N, K, (mode & 1) ? "smem" : "tmem", passes, shift, exactData ? "exact-data" : "random", errExact, errTrunc, ref2);
  cudaFree(dA), cudaFree(dB), cudaFree(dD);
  return (exactData && errExact > 1e-4) ? 1 : 0;
}

template <int N>
static void runRate(int mode, int grid, int nacc) {
  unsigned long long* dc;
  CK(cudaMalloc(&dc, 8));
  CK(cudaFuncSetAttribute(rateKernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int iters = 200, steps = 64;
  rateKernel<N><<<grid, 128, 64 * 1024>>>(iters, nacc, mode, dc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rateKernel<N><<<grid, 128, 64 * 1024>>>(iters, nacc, mode, dc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long cyc;
  CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  const double perMma = (double)cyc / (iters * steps);
  const double tflops = 2.0 * 128 * N * 8 * (double)iters * steps * grid / (ms * 1e-3) / 1e12;
  printf("rate N=%d A-from-%s accumulators=%d grid=%d: %.1f cycles per MMA (M128 N%d K8), %.3f ms, %.1f dense TF32 TFLOP/s\n",
         N, (mode & 1) ? "smem" : "tmem", nacc, grid, perMma, N, ms, tflops);
  cudaFree(dc);
}

int main() {
  int fails = 0;
  // 1. layouts, exact data
  fails += runProbe<32>(16, 0, 1, 0, true);
  fails += runProbe<32>(16, 0, 1, 3, true);
  fails += runProbe<32>(64, 0, 1, 5, true);
  fails += runProbe<32>(16, 1, 1, 3, true);
  fails += runProbe<16>(32, 0, 1, 1, true);
  fails += runProbe<64>(32, 0, 1, 2, true);
  // 2. precision: one pass on raw FP32 bits, then the 3-pass split
  runProbe<32>(64, 0, 1, 0, false);
  runProbe<32>(64, 0, 3, 0, false);
  runProbe<32>(64, 1, 3, 0, false);
  // 3. issue rate
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int nacc : {1, 2, 4, 8}) runRate<32>(0, sms, nacc);
  for (int nacc : {1, 2, 4}) runRate<64>(0, sms, nacc);
  for (int nacc : {1, 2}) runRate<128>(0, sms, nacc);
  runRate<16>(0, sms, 8);
  for (int nacc : {1, 4}) runRate<32>(1, sms, nacc);
  runRate<128>(1, sms, 2);
  printf(fails ? "PROBE FAILED (%d layout cases)\n" : "PROBE OK\n", fails);
  return fails ? 1 : 0;
}
