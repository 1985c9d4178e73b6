// tc_probe_f16.cu — measurement tool: tcgen05.mma kind::f16 with the A operand in tensor memory (two K elements per
// 32-bit column?) and the B operand as an un-swizzled K-major descriptor with 16-byte row pitch; exact-data layout
// check against the CPU, plus the issue rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe_f16 ...
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t st_ = (x); if (st_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorName(st_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned smemU32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smemU32(bar)), "r"(parity) : "memory");
}
// kind::f16: A, B = F16 (format 0), FP32 accumulate, both K-major
__host__ __device__ constexpr uint32_t idescF16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t smemDesc(unsigned addr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mmaTS(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mmaCommit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemU32(bar)) : "memory");
}

// D[128 x N] = A[128 x K] * B[K x N], K = 16 * steps.  A: row m in TMEM lane m, 8 columns per step, column j of a step
// = (k = 2j in the low half, k = 2j + 1 in the high half).  B: panels of 8 consecutive k (16 bytes per n), rows at a
// 16-byte pitch, panels `lbo` apart, start shifted by `shift` entries.
template <int N>
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int K, int shift) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmemBase;
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  const int steps = K / 16;
  const unsigned lbo = (N + 8) * 16;
  __half* bp = reinterpret_cast<__half*>(smem);  // [2*steps panels][N + 8][8]
  if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBase)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * steps * N * 8; i += 128) {
    const int e = i & 7, n = (i >> 3) % N, j = (i >> 3) / N;
    bp[((size_t)j * (N + 8) + shift + n) * 8 + e] = __float2half(B[(size_t)(8 * j + e) * N + n]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmemBase, colA = 64;
  for (int s = 0; s < steps; s++) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const __half2 h = __floats2half2_rn(A[(size_t)tid * K + 16 * s + 2 * j], A[(size_t)tid * K + 16 * s + 2 * j + 1]);
      v[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tb + ((warp * 32u) << 16) + colA + 8 * s),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int s = 0; s < steps; s++) {
      const __half* p0 = bp + ((size_t)(2 * s) * (N + 8) + shift) * 8;
      mmaTS(tb, tb + colA + 8 * s, smemDesc(smemU32(p0), lbo, 128), idescF16(128, N), s > 0);
    }
    mmaCommit(&bar);
  }
  mbarWait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tb + ((warp * 32u) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 8; e++) D[(size_t)tid * N + c + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128) rate(int iters, unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmemBase;
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  for (unsigned i = tid; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBase)), "r"(256) : "memory");
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
    const uint64_t bDesc = smemDesc(smemU32(smem), 1024, 128);
    const long long t0 = clock64();
    unsigned parity = 0;
    for (int it = 0; it < iters; it++) {
      if (leader) {
#pragma unroll
        for (int s = 0; s < 64; s++) mmaTS(tb, tb + 128, bDesc, idescF16(128, N), 1);
        mmaCommit(&bar);
      }
      __syncwarp();
      mbarWait(&bar, parity);
      parity ^= 1;
    }
    if (blockIdx.x == 0 && leader) cycles[0] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256) : "memory");
}

template <int N>
static int run(int K, int shift) {
  std::vector<float> A(128 * K), B((size_t)K * N), D(128 * N, -7.f);
  for (int m = 0; m < 128; m++) for (int k = 0; k < K; k++) A[m * K + k] = (float)((m % 7) - 3) + 0.25f * (float)(k % 5);
  for (int k = 0; k < K; k++) for (int n = 0; n < N; n++) B[(size_t)k * N + n] = 0.5f * (float)(((k * 3 + n) % 11) - 5);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const size_t smemBytes = (size_t)2 * (K / 16) * (N + 8) * 16 + 1024;
  probe<N><<<1, 128, smemBytes>>>(dA, dB, dD, K, shift);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double err = 0; int bad = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
    double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * B[(size_t)k * N + n];
    const double e = fabs(D[m * N + n] - ref); if (e > err) err = e;
    if (e > 1e-3 && bad < 4) { printf("   D[%d][%d] = %g, expected %g\n", m, n, D[m * N + n], ref); bad++; }
  }
  printf("f16 N=%d K=%d shift=%d: max|D - exact| = %.3e\n", N, K, shift, err);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err > 1e-3;
}

template <int N>
static void runRate(int grid) {
  unsigned long long* dc; CK(cudaMalloc(&dc, 8));
  CK(cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  rate<N><<<grid, 128, 64 * 1024>>>(200, dc); CK(cudaDeviceSynchronize());
  rate<N><<<grid, 128, 64 * 1024>>>(200, dc); CK(cudaDeviceSynchronize());
  unsigned long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  printf("rate kind::f16 N=%d A-from-tmem: %.1f cycles per MMA (M128 N%d K16)\n", N, (double)cyc / (200 * 64), N);
  cudaFree(dc);
}

int main() {
  int fails = 0;
  fails += run<32>(16, 0);
  fails += run<32>(64, 3);
  fails += run<16>(32, 1);
  fails += run<64>(32, 2);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  runRate<16>(sms); runRate<32>(sms); runRate<64>(sms); runRate<128>(sms);
  printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
  return fails;
}
