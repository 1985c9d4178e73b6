// fir_kernels.cuh — sm_100a kernels for the decimating FIR family and the fused NCO mix-down.
//
// Replaces the reference kernels k_Fir / k_FirDecimate (ref: src/fir.cu:26-71) and the device function
// k_AdjustFrequency (ref: src/adjustFrequency.cu:25-56).  Not a port: the reference runs one thread per output
// straight out of global memory; here each CTA stages a sample window once into shared memory in a
// polyphase ("phase-major") layout and every thread keeps R consecutive outputs and a sliding sample window
// in registers, so the inner loop is FFMA2 (fma.rn.f32x2: one complex sample x one real tap per instruction)
// fed by one LDS.128 per 2 samples / 4 taps.
//
// Math.  out[n] = sum_{i<T} x[n*D + i] * h[i].  Write i = j*D + p (phase p in [0,D), j in [0,J), J = ceil(T/D)):
//     out[n] = sum_p sum_j h[j*D+p] * x[(n+j)*D + p] = sum_p sum_j hs[p][j] * xs[p][n + j]
// with xs[p][m] = x[m*D + p] and hs[p][j] = h[j*D + p] (0 where j*D+p >= T).  For a fixed phase this is a
// stride-1 FIR, so a thread owning outputs n0..n0+R-1 slides a register window over xs[p][n0 ...]: each new
// sample is used by R FMAs and each tap by R FMAs.
//
// Shared-memory layout (per CTA, tile of BOUT = R*THREADS outputs):
//     hs[D][Jpad]   floats, Jpad = J rounded up to a multiple of R, zero padded
//     xs[D][pitch]  float2; element m of a row lives at position m + PAD*(m/R)  (PAD = 2 float2 of padding per R
//                   elements, so a thread's 16-byte window loads are bank-conflict free: the per-thread stride is
//                   (R+PAD)/2 = odd number of 16-byte units)
// Rows hold BOUT + Jpad elements; samples past the caller-guaranteed input extent are zero-filled
// (cp.async src-size 0), never read.
//
// FF (real input) reuses the same complex core on two half-tiles at once: element.x comes from the first half
// of a 2*BOUT-output tile, element.y from the second half, so one FFMA2 advances two real outputs.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gsdr_b200 {

enum PolyMode : int {
  kPolyFC = 0,         // complex input, real taps
  kPolyFF = 1,         // real input, real taps (two half-tiles packed as .x/.y)
  kPolyNcoExact = 2,   // FC with the 64-bit-phase NCO applied while staging
  kPolyNcoLiteral = 3  // FC with the reference's literal phase arithmetic applied while staging
};

struct PolyParams {
  const void* x;   // float2* (FC/NCO) or float* (FF)
  const float* h;  // taps
  void* y;         // float2* (FC/NCO) or float* (FF)
  unsigned long long nOut;     // outputs per channel
  unsigned long long nIn;      // valid input elements per channel: (nOut-1)*D + T
  unsigned long long xStride;  // channel strides, in elements
  unsigned long long yStride;
  unsigned long long hStride;  // 0 = taps shared by all channels
  unsigned tilesPerChannel;
  unsigned D, T, Jpad, pitch;
  unsigned stageElems;  // (BOUT + Jpad) * D elements staged per (half-)tile
  unsigned dp, dm;      // THREADS % D, THREADS / D
  unsigned posStep;     // shared-memory position step per staging iteration on the constant-stride path
  unsigned fastStage;   // 1 when dp == 0 and dm % R == 0 (every thread keeps its phase; positions advance uniformly)
  unsigned y16;         // output base and channel stride are 16-byte aligned
  // NCO (kPolyNco*)
  unsigned long long ncoStep;   // phase increment per sample, cycles * 2^64
  unsigned long long ncoFirst;  // absolute sample index of input[0]
  unsigned ncoFirst32;          // literal mode: (uint32_t)fmodf((float)firstSampleIndex, fs), ref: src/fm.cu:202
  float ncoFs, ncoF;
};

constexpr int kPolyPad = 2;

__host__ __device__ constexpr unsigned polyPos(unsigned m, unsigned R) { return m + kPolyPad * (m / R); }

__device__ __forceinline__ void cpAsync8(void* smemDst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpAsync8z(void* smemDst, const void* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpAsync4(void* smemDst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpAsync4z(void* smemDst, const void* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpAsyncCommitWaitAll() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// Stage `count` consecutive elements starting at global element index `g0` of channel base `src` into the
// phase-major rows.  ELEM = 8 (float2 -> whole element) or 4 (float -> component `comp` of the element).
template <int ELEM, int R, int THREADS>
__device__ __forceinline__ void stageWindow(
    float2* xs, const unsigned char* src, unsigned long long g0, unsigned long long nIn, const PolyParams& P,
    unsigned comp) {
  const unsigned tid = threadIdx.x;
  const unsigned total = P.stageElems;
  unsigned p = tid % P.D;
  unsigned m = tid / P.D;
  unsigned char* dstBase = reinterpret_cast<unsigned char*>(xs) + comp * 4u;
  const bool interior = g0 + total <= nIn;
  if (interior && P.fastStage) {
    // Every thread stays on its phase row and its position advances by a constant: no index arithmetic.
    const unsigned char* g = src + (g0 + tid) * ELEM;
    unsigned char* d = dstBase + (size_t)(p * P.pitch + polyPos(m, R)) * 8u;
    const unsigned dstep = P.posStep * 8u;
#pragma unroll 8
    for (unsigned s = tid; s < total; s += THREADS) {
      if (ELEM == 8) {
        cpAsync8(d, g);
      } else {
        cpAsync4(d, g);
      }
      g += (size_t)THREADS * ELEM;
      d += dstep;
    }
  } else {
    for (unsigned s = tid; s < total; s += THREADS) {
      const unsigned long long g = g0 + s;
      const bool valid = g < nIn;
      unsigned char* d = dstBase + (size_t)(p * P.pitch + polyPos(m, R)) * 8u;
      const unsigned char* gp = src + (valid ? g : 0ull) * ELEM;
      if (ELEM == 8) {
        cpAsync8z(d, gp, valid);
      } else {
        cpAsync4z(d, gp, valid);
      }
      p += P.dp;
      m += P.dm;
      if (p >= P.D) {
        p -= P.D;
        m += 1;
      }
    }
  }
}

// In-place NCO mix of the staged rows (runs after the async copies have landed, before the FIR loop).
// Each thread walks elements of one row at a stride; element (p, m) is input sample in0 + m*D + p.
template <int MODE, int R, int THREADS>
__device__ __forceinline__ void mixWindow(float2* xs, unsigned long long in0, unsigned rowLen, const PolyParams& P) {
  const unsigned tid = threadIdx.x;
  const unsigned total = P.D * rowLen;
  for (unsigned e = tid; e < total; e += THREADS) {
    const unsigned p = e / rowLen;
    const unsigned m = e - p * rowLen;
    const unsigned long long s = in0 + (unsigned long long)m * P.D + p;  // index relative to input[0]
    float2* q = xs + (size_t)p * P.pitch + polyPos(m, R);
    float sn, cs;
    if (MODE == kPolyNcoExact) {
      const unsigned long long phase = (P.ncoFirst + s) * P.ncoStep;  // mod 2^64
      const float v = (float)(int)(unsigned)(phase >> 32) * 4.656612873077392578125e-10f;  // theta / pi in [-1, 1]
      sincospif(v, &sn, &cs);
    } else {
      // ref: src/adjustFrequency.cu:23,35-50, with the uint32 wrap of ref: src/fm.cu:43-47
      const unsigned idx = P.ncoFirst32 + (unsigned)s;
      const float period = __frcp_rn(P.ncoF);
      const float t = __fdiv_rn(fmodf(__uint2float_rn(idx), P.ncoFs), P.ncoFs);
      const float u = fmodf(t, period);
      sincospif(u * 2.0f, &sn, &cs);
    }
    const float2 v = *q;
    float2 r;
    r.x = __fmaf_rn(v.x, cs, -__fmul_rn(v.y, sn));
    r.y = __fmaf_rn(v.x, sn, __fmul_rn(v.y, cs));
    *q = r;
  }
}

template <int MODE, int R, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) firPolyKernel(const PolyParams P) {
  static_assert(R % 4 == 0 && R >= 4, "R must be a multiple of 4");
  constexpr unsigned BOUT = R * THREADS;
  constexpr bool kReal = (MODE == kPolyFF);
  extern __shared__ __align__(16) unsigned char smemRaw[];
  float* hs = reinterpret_cast<float*>(smemRaw);                                       // [D][Jpad] (+R slack)
  float2* xs = reinterpret_cast<float2*>(smemRaw + ((size_t)P.D * P.Jpad + R) * 4u);  // [D][pitch]

  const unsigned tid = threadIdx.x;
  const unsigned chan = blockIdx.x / P.tilesPerChannel;
  const unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  const unsigned long long o0 = (unsigned long long)tile * (kReal ? 2u * BOUT : BOUT);
  const unsigned long long in0 = o0 * P.D;
  const float* h = P.h + (size_t)chan * P.hStride;

  // ---- stage the sample window (async, no registers) ----
  if (kReal) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.x) + (size_t)chan * P.xStride * 4u;
    stageWindow<4, R, THREADS>(xs, src, in0, P.nIn, P, 0u);
    stageWindow<4, R, THREADS>(xs, src, in0 + (unsigned long long)BOUT * P.D, P.nIn, P, 1u);
  } else {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.x) + (size_t)chan * P.xStride * 8u;
    stageWindow<8, R, THREADS>(xs, src, in0, P.nIn, P, 0u);
  }
  // ---- taps -> phase-major, zero padded (overlaps with the copies in flight) ----
  {
    const unsigned nh = P.D * P.Jpad;
    for (unsigned i = tid; i < nh + R; i += THREADS) {
      const unsigned p = i / P.Jpad;
      const unsigned j = i - p * P.Jpad;
      const unsigned ti = j * P.D + p;
      hs[i] = (i < nh && ti < P.T) ? __ldg(h + ti) : 0.0f;
    }
  }
  cpAsyncCommitWaitAll();
  __syncthreads();
  if (MODE == kPolyNcoExact || MODE == kPolyNcoLiteral) {
    mixWindow<MODE, R, THREADS>(xs, in0, BOUT + P.Jpad, P);
    __syncthreads();
  }

  // ---- polyphase FIR: R outputs per thread, sliding register window ----
  float2 acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = make_float2(0.0f, 0.0f);

  const float2* xrow = xs + (size_t)tid * (R + kPolyPad);
  const float* hp = hs;
  float4 hnext[R / 4];
#pragma unroll
  for (int k = 0; k < R / 4; k++) hnext[k] = *reinterpret_cast<const float4*>(hp + 4 * k);

  for (unsigned p = 0; p < P.D; p++, xrow += P.pitch) {
    const float2* xp = xrow;
    float2 w[R];
#pragma unroll
    for (int k = 0; k < R; k += 2) {
      const float4 v = *reinterpret_cast<const float4*>(xp + k);
      w[k] = make_float2(v.x, v.y);
      w[k + 1] = make_float2(v.z, v.w);
    }
    xp += R + kPolyPad;
    for (unsigned j0 = 0; j0 < P.Jpad; j0 += R, xp += R + kPolyPad) {
      float hv[R];
#pragma unroll
      for (int k = 0; k < R / 4; k++) {
        hv[4 * k + 0] = hnext[k].x;
        hv[4 * k + 1] = hnext[k].y;
        hv[4 * k + 2] = hnext[k].z;
        hv[4 * k + 3] = hnext[k].w;
      }
      hp += R;  // rows of hs are contiguous, so this also walks into the next phase; R floats of slack at the end
#pragma unroll
      for (int k = 0; k < R / 4; k++) hnext[k] = *reinterpret_cast<const float4*>(hp + 4 * k);
#pragma unroll
      for (int jj = 0; jj < R; jj++) {
        const float2 hh = make_float2(hv[jj], hv[jj]);
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = __ffma2_rn(w[(jj + r) % R], hh, acc[r]);
        if (jj & 1) {
          const float4 v = *reinterpret_cast<const float4*>(xp + (jj - 1));
          w[jj - 1] = make_float2(v.x, v.y);
          w[jj] = make_float2(v.z, v.w);
        }
      }
    }
  }

  // ---- store ----
  const unsigned long long ob = o0 + (unsigned long long)tid * R;
  if (!kReal) {
    float2* y = reinterpret_cast<float2*>(P.y) + (size_t)chan * P.yStride;
    if (P.y16 && ob + R <= P.nOut) {
#pragma unroll
      for (int r = 0; r < R; r += 2) {
        *reinterpret_cast<float4*>(y + ob + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (ob + r < P.nOut) y[ob + r] = acc[r];
      }
    }
  } else {
    float* y = reinterpret_cast<float*>(P.y) + (size_t)chan * P.yStride;
    const unsigned long long ob2 = ob + BOUT;
    if (P.y16 && ob2 + R <= P.nOut) {
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        *reinterpret_cast<float4*>(y + ob + r) = make_float4(acc[r].x, acc[r + 1].x, acc[r + 2].x, acc[r + 3].x);
        *reinterpret_cast<float4*>(y + ob2 + r) = make_float4(acc[r].y, acc[r + 1].y, acc[r + 2].y, acc[r + 3].y);
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (ob + r < P.nOut) y[ob + r] = acc[r].x;
        if (ob2 + r < P.nOut) y[ob2 + r] = acc[r].y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Direct kernel: any type combination, any D/T/alignment.  One output per thread, taps staged through shared
// memory in chunks, samples through the read-only path.  Accumulates in the reference's own order with the
// reference's own expression shapes (ref: src/fir.cu:64-70; src/cuComplexOperatorOverloads.cuh:25-31,57-62),
// so it reproduces the reference's bits; used for CC/CF and for shapes the polyphase kernel cannot hold in
// shared memory.
// ---------------------------------------------------------------------------------------------------------
struct DirectParams {
  const void* x;
  const void* h;
  void* y;
  unsigned long long nOut, D, T;
  unsigned long long xStride, yStride, hStride;
  unsigned blocksPerChannel;
};

constexpr int kDirectThreads = 256;
constexpr int kDirectTapChunk = 2048;

__device__ __forceinline__ void directMac(float& acc, float x, float h) { acc = __fmaf_rn(x, h, acc); }
__device__ __forceinline__ void directMac(float2& acc, float2 x, float h) {
  acc.x = __fmaf_rn(x.x, h, acc.x);
  acc.y = __fmaf_rn(x.y, h, acc.y);
}
__device__ __forceinline__ void directMac(float2& acc, float x, float2 h) {
  acc.x = __fmaf_rn(h.x, x, acc.x);
  acc.y = __fmaf_rn(h.y, x, acc.y);
}
__device__ __forceinline__ void directMac(float2& acc, float2 x, float2 h) {
  const float t1 = __fmul_rn(x.y, h.y);
  const float t2 = __fmul_rn(x.x, h.y);
  const float pre = __fmaf_rn(x.x, h.x, -t1);
  const float pim = __fmaf_rn(x.y, h.x, t2);
  acc.x = __fadd_rn(acc.x, pre);
  acc.y = __fadd_rn(acc.y, pim);
}
template <class T>
__device__ __forceinline__ T directZero();
template <>
__device__ __forceinline__ float directZero<float>() {
  return 0.0f;
}
template <>
__device__ __forceinline__ float2 directZero<float2>() {
  return make_float2(0.0f, 0.0f);
}

template <class IN_T, class OUT_T, class TAP_T>
__global__ void __launch_bounds__(kDirectThreads) firDirectKernel(const DirectParams P) {
  __shared__ TAP_T hs[kDirectTapChunk];
  const unsigned chan = blockIdx.x / P.blocksPerChannel;
  const unsigned blk = blockIdx.x - chan * P.blocksPerChannel;
  const unsigned long long n = (unsigned long long)blk * kDirectThreads + threadIdx.x;
  const bool live = n < P.nOut;
  const IN_T* x = reinterpret_cast<const IN_T*>(P.x) + (size_t)chan * P.xStride + (live ? n * P.D : 0ull);
  const TAP_T* h = reinterpret_cast<const TAP_T*>(P.h) + (size_t)chan * P.hStride;
  OUT_T acc = directZero<OUT_T>();
  for (unsigned long long t0 = 0; t0 < P.T; t0 += kDirectTapChunk) {
    const unsigned cnt = (unsigned)((P.T - t0 < (unsigned long long)kDirectTapChunk) ? (P.T - t0) : kDirectTapChunk);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < cnt; i += kDirectThreads) hs[i] = h[t0 + i];
    __syncthreads();
    if (live) {
      const IN_T* xp = x + t0;
#pragma unroll 4
      for (unsigned i = 0; i < cnt; i++) directMac(acc, __ldg(xp + i), hs[i]);
    }
  }
  if (live) reinterpret_cast<OUT_T*>(P.y)[(size_t)chan * P.yStride + n] = acc;
}

}  // namespace gsdr_b200
