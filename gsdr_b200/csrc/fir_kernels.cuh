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
//     hs[D][Jpad]   floats, Jpad = J rounded up to a multiple of 2R, zero padded (+2R floats of zero slack)
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

// Work-skipping measurement hooks (skip the copies / the FIR loop / the stores) exist only in the tuning build
// (-DGSDR_B200_TUNING, libgsdr_b200_tuning.so); the release library compiles them out.
#ifdef GSDR_B200_TUNING
#define GSDR_DBG(P) ((P).dbg)
#else
#define GSDR_DBG(P) 0u
#endif

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
  unsigned totalTiles;  // tilesPerChannel * numChannels
  unsigned D, T, Jpad, pitch;
  unsigned stageElems;  // (BOUT + Jpad) * D elements staged per (half-)tile
  unsigned dp, dm;      // NT % D, NT / D   (NT = threads per CTA)
  unsigned posStep;     // shared-memory position step per staging iteration on the constant-stride path
  unsigned fastStage;   // 1 when dp == 0 and dm % R == 0 (every thread keeps its phase; positions advance uniformly)
  unsigned y16;         // output base and channel stride are 16-byte aligned
  unsigned dbg;         // measurement hook (gsdrB200SetDebugFlags): bit0 = skip the window copies, bit1 = skip the FIR loop
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

// Stage stageElems consecutive elements starting at global element index `g0` of channel base `src` into the
// phase-major rows of one buffer.  ELEM = 8 (float2 -> whole element) or 4 (float -> component `comp`).
template <int ELEM, int R, int NT>
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
    for (unsigned s = tid; s < total; s += NT) {
      if (ELEM == 8) {
        cpAsync8(d, g);
      } else {
        cpAsync4(d, g);
      }
      g += (size_t)NT * ELEM;
      d += dstep;
    }
  } else {
    for (unsigned s = tid; s < total; s += NT) {
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
template <int MODE, int R, int NT>
__device__ __forceinline__ void mixWindow(float2* xs, unsigned long long in0, unsigned rowLen, const PolyParams& P) {
  const unsigned tid = threadIdx.x;
  const unsigned total = P.D * rowLen;
  for (unsigned e = tid; e < total; e += NT) {
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

__device__ __forceinline__ void loadPair(float2& a, float2& b, const float2* src) {
  const float4 v = *reinterpret_cast<const float4*>(src);
  a = make_float2(v.x, v.y);
  b = make_float2(v.z, v.w);
}
template <int R>
__device__ __forceinline__ void loadSamples(float2 (&x)[R], const float2* src) {
#pragma unroll
  for (int k = 0; k < R; k += 2) loadPair(x[k], x[k + 1], src + k);
}
template <int R>
__device__ __forceinline__ void loadTaps4(float (&h)[R], int off, const float* src) {
  const float4 v = *reinterpret_cast<const float4*>(src);
  h[off] = v.x;
  h[off + 1] = v.y;
  h[off + 2] = v.z;
  h[off + 3] = v.w;
}
template <int R>
__device__ __forceinline__ void loadTaps(float (&h)[R], const float* src) {
#pragma unroll
  for (int k = 0; k < R; k += 4) loadTaps4<R>(h, k, src + k);
}
__device__ __forceinline__ float2 macTap(float2 x, float h, float2 acc) {
  return __ffma2_rn(x, make_float2(h, h), acc);  // FFMA2 acc, x.F32x2, h.F32 (scalar broadcast), acc
}

// The FIR inner loop is SAMPLE-stationary: window element e of a phase (sample xs[p][n0 + e]) is multiplied by
// the R taps h[e - r], r = 0..R-1, into the R accumulators, then it is dead.  Consecutive FFMA2 therefore share
// their sample operand (served by the operand-reuse cache), and each reads only its accumulator pair and one
// tap from the register file: two registers per bank, so FFMA2 issues every 2 cycles.  (Tap-stationary order —
// one tap, R samples — needs 5 distinct registers per FFMA2, 3 of them in one bank, and runs at 2/3 rate: measured
// 61 % FMA-pipe utilisation, profiles/r01_notes.md.)  The taps slide through two register blocks of R.
//
// Elements of one phase (J' = Jpad taps, blocks of R):
//   prologue   e = 0 .. R-1        element e feeds outputs r <= e            (taps block 0)
//   steady b   e = R(b+1) + i      r <= i: new block b+1, slot i-r; r > i: old block b, slot R+i-r
//   tail       e = J' + k, k<R-1   feeds outputs r > k                       (taps block J'/R - 1)

// Block b+1 of samples in `x`; `ho` = tap block b (dies slot by slot, refilled with block b+2 from hNext),
// `hn` = tap block b+1.  Each sample pair is refilled from xNext (same slots, two sample blocks ahead).
template <int R>
__device__ __forceinline__ void firSteady(
    float2 (&acc)[R], float2 (&x)[R], float (&ho)[R], const float (&hn)[R], const float2* xNext, const float* hNext) {
#pragma unroll
  for (int i = 0; i < R; i++) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const float h = (r <= i) ? hn[(r <= i) ? i - r : 0] : ho[(r <= i) ? 0 : R + i - r];
      acc[r] = macTap(x[i], h, acc[r]);
    }
    if (i & 1) loadPair(x[i - 1], x[i], xNext + (i - 1));
    if ((i & 3) == 3) loadTaps4<R>(ho, i - 3, hNext + (i - 3));  // old slots <= i are dead after element i
  }
}
template <int R>
__device__ __forceinline__ void firPrologue(float2 (&acc)[R], const float2 (&x)[R], const float (&h0)[R]) {
#pragma unroll
  for (int e = 0; e < R; e++) {
#pragma unroll
    for (int r = 0; r <= e; r++) acc[r] = macTap(x[e], h0[e - r], acc[r]);
  }
}
template <int R>
__device__ __forceinline__ void firTail(float2 (&acc)[R], const float2 (&x)[R], const float (&hl)[R]) {
#pragma unroll
  for (int k = 0; k < R - 1; k++) {
#pragma unroll
    for (int r = k + 1; r < R; r++) acc[r] = macTap(x[k], hl[R + k - r], acc[r]);
  }
}

// MODE: PolyMode.  R: outputs per thread.  TG: threads per phase group (tile = R*TG outputs).  PSPLIT: phase
// groups; group g accumulates phases [g*D/PSPLIT, (g+1)*D/PSPLIT) and the partial sums are added through shared
// memory.  NBUF: sample-window buffers (2 = the next tile is copied in while this one is filtered).
// Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
template <int MODE, int R, int TG, int PSPLIT, int NBUF, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT, MINB) firPolyKernel(const PolyParams P) {
  static_assert(R % 4 == 0 && R >= 4, "R must be a multiple of 4");
  static_assert(NBUF == 1 || NBUF == 2, "one or two window buffers");
  constexpr unsigned NT = TG * PSPLIT;
  constexpr unsigned BOUT = R * TG;
  constexpr bool kReal = (MODE == kPolyFF);
  constexpr unsigned kTileOut = kReal ? 2u * BOUT : BOUT;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  float* hs = reinterpret_cast<float*>(smemRaw);  // [D][Jpad] (+2R slack)
  float2* xsBase = reinterpret_cast<float2*>(smemRaw + ((size_t)P.D * P.Jpad + 2 * R) * 4u);
  const unsigned bufElems = P.D * P.pitch;  // float2 per buffer

  const unsigned tid = threadIdx.x;
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  const unsigned pBegin = (grp * P.D) / PSPLIT;
  const unsigned pEnd = ((grp + 1) * P.D) / PSPLIT;
  const unsigned elemBytes = kReal ? 4u : 8u;

  auto stageTile = [&](unsigned work, float2* xs) {
    if (GSDR_DBG(P) & 1u) return;
    const unsigned chan = work / P.tilesPerChannel;
    const unsigned tile = work - chan * P.tilesPerChannel;
    const unsigned long long in0 = (unsigned long long)tile * kTileOut * P.D;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.x) + (size_t)chan * P.xStride * elemBytes;
    if (kReal) {
      stageWindow<4, R, NT>(xs, src, in0, P.nIn, P, 0u);
      stageWindow<4, R, NT>(xs, src, in0 + (unsigned long long)BOUT * P.D, P.nIn, P, 1u);
    } else {
      stageWindow<8, R, NT>(xs, src, in0, P.nIn, P, 0u);
    }
  };

  unsigned work = blockIdx.x;
  if (NBUF == 2) {
    if (work < P.totalTiles) stageTile(work, xsBase);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  unsigned tapsChan = 0xffffffffu;

  for (unsigned it = 0; work < P.totalTiles; work += gridDim.x, it++) {
    float2* xs = xsBase + (NBUF == 2 ? (it & 1u) * bufElems : 0u);
    const unsigned chan = work / P.tilesPerChannel;
    const unsigned tile = work - chan * P.tilesPerChannel;
    const unsigned long long o0 = (unsigned long long)tile * kTileOut;

    if (NBUF == 2) {
      const unsigned nextWork = work + gridDim.x;
      if (nextWork < P.totalTiles) stageTile(nextWork, xsBase + ((it + 1u) & 1u) * bufElems);
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    } else {
      stageTile(work, xs);
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    // taps -> phase-major, zero padded; once per CTA unless the channel (and its tap set) changes
    if (chan != tapsChan && (tapsChan == 0xffffffffu || P.hStride != 0)) {
      const float* h = P.h + (size_t)chan * P.hStride;
      const unsigned nh = P.D * P.Jpad;
      unsigned p = tid / P.Jpad;
      unsigned j = tid - p * P.Jpad;
      const unsigned dpj = NT / P.Jpad, djj = NT - dpj * P.Jpad;
      for (unsigned i = tid; i < nh + 2 * R; i += NT) {
        const unsigned ti = j * P.D + p;
        hs[i] = (i < nh && ti < P.T) ? __ldg(h + ti) : 0.0f;
        p += dpj;
        j += djj;
        if (j >= P.Jpad) {
          j -= P.Jpad;
          p += 1;
        }
      }
    }
    tapsChan = chan;
    if (NBUF == 2) {
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    if (MODE == kPolyNcoExact || MODE == kPolyNcoLiteral) {
      mixWindow<MODE, R, NT>(xs, o0 * P.D, BOUT + P.Jpad, P);
      __syncthreads();
    }

    // ---- polyphase FIR: R outputs per thread, sample-stationary, taps sliding through registers ----
    float2 acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = make_float2(0.0f, 0.0f);
    const unsigned pStop = (GSDR_DBG(P) & 2u) ? pBegin : pEnd;
    if (pBegin < pStop) {
      constexpr unsigned BLK = R + kPolyPad;  // float2 positions per block of R samples
      const float2* xrow = xs + (size_t)t * BLK + (size_t)pBegin * P.pitch;
      const float* hp = hs + (size_t)pBegin * P.Jpad;
      const unsigned nbk = P.Jpad / R;  // even, >= 2
      float hA[R], hB[R];
      float2 xA[R], xB[R], xC[R];
      loadTaps<R>(hA, hp);
      loadTaps<R>(hB, hp + R);
      loadSamples<R>(xC, xrow);
      loadSamples<R>(xB, xrow + BLK);
      for (unsigned p = pBegin; p < pStop; p++, xrow += P.pitch, hp += P.Jpad) {
        const bool more = p + 1 < pStop;
        const float2* xrowNext = more ? xrow + P.pitch : xrow;  // last phase: harmless re-read of this row
        firPrologue<R>(acc, xC, hA);
        loadSamples<R>(xA, xrow + 2 * BLK);
        const float2* xq = xrow + 3 * BLK;
        const float* hq = hp + 2 * R;
        for (unsigned b = 0; b + 2 < nbk; b += 2) {
          firSteady<R>(acc, xB, hA, hB, xq, hq);
          firSteady<R>(acc, xA, hB, hA, xq + BLK, hq + R);
          xq += 2 * BLK;
          hq += 2 * R;
        }
        // last steady block of the phase: its refills already fetch the NEXT phase (tap rows are contiguous, so
        // hq points at the next phase's block 0; samples come from the next row's block 1), and block 0 of the
        // next row goes to xC, which has been free since the prologue.
        loadSamples<R>(xC, xrowNext);
        firSteady<R>(acc, xB, hA, hB, xrowNext + BLK, hq);
        firTail<R>(acc, xA, hB);
        loadTaps<R>(hB, hq + R);
      }
    }

    // ---- add the phase groups' partial sums (fixed order: deterministic) ----
    if (PSPLIT > 1) {
      __syncthreads();  // everyone is done reading this window; reuse it as scratch
      float4* red = reinterpret_cast<float4*>(xs);
      if (grp > 0) {
#pragma unroll
        for (int q = 0; q < R / 2; q++) {
          red[((grp - 1) * (R / 2) + q) * TG + t] =
              make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        }
      }
      __syncthreads();
      if (grp == 0) {
#pragma unroll
        for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
          for (int q = 0; q < R / 2; q++) {
            const float4 v = red[((g - 1) * (R / 2) + q) * TG + t];
            acc[2 * q].x += v.x;
            acc[2 * q].y += v.y;
            acc[2 * q + 1].x += v.z;
            acc[2 * q + 1].y += v.w;
          }
        }
      }
    }

    // ---- store ----
    if (grp == 0) {
      const unsigned long long ob = o0 + (unsigned long long)t * R;
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
    __syncthreads();  // the window (and the scratch in it) is free before anyone refills it
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
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

// Direct kernel with the NCO mix-down in front of every tap (any D / T): the fallback for shapes whose window does
// not fit the staged kernels.  One sincospi per tap per output — T/D times more phasor work than the staged kernels,
// which mix every input sample once — but no shared-memory limit.  Same phase laws as mixWindow().
struct DirectNcoParams {
  const float2* x;
  const float* h;
  float2* y;
  unsigned long long nOut, D, T;
  unsigned long long xStride, yStride, hStride;
  unsigned blocksPerChannel;
  unsigned long long ncoStep, ncoFirst;
  unsigned ncoFirst32;
  float ncoFs, ncoF;
};

template <int MODE>
__global__ void __launch_bounds__(kDirectThreads) firDirectNcoKernel(const DirectNcoParams P) {
  __shared__ float hs[kDirectTapChunk];
  const unsigned chan = blockIdx.x / P.blocksPerChannel;
  const unsigned blk = blockIdx.x - chan * P.blocksPerChannel;
  const unsigned long long n = (unsigned long long)blk * kDirectThreads + threadIdx.x;
  const bool live = n < P.nOut;
  const unsigned long long s0 = live ? n * P.D : 0ull;  // index of the window's first sample, relative to input[0]
  const float2* x = P.x + (size_t)chan * P.xStride + s0;
  const float* h = P.h + (size_t)chan * P.hStride;
  float2 acc = make_float2(0.0f, 0.0f);
  for (unsigned long long t0 = 0; t0 < P.T; t0 += kDirectTapChunk) {
    const unsigned cnt = (unsigned)((P.T - t0 < (unsigned long long)kDirectTapChunk) ? (P.T - t0) : kDirectTapChunk);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < cnt; i += kDirectThreads) hs[i] = h[t0 + i];
    __syncthreads();
    if (live) {
      const float2* xp = x + t0;
      unsigned long long phase = (P.ncoFirst + s0 + t0) * P.ncoStep;  // mod 2^64
      for (unsigned i = 0; i < cnt; i++) {
        float sn, cs;
        if (MODE == kPolyNcoExact) {
          sincospif((float)(int)(unsigned)(phase >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
          phase += P.ncoStep;
        } else {
          // ref: src/adjustFrequency.cu:23,35-50, with the uint32 wrap of ref: src/fm.cu:43-47
          const unsigned idx = P.ncoFirst32 + (unsigned)(s0 + t0 + i);
          const float period = __frcp_rn(P.ncoF);
          const float tt = __fdiv_rn(fmodf(__uint2float_rn(idx), P.ncoFs), P.ncoFs);
          sincospif(fmodf(tt, period) * 2.0f, &sn, &cs);
        }
        const float2 v = __ldg(xp + i);
        const float mre = __fmaf_rn(v.x, cs, -__fmul_rn(v.y, sn));
        const float mim = __fmaf_rn(v.x, sn, __fmul_rn(v.y, cs));
        acc.x = __fmaf_rn(hs[i], mre, acc.x);
        acc.y = __fmaf_rn(hs[i], mim, acc.y);
      }
    }
  }
  if (live) P.y[(size_t)chan * P.yStride + n] = acc;
}

// Direct kernel for int8 IQ input (any D / T / alignment): conversion (ref: src/conversion.cu:26) and the optional
// exact NCO per tap; the fallback of firTmaInt8Kernel.
struct DirectInt8Params {
  const signed char* x;  // interleaved I, Q
  const float* h;
  float2* y;
  unsigned long long nOut, D, T;
  unsigned long long xStride, yStride;  // in samples / outputs
  unsigned blocksPerChannel;
  unsigned long long ncoStep, ncoFirst;
};

template <bool NCO>
__global__ void __launch_bounds__(kDirectThreads) firDirectInt8Kernel(const DirectInt8Params P) {
  __shared__ float hs[kDirectTapChunk];
  const unsigned chan = blockIdx.x / P.blocksPerChannel;
  const unsigned blk = blockIdx.x - chan * P.blocksPerChannel;
  const unsigned long long n = (unsigned long long)blk * kDirectThreads + threadIdx.x;
  const bool live = n < P.nOut;
  const unsigned long long s0 = live ? n * P.D : 0ull;
  const signed char* x = P.x + 2u * ((size_t)chan * P.xStride + s0);
  float2 acc = make_float2(0.0f, 0.0f);
  for (unsigned long long t0 = 0; t0 < P.T; t0 += kDirectTapChunk) {
    const unsigned cnt = (unsigned)((P.T - t0 < (unsigned long long)kDirectTapChunk) ? (P.T - t0) : kDirectTapChunk);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < cnt; i += kDirectThreads) hs[i] = P.h[t0 + i];
    __syncthreads();
    if (live) {
      const signed char* xp = x + 2u * t0;
      unsigned long long phase = (P.ncoFirst + s0 + t0) * P.ncoStep;
      for (unsigned i = 0; i < cnt; i++) {
        float re = fmaxf(-1.0f, __fdiv_rn((float)xp[2 * i], 127.0f));
        float im = fmaxf(-1.0f, __fdiv_rn((float)xp[2 * i + 1], 127.0f));
        if (NCO) {
          float sn, cs;
          sincospif((float)(int)(unsigned)(phase >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
          phase += P.ncoStep;
          const float mre = __fmaf_rn(re, cs, -__fmul_rn(im, sn));
          const float mim = __fmaf_rn(re, sn, __fmul_rn(im, cs));
          re = mre, im = mim;
        }
        acc.x = __fmaf_rn(hs[i], re, acc.x);
        acc.y = __fmaf_rn(hs[i], im, acc.y);
      }
    }
  }
  if (live) P.y[(size_t)chan * P.yStride + n] = acc;
}

// ref: src/conversion.cu:20-27 (k_int8ToFloat), with the bounds check the reference gets wrong (x > numElements)
static __global__ void int8ToNormFloatKernel(const signed char* in, float* out, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaxf(-1.0f, __fdiv_rn((float)in[i], 127.0f));
}

}  // namespace gsdr_b200
