// fir_tma_kernel.cuh — the fast path for complex input x real taps (gsdrFirFC and the fused NCO stage):
// a persistent, TMA-fed polyphase FIR for even decimations with rows of at most 128 bytes (D <= 16).
// Replaces ref: src/fir.cu:49-71 (k_FirDecimate<cuComplex,cuComplex,float>) and the per-tap arithmetic of
// ref: src/adjustFrequency.cu:36-55.
//
// Why TMA.  On sm_100a an FFMA2 holds the issue port for two cycles and EVERY other instruction costs one more
// (tools/ubench_fp32.cu, profiles/r01_notes.md), so the FIR is bound by issue slots, not by the FMA pipe alone.
// Staging the window with per-thread cp.async costs ~6 issue cycles per 32 samples; a TMA tensor copy costs
// none: one elected thread issues ONE cp.async.bulk.tensor per tile and the copy engine does the rest at HBM speed
// (6.7 TB/s measured for this box shape, tools/tma_probe.cu).
//
// Shared-memory layout of one window buffer.  The input is viewed as rows of D samples (row m = samples
// m*D .. m*D+D-1, G = 8*D bytes).  A 3-D tensor map (row bytes | mh, stride 8 rows | ml, stride 1 row) with box
// (G bytes, MHP, 8) lands the window as
//        buf[ml][mh][G bytes],        m = 8*mh + ml,
// i.e. eight planes holding every 8th row.  A thread owns outputs n0 = 8*t .. 8*t+7 of the tile, so the sample it
// needs for window element e = 8*c + i is row (ml = i, mh = t + c): consecutive lanes read consecutive mh — G
// bytes apart.  Rows whose 16-byte chunk count (D/2) is odd are bank-conflict free as they are; for D = 4, 8, 16
// the TMA swizzle mode (32/64/128 B) XORs the chunk index with row-address bits and the reader applies the same
// XOR, which makes the eight 16-byte loads of a quarter warp hit eight different bank groups.
//
// Inner loop.  One LDS.128 fetches the samples of two adjacent polyphase branches (2*pp, 2*pp+1) of one row; it
// feeds 2 x 8 FFMA2 (sample-stationary: consecutive FFMA2 share the sample operand).  Taps of the two branches
// are stored interleaved, hs2[pp][j] = (h[j*D+2pp], h[j*D+2pp+1]), contiguous across pairs, and slide through
// two register blocks.  Samples run through a ring of 8 float4 registers, loaded 6 elements ahead.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fir_kernels.cuh"

namespace gsdr_b200 {

struct TmaParams {
  const float2* x;  // channel 0 input (edge tiles and unaligned fallbacks read it directly)
  const float* h;
  float2* y;
  unsigned long long nOut, nIn;
  unsigned long long xStride, yStride, hStride;
  unsigned tilesPerChannel, totalTiles, numChannels;
  unsigned strideChan, strideTile;  // gridDim.x split as strideChan * tilesPerChannel + strideTile
  unsigned D, T, Jpad;
  unsigned rowBytes;     // G = 8*D
  unsigned segBytes;     // min(G, 128): bytes of a row stored contiguously in a plane
  unsigned mhp;          // rows per plane in a buffer (box dim 1)
  unsigned planeBytes;   // mhp * segBytes, a whole number of swizzle periods
  unsigned swzShift;     // address bit the chunk XOR takes its source from, relative to mh: see tmaSwizzle()
  unsigned swzMask;      // 0 (no swizzle), 1, 3 or 7
  unsigned tmaRows;      // rows visible to TMA per channel (multiple of 8); windows reaching past it use cp.async
  unsigned y16;
  unsigned dbg;
  // fused output stage (launch.h: FirEpilogue): 0 = complex outputs, 1 = AM envelope, 2 = FM quadrature demodulation
  unsigned epi;
  float epiGain;     // FM: sampleRate / (2 * pi * deviation)
  unsigned tileOut;  // outputs a tile advances by: 8 * TG, or 8 * TG - 8 for FM (tiles overlap by one row group,
                     // so every output but a tile's last eight finds its successor inside the tile)
  unsigned long long ncoStep, ncoFirst;
  unsigned ncoFirst32;
  float ncoFs, ncoF;
  unsigned long long ncoRow0;  // ncoFirst / D: absolute row number of the channel's first row (exact NCO)
  unsigned ncoRho;             // ncoFirst % D: offset of every row's first sample inside its absolute row
};

// firTmaRealKernel: x, y and the strides count floats, D is twice the caller's decimation
struct RealParams : TmaParams {
  unsigned long long nOutReal;  // the caller's output count (nOut = ceil(nOutReal / 2) output pairs)
  unsigned rawFloats;           // floats of input one tile reads, a multiple of 4
};

constexpr int kTmaR = 8;
constexpr unsigned kTmaJpadCap = 64;  // compile-time-geometry kernels hold up to 64 taps per branch (T <= 64*D)

// Rows per plane of a window buffer: TG row groups for the outputs + Jpad/8 for the taps' reach, rounded so
// that a plane is a whole number of swizzle periods (swizzled rows) or of 128-byte TMA units (plain rows).
// Rows wider than 128 bytes (D > 16, a multiple of 16) are stored as row SEGMENTS of 128 bytes, each segment in
// its own set of 8 planes and fetched by its own tensor copy.
__host__ __device__ constexpr unsigned tmaSegBytes(unsigned D) { return 8u * D > 128u ? 128u : 8u * D; }
__host__ __device__ constexpr unsigned tmaPlaneRows(unsigned tg, unsigned jpad, unsigned D) {
  const unsigned G = tmaSegBytes(D);
  const unsigned unit = (G == 32u) ? 256u : (G == 64u) ? 512u : (G == 128u) ? 1024u : 128u;
  unsigned mhp = tg + jpad / 8u;
  while ((mhp * G) % unit) mhp++;
  return mhp;
}

// XOR applied to the 16-byte chunk index of row-group mh (what the TMA swizzle modes do to address bits 4..6):
//   32B  mode (G = 32):  chunk ^= addr bit 7      = (mh >> 2) & 1
//   64B  mode (G = 64):  chunk ^= addr bits 7..8  = (mh >> 1) & 3
//   128B mode (G = 128): chunk ^= addr bits 7..9  =  mh       & 7
// DT is the compile-time decimation (0 = take everything from the parameter block).
template <int DT>
__device__ __forceinline__ unsigned tmaSwizzle(unsigned mh, const TmaParams& P) {
  if (DT == 4) return (mh >> 2) & 1u;
  if (DT == 8) return (mh >> 1) & 3u;
  if (DT >= 16 && DT % 16 == 0) return mh & 7u;
  if (DT != 0) return 0u;
  return (mh >> P.swzShift) & P.swzMask;
}

// byte offset inside a buffer of branch pair pp of row group mh, plane 0
template <int DT>
__device__ __forceinline__ unsigned tmaPairOffset(unsigned mh, unsigned pp, unsigned planeBytes, const TmaParams& P) {
  const unsigned segBytes = DT ? tmaSegBytes(DT ? DT : 2) : P.segBytes;
  const unsigned pairsPerSeg = segBytes >> 4;
  unsigned seg = 0, chunk = pp;
  if (!DT || 8u * DT > 128u) {
    seg = pp / pairsPerSeg;
    chunk = pp - seg * pairsPerSeg;
  }
  return seg * 8u * planeBytes + mh * segBytes + ((chunk ^ tmaSwizzle<DT>(mh, P)) << 4);
}

// byte offset inside a buffer of sample (row m, phase p)
template <int DT>
__device__ __forceinline__ unsigned tmaSampleOffset(unsigned m, unsigned p, unsigned planeBytes, const TmaParams& P) {
  return (m & 7u) * planeBytes + tmaPairOffset<DT>(m >> 3, p >> 1, planeBytes, P) + (p & 1u) * 8u;
}

__device__ __forceinline__ unsigned smemU32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemU32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemU32(bar)) : "memory");
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
__device__ __forceinline__ void tmaLoad4(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                         int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smemU32(dst)),
      "l"(map), "r"(smemU32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- fused output stages ------------------------------------------------------------------------------------
enum : unsigned { kEpiNone = 0u, kEpiAmEnvelope = 1u, kEpiFmDemod = 2u };

// ref: src/am.cu:49 / src/quad_demod.cu:46-49 — 2 * saturate(|v|) - 1
__device__ __forceinline__ float amEnvelope(float2 v) { return __fadd_rn(scalbnf(__saturatef(hypotf(v.x, v.y)), 1), -1.0f); }

// gain * arg(next * conj(cur)) with the expression shape nvcc gives the reference's cuCmulf(next, cuConjf(cur)):
// FMUL, FMUL, FFMA, FFMA (ref: src/quad_demod.cu:30-31; same libdevice atan2f => same bits as the reference kernel)
__device__ __forceinline__ float quadFmValue(float2 cur, float2 next, float gain) {
  const float a = __fmul_rn(next.x, cur.y);
  const float b = __fmul_rn(next.y, cur.y);
  const float im = __fmaf_rn(next.y, cur.x, -a);
  const float re = __fmaf_rn(next.x, cur.x, b);
  return __fmul_rn(gain, atan2f(im, re));
}

// Complex outputs ob .. ob+7 of one thread -> global memory.
__device__ __forceinline__ void tmaStoreComplex(const TmaParams& P, unsigned chan, unsigned long long ob,
                                                const float2 (&acc)[8]) {
  float2* y = P.y + (size_t)chan * P.yStride;
  if (P.y16 && ob + 8 <= P.nOut) {
#pragma unroll
    for (int r = 0; r < 8; r += 2) {
      *reinterpret_cast<float4*>(y + ob + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      if (ob + r < P.nOut) y[ob + r] = acc[r];
    }
  }
}

// Real outputs (AM envelope / FM): values v[0..7] for output indices ob .. ob+7, `count` outputs exist in the channel.
__device__ __forceinline__ void tmaStoreReal(const TmaParams& P, unsigned chan, unsigned long long ob, const float (&v)[8],
                                             unsigned long long count) {
  float* y = reinterpret_cast<float*>(P.y) + (size_t)chan * P.yStride;  // yStride counts floats here
  if (P.y16 && ob + 8 <= count && (ob & 3ull) == 0) {
    *reinterpret_cast<float4*>(y + ob) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(y + ob + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      if (ob + r < count) y[ob + r] = v[r];
    }
  }
}

// Outputs of thread t of a tile whose first output is o0: complex, or the AM envelope of each.
__device__ __forceinline__ void tmaStoreTile(const TmaParams& P, unsigned chan, unsigned long long ob,
                                             const float2 (&acc)[8]) {
  if (P.epi == kEpiAmEnvelope) {
    float v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = amEnvelope(acc[r]);
    tmaStoreReal(P, chan, ob, v, P.nOut);
  } else {
    tmaStoreComplex(P, chan, ob, acc);
  }
}

enum TmaBlockKind : int { kBlkPrologue = 0, kBlkSteady = 1, kBlkTail = 2 };

// Elements the sample ring runs ahead of the FFMA2s (ring of 8 slots: 1..8).  Tuning knob, see profiles/r01_notes.md.
#ifndef GSDR_TMA_LOOKAHEAD
#define GSDR_TMA_LOOKAHEAD 6
#endif
constexpr int kTmaAhead = GSDR_TMA_LOOKAHEAD;
static_assert(kTmaAhead >= 1 && kTmaAhead <= 8, "the sample ring has 8 slots");

// One block of 8 window elements of a branch pair.  q[]: sample ring (slot = element index & 7).
// hnP/hnQ: taps block b ("new") of branches P/Q, hoP/hoQ: taps block b-1 ("old").  Element i feeds output r with
// tap (i-r) of the new block when r <= i, with tap (8+i-r) of the old block when r > i.  After element i its ring
// slot is free: the element six ahead is loaded (planes 6,7 of this block from a0, planes 0..5 of the next block
// from a1).  Old-tap slots i-1, i are dead after odd i and are refilled with the block after `new` (tapNext).
template <int KIND, bool REFILL_TAPS>
__device__ __forceinline__ void firPairBlock(
    float2 (&acc)[kTmaR], float4 (&q)[8], float (&hoP)[8], float (&hoQ)[8], const float (&hnP)[8],
    const float (&hnQ)[8], const unsigned char* a0, const unsigned char* a1, unsigned planeBytes,
    const float* tapNext) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (!(KIND == kBlkTail && i == 7)) {
      const float2 xP = make_float2(q[i].x, q[i].y);
      const float2 xQ = make_float2(q[i].z, q[i].w);
#pragma unroll
      for (int r = 0; r < kTmaR; r++) {
        const bool useNew = r <= i;
        if ((useNew && KIND != kBlkTail) || (!useNew && KIND != kBlkPrologue)) {
          const float h = useNew ? hnP[useNew ? i - r : 0] : hoP[useNew ? 0 : 8 + i - r];
          acc[r] = macTap(xP, h, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < kTmaR; r++) {
        const bool useNew = r <= i;
        if ((useNew && KIND != kBlkTail) || (!useNew && KIND != kBlkPrologue)) {
          const float h = useNew ? hnQ[useNew ? i - r : 0] : hoQ[useNew ? 0 : 8 + i - r];
          acc[r] = macTap(xQ, h, acc[r]);
        }
      }
    }
    // sample kTmaAhead (six) elements ahead -> the ring slot that element (i-2) vacated
    {
      const int e = i + kTmaAhead;
      const bool skip = (KIND == kBlkTail && e == 7);  // element 7 of the tail block is never used
      if (!skip) {
        const unsigned char* addr = (e < 8) ? a0 + (unsigned)e * planeBytes : a1 + (unsigned)(e - 8) * planeBytes;
        q[e & 7] = *reinterpret_cast<const float4*>(addr);
      }
    }
    if (REFILL_TAPS && (i & 1)) {
      const float4 v = *reinterpret_cast<const float4*>(tapNext + 2 * (i - 1));
      hoP[i - 1] = v.x;
      hoQ[i - 1] = v.y;
      hoP[i] = v.z;
      hoQ[i] = v.w;
    }
  }
}

// Slow stager for windows that reach past the TMA-visible rows (the last tile(s) of a channel): same layout,
// 8-byte cp.async with zero fill beyond the caller-guaranteed extent.
template <int NT, int DT>
__device__ __forceinline__ void tmaStageSlow(unsigned char* buf, const float2* src, unsigned long long in0,
                                             unsigned rows, unsigned planeBytes, const TmaParams& P,
                                             unsigned lt = threadIdx.x) {
  const unsigned D = DT ? (unsigned)DT : P.D;
  const unsigned total = rows * D;
  unsigned p = lt % D, m = lt / D;
  const unsigned dp = NT % D, dm = NT / D;
  for (unsigned s = lt; s < total; s += NT) {
    const unsigned long long g = in0 + s;
    const bool valid = g < P.nIn;
    cpAsync8z(buf + tmaSampleOffset<DT>(m, p, planeBytes, P), src + (valid ? g : 0ull), valid);
    p += dp;
    m += dm;
    if (p >= D) {
      p -= D;
      m += 1;
    }
  }
}

// Phasor of absolute sample n for the exact NCO: top 32 bits of (n * step mod 2^64) -> angle / pi in [-1, 1).
__device__ __forceinline__ float2 ncoExactPhasor(unsigned long long n, unsigned long long step) {
  const unsigned long long phase = n * step;
  const float a = (float)(int)(unsigned)(phase >> 32) * 4.656612873077392578125e-10f;
  float sn, cs;
  sincospif(a, &sn, &cs);
  return make_float2(cs, sn);
}
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}

// Barrier among the NT threads that build / mix a window: the whole CTA (id 0) or the producer warps only.
template <int NT, int BARRIER_ID>
__device__ __forceinline__ void mixBarrier() {
  if (BARRIER_ID == 0) {
    __syncthreads();
  } else {
    asm volatile("bar.sync %0, %1;" ::"n"(BARRIER_ID), "n"(NT) : "memory");
  }
}

// Row anchors A[m] = exp(j * phase(first sample of row m)) of a window, two-level: one sincospi per EIGHT rows,
//     A = coarse[a >> 3] * fine[a & 7],   a = absolute row number (row0 + m),
// coarse[g] the phasor of the first sample of absolute row 8g, fine[i] = exp(j * i * D * step) a per-CTA table (fine[0]
// is exactly 1).  A sincospi costs ~45 instructions, i.e. 5.6 per sample at D = 8 when taken per row; this way it is
// ~1.  Grouping by ABSOLUTE row number keeps the anchor a pure function of the absolute sample index, so time shards and
// stream blocks (which start at multiples of D from the capture's first sample) still reproduce the one-shot bits.
// Layout of the anchors: [m & 7][m >> 3], read by pass 2 with consecutive lanes on consecutive row groups.
// scratch: (rows / 8 + 2) float2.  Contains one barrier; the caller adds the one after it.
template <int NT, int BARRIER_ID>
__device__ __forceinline__ void ncoRowAnchors(float2* ncoA, float2* coarse, const float2* fine, unsigned rows,
                                              unsigned mhCount, unsigned long long row0, unsigned D, unsigned rho,
                                              unsigned long long step, unsigned lt) {
  const unsigned long long g0 = row0 >> 3;
  const unsigned groups = (unsigned)(((row0 + rows - 1u) >> 3) - g0) + 1u;
  for (unsigned i = lt; i < groups; i += NT) coarse[i] = ncoExactPhasor(((g0 + i) << 3) * D + rho, step);
  mixBarrier<NT, BARRIER_ID>();
  const unsigned r7 = (unsigned)(row0 & 7ull);
  for (unsigned m = lt; m < rows; m += NT) {
    const unsigned a = r7 + m;
    ncoA[(m & 7u) * mhCount + (m >> 3)] = cmulf(coarse[a >> 3], fine[a & 7u]);
  }
}

// shared-memory carve-up of the exact NCO's tables behind the row anchors: R[p], p < D (padded to an even count), the
// eight fine anchors, then the coarse-anchor scratch
__device__ __forceinline__ float2* ncoFineOf(float2* ncoR, unsigned D) { return ncoR + D + (D & 1u); }
__device__ __forceinline__ float2* ncoCoarseOf(float2* ncoR, unsigned D) { return ncoR + D + (D & 1u) + 8u; }
__host__ __device__ constexpr unsigned ncoTableFloat2s(unsigned D, unsigned rows) {
  return D + (D & 1u) + 8u + rows / 8u + 2u + ((rows / 8u) & 1u);  // even count: what follows stays 16-byte aligned
}

// In-place NCO mix of a landed window (all threads).  Element (m, p) is input sample in0 + m*D + p.
//
// Exact mode: phasor(row m, phase p) = A[m] * R[p] with A[m] = exp(j*phase(first sample of row m)) evaluated by
// sincospi from the exact 64-bit phase (one per row, pass 1) and R[p] = exp(j*p*step) a per-CTA table.  Two
// roundings per phasor, no recurrence chain (every element is independent: full ILP), and a pure function of the
// absolute sample index given the row alignment — which is identical for sharded and unsharded runs (shards and
// tiles start on multiples of D), so time shards reproduce the unsharded bits.  Work items are (branch pair,
// plane, row group) with consecutive lanes on consecutive row groups, like the FIR reader (conflict free).
// Literal mode (parity with the reference only): the reference's arithmetic per sample.
// Barrier among the NT threads that run a mix pass: the whole CTA (id 0) or the mixer warps only (named barrier).
// Pass 2 of the exact mix for compile-time decimations.  A thread keeps ONE residue s = mh & 7 of the row groups and
// one chunk of CH branch pairs for the whole window, so that
//   * its rotation-table entries R[p] live in registers for the whole window (no table loads);
//   * the swizzle XOR (a function of mh & 7 only) is a per-thread constant: the 16-byte slot of pair pp0 + j inside
//     a row is a per-thread register, and the only per-item address arithmetic is the row pointer;
//   * the eight lanes of a quarter warp (s = 0..7, same plane, same pair) read eight consecutive row groups, which
//     the swizzle spreads over all bank groups, exactly like the FIR reader.
// Per work item (CH pairs of one row): 1 LDS.64 (row anchor) + CH x (LDS.128, 4 complex multiplies, STS.128) and a
// handful of integer instructions; the arithmetic (w = A[m] * R[p], then x * w) and hence every result bit is that
// of the generic loop in tmaMixWindow.
template <int NT, int DT>
struct MixStatic {
  static constexpr unsigned PAIRS = (DT ? DT : 2) / 2;
  static constexpr unsigned CH = (PAIRS % 4 == 0) ? 4 : PAIRS;  // pairs per chunk
  static constexpr unsigned CHUNKS = PAIRS / CH;
  static constexpr unsigned G = NT / 8;  // threads per residue
  static constexpr bool ok = DT != 0 && NT % 8 == 0 && CH <= 7 && (G % CHUNKS == 0 || CHUNKS % G == 0);
};

template <int NT, int DT>
__device__ __forceinline__ void tmaMixRowsStatic(unsigned char* buf, unsigned mhCount, unsigned planeBytes,
                                                 const float2* ncoA, const float2* ncoR, const TmaParams& P,
                                                 unsigned lt, int dstDelta = 0) {  // dstDelta: mixed window goes to buf + dstDelta
  using M = MixStatic<NT, DT>;
  constexpr unsigned CH = M::CH, CHUNKS = M::CHUNKS, G = M::G;
  constexpr unsigned SEG = tmaSegBytes(DT ? DT : 2);
  constexpr unsigned PAIRS_PER_SEG = SEG / 16;
  constexpr unsigned CK_LANES = G < CHUNKS ? G : CHUNKS;
  constexpr unsigned V_LANES = G / CK_LANES;
  const unsigned s = lt & 7u;
  const unsigned rest = lt >> 3;
  const unsigned ckLane = rest % CK_LANES, vLane = rest / CK_LANES;
  const unsigned sw = tmaSwizzle<DT>(s, P);  // depends on mh & 7 only for every compile-time decimation
  const unsigned vTotal = 8u * ((mhCount + 7u) >> 3);
  for (unsigned ck = ckLane; ck < CHUNKS; ck += CK_LANES) {
    const unsigned pp0 = ck * CH;
    const unsigned seg = pp0 / PAIRS_PER_SEG, c0 = pp0 % PAIRS_PER_SEG;
    float2 r[2 * CH];
    unsigned off[CH];  // byte offset of pair pp0 + j inside a row segment of this thread's rows
#pragma unroll
    for (unsigned j = 0; j < CH; j++) {
      const float4 rr = *reinterpret_cast<const float4*>(ncoR + 2u * (pp0 + j));
      r[2 * j] = make_float2(rr.x, rr.y);
      r[2 * j + 1] = make_float2(rr.z, rr.w);
      off[j] = ((c0 + j) ^ sw) << 4;
    }
    unsigned char* base = buf + seg * 8u * planeBytes + s * SEG;
    const float2* anchors = ncoA + s;
    for (unsigned v = vLane; v < vTotal; v += V_LANES) {
      const unsigned ml = v & 7u, u = v >> 3;
      if (s + 8u * u >= mhCount) continue;
      const float2 an = anchors[ml * mhCount + 8u * u];
      unsigned char* row = base + ml * planeBytes + u * (8u * SEG);
      float4 x[CH];
#pragma unroll
      for (unsigned j = 0; j < CH; j++) x[j] = *reinterpret_cast<const float4*>(row + off[j]);
#pragma unroll
      for (unsigned j = 0; j < CH; j++) {
        const float2 w0 = cmulf(an, r[2 * j]);
        const float2 w1 = cmulf(an, r[2 * j + 1]);
        const float2 a = cmulf(make_float2(x[j].x, x[j].y), w0);
        const float2 c = cmulf(make_float2(x[j].z, x[j].w), w1);
        *reinterpret_cast<float4*>(row + off[j] + dstDelta) = make_float4(a.x, a.y, c.x, c.y);
      }
    }
  }
}

template <int MODE, int NT, int DT, int BARRIER_ID = 0>
__device__ __forceinline__ void tmaMixWindow(unsigned char* buf, unsigned long long in0, unsigned long long row0,
                                             unsigned rows, unsigned planeBytes, float2* ncoA, float2* ncoR,
                                             const TmaParams& P, unsigned lt = threadIdx.x) {  // row0: in0 / D
  const unsigned D = DT ? (unsigned)DT : P.D;
  const unsigned pairsPerRow = D >> 1;
  if (MODE == kPolyNcoExact) {
    const unsigned mhCount = rows >> 3;  // rows is a multiple of 8
    // pass 1: row anchors, stored [ml][mh] so that pass 2 reads them with consecutive lanes
    ncoRowAnchors<NT, BARRIER_ID>(ncoA, ncoCoarseOf(ncoR, D), ncoFineOf(ncoR, D), rows, mhCount, P.ncoRow0 + row0, D,
                                  P.ncoRho, P.ncoStep, lt);
    mixBarrier<NT, BARRIER_ID>();
    if constexpr (MixStatic<NT, DT>::ok) {
      tmaMixRowsStatic<NT, DT>(buf, mhCount, planeBytes, ncoA, ncoR, P, lt);
      return;
    }
    // pass 2: work item = (chunk of up to 4 branch pairs, plane, row group); the row anchor is loaded once per item
    const unsigned pairsPerChunk = (pairsPerRow % 4u == 0u) ? 4u : 1u;  // finer items balance better for odd counts
    const unsigned chunks = pairsPerRow / pairsPerChunk;
    const unsigned total = chunks * rows;
    unsigned g = lt / mhCount;  // g = chunk * 8 + ml
    unsigned mh = lt - g * mhCount;
    const unsigned dg = NT / mhCount, dmh = NT - dg * mhCount;
    for (unsigned e = lt; e < total; e += NT) {
      const unsigned ck = g >> 3, ml = g & 7u;
      const float2 an = ncoA[ml * mhCount + mh];
      unsigned char* rowBase = buf + ml * planeBytes;
      const unsigned pp0 = ck * pairsPerChunk;
#pragma unroll 4
      for (unsigned k = 0; k < pairsPerChunk; k++) {
        const unsigned pp = pp0 + k;
        const float4 rr = *reinterpret_cast<const float4*>(ncoR + 2u * pp);
        float4* q = reinterpret_cast<float4*>(rowBase + tmaPairOffset<DT>(mh, pp, planeBytes, P));
        const float4 v = *q;
        const float2 w0 = cmulf(an, make_float2(rr.x, rr.y));
        const float2 w1 = cmulf(an, make_float2(rr.z, rr.w));
        const float2 a = cmulf(make_float2(v.x, v.y), w0);
        const float2 c = cmulf(make_float2(v.z, v.w), w1);
        *q = make_float4(a.x, a.y, c.x, c.y);
      }
      g += dg;
      mh += dmh;
      if (mh >= mhCount) {
        mh -= mhCount;
        g += 1;
      }
    }
  } else {
    const unsigned total = rows * pairsPerRow;
    for (unsigned e = lt; e < total; e += NT) {
      const unsigned m = e / pairsPerRow;
      const unsigned pp = e - m * pairsPerRow;
      float4* q = reinterpret_cast<float4*>(buf + tmaSampleOffset<DT>(m, 2 * pp, planeBytes, P));
      const float4 v = *q;
      float2 w[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const unsigned long long s = in0 + (unsigned long long)m * D + 2 * pp + k;
        const unsigned idx = P.ncoFirst32 + (unsigned)s;  // ref: src/adjustFrequency.cu:23,35-50; src/fm.cu:43-47
        const float period = __frcp_rn(P.ncoF);
        const float tt = __fdiv_rn(fmodf(__uint2float_rn(idx), P.ncoFs), P.ncoFs);
        const float u = fmodf(tt, period);
        sincospif(u * 2.0f, &w[k].y, &w[k].x);
      }
      const float2 a = cmulf(make_float2(v.x, v.y), w[0]);
      const float2 c = cmulf(make_float2(v.z, v.w), w[1]);
      *q = make_float4(a.x, a.y, c.x, c.y);
    }
  }
}

// The FIR proper for one thread: outputs 8t..8t+7 of the tile in `buf`, branch pairs [ppBegin, ppStop).
template <int DT>
__device__ __forceinline__ void firComputePairs(float2 (&acc)[kTmaR], const unsigned char* buf, const float* hs,
                                                unsigned t, unsigned ppBegin, unsigned ppStop, unsigned Jpad,
                                                unsigned planeBytes, const TmaParams& P) {
  const unsigned nbk = Jpad >> 3;  // even, >= 2
  // address of (plane 0, row group t + c, branch pair pp)
  auto blockAddr = [&](unsigned pp, unsigned c) -> const unsigned char* {
    return buf + tmaPairOffset<DT>(t + c, pp, planeBytes, P);
  };
  const float* hp = hs + (size_t)ppBegin * 2u * Jpad;
  float hAP[8], hAQ[8], hBP[8], hBQ[8];
  float4 q[8];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const float4 v = *reinterpret_cast<const float4*>(hp + 2 * k);
    hAP[k] = v.x, hAQ[k] = v.y, hAP[k + 1] = v.z, hAQ[k + 1] = v.w;
    const float4 w = *reinterpret_cast<const float4*>(hp + 16 + 2 * k);
    hBP[k] = w.x, hBQ[k] = w.y, hBP[k + 1] = w.z, hBQ[k + 1] = w.w;
  }
  const unsigned char* a0 = blockAddr(ppBegin, 0);
#pragma unroll
  for (int e = 0; e < kTmaAhead; e++) q[e] = *reinterpret_cast<const float4*>(a0 + (unsigned)e * planeBytes);
#pragma unroll 1  // one copy of the ~600-instruction body: a fully unrolled pair loop overflows the instruction cache
  for (unsigned pp = ppBegin; pp < ppStop; pp++) {
    // block 0: prologue (new = A).  Its old set B already holds tap block 1 (initial load / previous tail).
    const unsigned char* a1 = blockAddr(pp, 1);
    firPairBlock<kBlkPrologue, false>(acc, q, hBP, hBQ, hAP, hAQ, a0, a1, planeBytes, hp);
    const float* tapNext = hp + 32;  // tap block 2
    unsigned c = 1;
    for (; c + 2 < nbk; c += 2) {
      a0 = a1, a1 = blockAddr(pp, c + 1);
      firPairBlock<kBlkSteady, true>(acc, q, hAP, hAQ, hBP, hBQ, a0, a1, planeBytes, tapNext);
      a0 = a1, a1 = blockAddr(pp, c + 2);
      firPairBlock<kBlkSteady, true>(acc, q, hBP, hBQ, hAP, hAQ, a0, a1, planeBytes, tapNext + 16);
      tapNext += 32;
    }
    // last steady block (odd c = nbk-1): old = A, new = B; refills A with tap block nbk = next pair's block 0
    a0 = a1, a1 = blockAddr(pp, nbk);
    firPairBlock<kBlkSteady, true>(acc, q, hAP, hAQ, hBP, hBQ, a0, a1, planeBytes, tapNext);
    // tail block (even): old = B; its sample look-ahead and tap refills already belong to the next pair
    a0 = a1;
    a1 = (pp + 1 < ppStop) ? blockAddr(pp + 1, 0) : a0;
    firPairBlock<kBlkTail, true>(acc, q, hBP, hBQ, hAP, hAQ, a0, a1, planeBytes, tapNext + 16);
    a0 = a1;
    hp += 2u * Jpad;
  }
}

// Same, for at most 8 taps per branch (Jpad == 8, one tap block): each pair is a prologue block followed at once by
// the tail block, which reads block 0 (A) as its old set and refills it with the next pair's block 0.
template <int DT>
__device__ __forceinline__ void firComputePairsShort(float2 (&acc)[kTmaR], const unsigned char* buf, const float* hs,
                                                     unsigned t, unsigned ppBegin, unsigned ppStop,
                                                     unsigned planeBytes, const TmaParams& P) {
  auto blockAddr = [&](unsigned pp, unsigned c) -> const unsigned char* {
    return buf + tmaPairOffset<DT>(t + c, pp, planeBytes, P);
  };
  const float* hp = hs + (size_t)ppBegin * 16u;
  float hAP[8], hAQ[8], hBP[8], hBQ[8];
  float4 q[8];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const float4 v = *reinterpret_cast<const float4*>(hp + 2 * k);
    hAP[k] = v.x, hAQ[k] = v.y, hAP[k + 1] = v.z, hAQ[k + 1] = v.w;
    hBP[k] = 0.0f, hBQ[k] = 0.0f, hBP[k + 1] = 0.0f, hBQ[k + 1] = 0.0f;  // never read: the prologue uses A only
  }
  const unsigned char* a0 = blockAddr(ppBegin, 0);
#pragma unroll
  for (int e = 0; e < kTmaAhead; e++) q[e] = *reinterpret_cast<const float4*>(a0 + (unsigned)e * planeBytes);
#pragma unroll 1
  for (unsigned pp = ppBegin; pp < ppStop; pp++) {
    const unsigned char* a1 = blockAddr(pp, 1);
    firPairBlock<kBlkPrologue, false>(acc, q, hBP, hBQ, hAP, hAQ, a0, a1, planeBytes, hp);
    a0 = a1;
    a1 = (pp + 1 < ppStop) ? blockAddr(pp + 1, 0) : a0;
    firPairBlock<kBlkTail, true>(acc, q, hAP, hAQ, hBP, hBQ, a0, a1, planeBytes, hp + 16);
    a0 = a1;
    hp += 16;
  }
}

// MODE: kPolyFC / kPolyNcoExact / kPolyNcoLiteral.  TG threads own 8 outputs each (tile = 8*TG outputs); PSPLIT
// thread groups split the branch pairs.  DT: compile-time decimation (row size, swizzle and plane pitch become
// immediates; needs Jpad <= kTmaJpadCap) or 0 for run-time geometry.  Two window buffers: tile k+1 is in flight
// while tile k is filtered.
template <int MODE, int TG, int PSPLIT, int DT, int NBUF, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT, MINB)
    firTmaKernel(const __grid_constant__ CUtensorMap map, const TmaParams P) {
  static_assert(NBUF == 1 || NBUF == 2, "one or two window buffers");
  constexpr unsigned NT = TG * PSPLIT;
  constexpr unsigned BOUT = kTmaR * TG;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long fullBar[2];
  // fused output stages exist in the NCO modes only (gsdrAmDemod / gsdrFmDemodFused): the plain FIR keeps its schedule
  constexpr bool kHasEpilogue = MODE != kPolyFC;
  __shared__ float2 nextFirst[kHasEpilogue ? TG : 1];  // FM: every thread's first output, for its predecessor
  const unsigned D = DT ? (unsigned)DT : P.D;
  const unsigned rowBytes = 8u * D;
  const unsigned segBytes = DT ? tmaSegBytes(DT ? DT : 2) : P.segBytes;
  const unsigned numSegs = rowBytes / segBytes;
  const unsigned planeBytes = DT ? tmaPlaneRows(TG, kTmaJpadCap, DT ? DT : 2) * segBytes : P.planeBytes;
  const unsigned bufBytes = numSegs * 8u * planeBytes;
  // the swizzle patterns are functions of absolute shared-memory address bits: align the buffers to 1024 bytes
  unsigned char* bufBase = smemRaw + ((1024u - (smemU32(smemRaw) & 1023u)) & 1023u);  // NBUF x bufBytes
  float4* scratch = reinterpret_cast<float4*>(bufBase + NBUF * bufBytes);  // 2 x (PSPLIT-1) x TG x 64 B of partial sums
  float* hs = reinterpret_cast<float*>(scratch + 2u * (PSPLIT - 1) * (kTmaR / 2) * TG);  // [D/2][Jpad][2] (+32 zeros)
  // exact NCO only: per-row anchors of the current window and the per-CTA table exp(j*p*step), p < D
  float2* ncoA = reinterpret_cast<float2*>(hs + (size_t)(DT ? DT : P.D) * P.Jpad + 32u);
  float2* ncoR = ncoA + (BOUT + P.Jpad);

  const unsigned tid = threadIdx.x;
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  const unsigned numPairs = D >> 1;
  const unsigned ppBegin = (grp * numPairs) / PSPLIT;
  const unsigned ppEnd = ((grp + 1) * numPairs) / PSPLIT;
  const unsigned rowsStaged = BOUT + P.Jpad;  // rows a tile needs (8 per output block + the taps' reach)
  const unsigned tileOut = kHasEpilogue ? P.tileOut : BOUT;  // FM tiles overlap by one row group

  if (tid == 0) {
    mbarInit(&fullBar[0], 1);
    mbarInit(&fullBar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == kPolyNcoExact) {
    for (unsigned p = tid; p < D; p += NT) ncoR[p] = ncoExactPhasor((unsigned long long)p, P.ncoStep);
    for (unsigned p = tid; p < 8u; p += NT) ncoFineOf(ncoR, D)[p] = ncoExactPhasor((unsigned long long)p * D, P.ncoStep);
  }
  __syncthreads();
#ifndef GSDR_NO_PDL
  // Programmatic dependent launch: the next launch's CTAs may be scheduled as this grid drains (they then wait
  // here), and this grid touches no global memory before the previous work in the stream has completed — so
  // stream order is unchanged for the caller.  Back-to-back launches: 167.0 -> 165.3 us on config 2.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif

  // (channel, tile) of this CTA's current and next work item; the grid stride is pre-split on the host so that no
  // division is needed per tile.
  unsigned chan = blockIdx.x / P.tilesPerChannel;
  unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  auto advance = [&](unsigned& c, unsigned& tl) {
    c += P.strideChan;
    tl += P.strideTile;
    if (tl >= P.tilesPerChannel) {
      tl -= P.tilesPerChannel;
      c += 1;
    }
  };
  // A tile is TMA-fed when every row it stages is visible to the tensor map.
  auto tileIsFast = [&](unsigned tl) -> bool { return tl * tileOut + rowsStaged <= P.tmaRows; };
  auto issueTile = [&](unsigned c, unsigned tl, unsigned b) {
    if (GSDR_DBG(P) & 1u) return;
    unsigned char* buf = bufBase + b * bufBytes;
    if (tileIsFast(tl)) {
      if (tid == 0) {
        // generic-proxy accesses to this buffer (mix pass, partial-sum scratch) are ordered before the async-proxy
        // writes of the tensor copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&fullBar[b], bufBytes);
        for (unsigned sg = 0; sg < numSegs; sg++) {
          tmaLoad4(buf + sg * 8u * planeBytes, &map, &fullBar[b], (int)(sg * (segBytes / 4u)), (int)(tl * (tileOut / 8)), 0,
                   (int)c);
        }
      }
    } else {
      tmaStageSlow<NT, DT>(buf, P.x + (size_t)c * P.xStride, (unsigned long long)tl * tileOut * D, rowsStaged, planeBytes, P);
    }
  };

  unsigned phaseBits = 0;  // parity of each buffer's mbarrier
  if (NBUF == 2) {
    if (chan < P.numChannels) issueTile(chan, tile, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  unsigned tapsChan = 0xffffffffu;

  for (unsigned it = 0; chan < P.numChannels; it++) {
    const unsigned b = (NBUF == 2) ? (it & 1u) : 0u;
    unsigned char* buf = bufBase + b * bufBytes;
    const unsigned long long o0 = (unsigned long long)tile * tileOut;
    unsigned nextChan = chan, nextTile = tile;
    advance(nextChan, nextTile);
    if (NBUF == 2) {
      if (nextChan < P.numChannels) issueTile(nextChan, nextTile, b ^ 1u);
    } else {
      issueTile(chan, tile, 0);  // single buffer: other resident CTAs compute while this copy is in flight
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");

    // taps -> hs2[pp][j] = (h[j*D + 2pp], h[j*D + 2pp + 1]); once per CTA unless the channel's tap set changes
    const bool tapsReloaded = chan != tapsChan && (tapsChan == 0xffffffffu || P.hStride != 0);
    if (tapsReloaded) {
      const float* h = P.h + (size_t)chan * P.hStride;
      const unsigned nh = D * P.Jpad;
      for (unsigned i = tid; i < nh + 32u; i += NT) {
        const unsigned pp = i / (2u * P.Jpad);
        const unsigned rem = i - pp * 2u * P.Jpad;
        const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
        hs[i] = (i < nh && ti < P.T) ? __ldg(h + ti) : 0.0f;
      }
    }
    tapsChan = chan;

    // Data-ready: every thread waits on the tile's mbarrier itself (TMA tiles), so no CTA barrier is needed; the
    // rare cp.async tiles and tap reloads are written by other threads and do need one.
    const bool fast = tileIsFast(tile);
    if (!(GSDR_DBG(P) & 1u) && fast) {
      mbarWait(&fullBar[b], (phaseBits >> b) & 1u);
      phaseBits ^= 1u << b;
    }
    if (!fast || tapsReloaded) {
      if (NBUF == 2) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");  // this tile's slow-path copies have landed
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      }
      __syncthreads();
    }
    if (MODE == kPolyNcoExact || MODE == kPolyNcoLiteral) {
      tmaMixWindow<MODE, NT, DT>(buf, o0 * D, o0, rowsStaged, planeBytes, ncoA, ncoR, P);
      __syncthreads();
    }

    float2 acc[kTmaR];
#pragma unroll
    for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
    const unsigned ppStop = (GSDR_DBG(P) & 2u) ? ppBegin : ppEnd;
    if (ppBegin < ppStop) firComputePairs<DT>(acc, buf, hs, t, ppBegin, ppStop, P.Jpad, planeBytes, P);

    // Partial sums of the branch-pair groups go through a small double-buffered scratch area, so ONE barrier per
    // tile both publishes them and tells thread 0 that this window may be overwritten by the tile after next.
    // With two groups the collecting group alternates from tile to tile: the collector's extra work (partial-sum
    // reads, stores) then loads the two warps' schedulers evenly (a + b == b + a: same bits either way).
    const unsigned collector = (PSPLIT == 2) ? (it & 1u) : 0u;
    if (PSPLIT > 1 && grp != collector) {
      float4* red = scratch + (size_t)(it & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
      const unsigned slot = (PSPLIT == 2) ? 0u : grp - 1u;
#pragma unroll
      for (int k = 0; k < kTmaR / 2; k++) {
        red[(slot * (kTmaR / 2) + k) * TG + t] =
            make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
      }
    }
    __syncthreads();
    if (grp == collector) {
      if (PSPLIT > 1) {
        const float4* red = scratch + (size_t)(it & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
#pragma unroll
        for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
          for (int k = 0; k < kTmaR / 2; k++) {
            const float4 v = red[((g - 1) * (kTmaR / 2) + k) * TG + t];
            acc[2 * k].x += v.x;
            acc[2 * k].y += v.y;
            acc[2 * k + 1].x += v.z;
            acc[2 * k + 1].y += v.w;
          }
        }
      }
      const unsigned long long ob = o0 + (unsigned long long)t * kTmaR;
      if (GSDR_DBG(P) & 4u) {
        if (acc[0].x == 123.456f) P.y[(size_t)chan * P.yStride] = acc[1];  // measurement hook: no output traffic
      } else if (kHasEpilogue && P.epi == kEpiFmDemod) {
        // d[n] = gain * arg(y[n+1] * conj(y[n])) (ref: src/quad_demod.cu:23-37, src/fm.cu:58-68): the successor of this
        // thread's last output is the next thread's first one; the tile's last thread only supplies it (its own
        // outputs belong to the next tile, which starts 8 outputs before this one ends)
        nextFirst[t] = acc[0];
        asm volatile("bar.sync 3, %0;" ::"n"(TG) : "memory");
        const float2 nx = (t + 1 < (unsigned)TG) ? nextFirst[t + 1] : make_float2(0.0f, 0.0f);
        float v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = quadFmValue(acc[r], r < 7 ? acc[r + 1] : nx, P.epiGain);
        if (t + 1 < (unsigned)TG) tmaStoreReal(P, chan, ob, v, P.nOut - 1);  // nOut low-pass values, nOut - 1 phase steps
      } else if (kHasEpilogue) {
        tmaStoreTile(P, chan, ob, acc);
      } else {
        tmaStoreComplex(P, chan, ob, acc);
      }
    }
    chan = nextChan;
    tile = nextTile;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Warp-specialised variant for the fused NCO (exact phase law): the CTA's first TG*PSPLIT threads only filter,
// MIXW extra warps only copy and mix.  Tile k+1 is fetched (TMA) and mixed in place by the mixer warps while the
// FIR warps filter tile k; everything is handed over with mbarriers, so the mix pass (latency bound: sincospi,
// table look-ups, read-modify-write of shared memory) fills the issue slots the FIR leaves, instead of
// alternating with it.
//   fullRaw[b]  TMA bytes of buffer b have landed            (tx count, armed by mixer thread 0)
//   fullMix[b]  buffer b is mixed                             (every mixer thread arrives)
//   empty[b]    the FIR warps are done reading buffer b       (every FIR thread arrives)
// One tap set for all channels only (hStride == 0); the host falls back to firTmaKernel otherwise.
// ---------------------------------------------------------------------------------------------------------
template <int TG, int PSPLIT, int DT, int MIXW, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT + 32 * MIXW, MINB)
    firTmaNcoSpecKernel(const __grid_constant__ CUtensorMap map, const TmaParams P) {
  constexpr unsigned NTF = TG * PSPLIT;  // filter threads
  constexpr unsigned NTM = 32 * MIXW;    // mixer threads
  constexpr unsigned BOUT = kTmaR * TG;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long fullRaw[2], fullMix[2], emptyBar[2];
  const unsigned D = DT ? (unsigned)DT : P.D;
  const unsigned rowBytes = 8u * D;
  const unsigned segBytes = DT ? tmaSegBytes(DT ? DT : 2) : P.segBytes;
  const unsigned numSegs = rowBytes / segBytes;
  const unsigned planeBytes = DT ? tmaPlaneRows(TG, kTmaJpadCap, DT ? DT : 2) * segBytes : P.planeBytes;
  const unsigned bufBytes = numSegs * 8u * planeBytes;
  unsigned char* bufBase = smemRaw + ((1024u - (smemU32(smemRaw) & 1023u)) & 1023u);
  float4* scratch = reinterpret_cast<float4*>(bufBase + 2u * bufBytes);
  float* hs = reinterpret_cast<float*>(scratch + 2u * (PSPLIT - 1) * (kTmaR / 2) * TG);
  float2* ncoA = reinterpret_cast<float2*>(hs + (size_t)D * P.Jpad + 32u);
  float2* ncoR = ncoA + (BOUT + P.Jpad);

  const unsigned tid = threadIdx.x;
  const unsigned rowsStaged = BOUT + P.Jpad;
  if (tid == 0) {
    for (int b = 0; b < 2; b++) {
      mbarInit(&fullRaw[b], 1);
      mbarInit(&fullMix[b], NTM);
      mbarInit(&emptyBar[b], NTF);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (unsigned p = tid; p < D; p += NTF + NTM) ncoR[p] = ncoExactPhasor((unsigned long long)p, P.ncoStep);
  for (unsigned p = tid; p < 8u; p += NTF + NTM) ncoFineOf(ncoR, D)[p] = ncoExactPhasor((unsigned long long)p * D, P.ncoStep);
  {
    const unsigned nh = D * P.Jpad;
    for (unsigned i = tid; i < nh + 32u; i += NTF + NTM) {
      const unsigned pp = i / (2u * P.Jpad);
      const unsigned rem = i - pp * 2u * P.Jpad;
      const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
      hs[i] = (i < nh && ti < P.T) ? __ldg(P.h + ti) : 0.0f;
    }
  }
  __syncthreads();

  unsigned chan = blockIdx.x / P.tilesPerChannel;
  unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  auto advance = [&](unsigned& c, unsigned& tl) {
    c += P.strideChan;
    tl += P.strideTile;
    if (tl >= P.tilesPerChannel) {
      tl -= P.tilesPerChannel;
      c += 1;
    }
  };
  auto tileIsFast = [&](unsigned tl) -> bool { return tl * BOUT + rowsStaged <= P.tmaRows; };

  if (tid >= NTF) {
    // ===================== mixer warps: copy + mix, up to two tiles ahead of the filter warps =====================
    const unsigned lt = tid - NTF;
    for (unsigned k = 0; chan < P.numChannels; k++, advance(chan, tile)) {
      const unsigned b = k & 1u;
      unsigned char* buf = bufBase + b * bufBytes;
      if (k >= 2) mbarWait(&emptyBar[b], ((k >> 1) - 1u) & 1u);  // the filter warps have left tile k-2
      const bool fast = tileIsFast(tile);
      if (fast) {
        if (lt == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbarExpectTx(&fullRaw[b], bufBytes);
          for (unsigned sg = 0; sg < numSegs; sg++) {
            tmaLoad4(buf + sg * 8u * planeBytes, &map, &fullRaw[b], (int)(sg * (segBytes / 4u)),
                     (int)(tile * (BOUT / 8)), 0, (int)chan);
          }
        }
        mbarWait(&fullRaw[b], (k >> 1) & 1u);
      } else {
        if (lt == 0) mbarArrive(&fullRaw[b]);  // keep the phase of the unused barrier in step with k
        tmaStageSlow<NTM, DT>(buf, P.x + (size_t)chan * P.xStride, (unsigned long long)tile * BOUT * D, rowsStaged,
                              planeBytes, P, lt);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        mixBarrier<NTM, 1>();
      }
      tmaMixWindow<kPolyNcoExact, NTM, DT, 1>(buf, (unsigned long long)tile * BOUT * D, (unsigned long long)tile * BOUT,
                                              rowsStaged, planeBytes, ncoA, ncoR, P, lt);
      mbarArrive(&fullMix[b]);   // release: the mixed window is visible to whoever acquires the barrier
      mixBarrier<NTM, 1>();      // ncoA may be overwritten by the next tile's anchors
    }
    return;
  }

  // ============================== filter warps ==============================
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  const unsigned numPairs = D >> 1;
  const unsigned ppBegin = (grp * numPairs) / PSPLIT;
  const unsigned ppEnd = ((grp + 1) * numPairs) / PSPLIT;
  for (unsigned k = 0; chan < P.numChannels; k++, advance(chan, tile)) {
    const unsigned b = k & 1u;
    const unsigned char* buf = bufBase + b * bufBytes;
    const unsigned long long o0 = (unsigned long long)tile * BOUT;
    mbarWait(&fullMix[b], (k >> 1) & 1u);
    float2 acc[kTmaR];
#pragma unroll
    for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
    if (ppBegin < ppEnd) firComputePairs<DT>(acc, buf, hs, t, ppBegin, ppEnd, P.Jpad, planeBytes, P);
    mbarArrive(&emptyBar[b]);  // this thread has no more reads of the window
    if (PSPLIT > 1) {
      float4* red = scratch + (size_t)(k & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
      if (grp > 0) {
#pragma unroll
        for (int q = 0; q < kTmaR / 2; q++) {
          red[((grp - 1) * (kTmaR / 2) + q) * TG + t] =
              make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");
      if (grp == 0) {
#pragma unroll
        for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
          for (int q = 0; q < kTmaR / 2; q++) {
            const float4 v = red[((g - 1) * (kTmaR / 2) + q) * TG + t];
            acc[2 * q].x += v.x;
            acc[2 * q].y += v.y;
            acc[2 * q + 1].x += v.z;
            acc[2 * q + 1].y += v.w;
          }
        }
      }
    }
    if (grp == 0) {
      tmaStoreTile(P, chan, o0 + (unsigned long long)t * kTmaR, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Real input x real taps (gsdrFirFF; replaces ref: src/fir.cu:26-71 for <float, float, float>) on the same
// inner loop.  Two consecutive outputs are one FFMA2 lane pair:
//     (out[2n], out[2n+1]) = sum_i (x[2n*D + i], x[2n*D + i + D]) * h[i]
// i.e. a complex-input FIR with decimation 2D over the stream c[k] = (x[k], x[k + D]), whose float2 outputs ARE the
// real outputs in order.  c is never in HBM: producer warps bulk-copy the raw float window of a tile into shared
// memory (cp.async.bulk, one instruction per tile, completion on an mbarrier), then rearrange it into the plane
// layout the FIR loop reads (row m' of c = floats [2D*m', 2D*m' + 3D) of the window).  Filter warps run
// firComputePairs unchanged.  Hand-over as in the fused-NCO kernel:
//   rawBar[r]   bytes of raw buffer r have landed                 (tx count, armed by producer thread 0)
//   fullMix[b]  window b is laid out                              (every producer thread arrives)
//   empty[b]    the filter warps are done reading window b        (every filter thread arrives)
// Plane pitch is 16 bytes more than a multiple of 128 so that the eight planes a quarter-warp of producers writes
// start in eight different bank groups.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulkLoad1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemU32(dst)),
               "l"(src), "r"(bytes), "r"(smemU32(bar))
               : "memory");
}

constexpr unsigned kRealPlanePad = 16;

// NPLANE = 2: complex taps (gsdrFirCF; ref: src/fir.cu:26-71 for <float, cuComplex, cuComplex>).  As in
// firTmaCcKernel the branch groups split into a real and an imaginary tap plane over the same window (PSPLIT even);
// group 0 collects the real plane into (re[2n], re[2n+1]) and the imaginary plane into (im[2n], im[2n+1]) and stores
// the two complex outputs (re[2n], im[2n], re[2n+1], im[2n+1]) with one 16-byte store.
template <int TG, int PSPLIT, int DT, int MIXW, int NWIN, int NRAW, int MINB, int NPLANE = 1>
__global__ void __launch_bounds__(TG* PSPLIT + 32 * MIXW, MINB) firTmaRealKernel(const RealParams P) {
  static_assert(NWIN == 1 || NWIN == 2, "one or two window buffers");
  static_assert(NPLANE == 1 || (NPLANE == 2 && PSPLIT % 2 == 0), "two tap planes need an even number of groups");
  static_assert(NRAW >= 2 && NRAW <= 4, "raw ring of 2..4 buffers (NRAW-1 bulk copies in flight per CTA)");
  constexpr unsigned NTF = TG * PSPLIT;  // filter threads
  constexpr unsigned NTM = 32 * MIXW;    // producer threads
  constexpr unsigned BOUT = kTmaR * TG;  // output PAIRS per tile
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long rawBar[NRAW], fullMix[2], emptyBar[2];
  const unsigned D = DT ? (unsigned)DT : P.D;  // pair-stream decimation = 2 x the caller's
  const unsigned Dr = D >> 1;
  const unsigned rowBytes = 8u * D;
  const unsigned segBytes = DT ? tmaSegBytes(DT ? DT : 2) : P.segBytes;
  const unsigned numSegs = rowBytes / segBytes;
  const unsigned planeBytes =
      DT ? tmaPlaneRows(TG, kTmaJpadCap, DT ? DT : 2) * segBytes + kRealPlanePad : P.planeBytes;
  const unsigned bufBytes = numSegs * 8u * planeBytes;
  unsigned char* bufBase = smemRaw;
  float4* scratch = reinterpret_cast<float4*>(bufBase + NWIN * bufBytes);
  float* hs = reinterpret_cast<float*>(scratch + 2u * (PSPLIT - 1) * (kTmaR / 2) * TG);
  const unsigned tapPlaneFloats = D * P.Jpad + 32u;
  float* raw = hs + (size_t)NPLANE * tapPlaneFloats;  // NRAW x rawFloats

  const unsigned tid = threadIdx.x;
  const unsigned rowsStaged = BOUT + P.Jpad;
  if (tid == 0) {
    for (int b = 0; b < NRAW; b++) mbarInit(&rawBar[b], 1);
    for (int b = 0; b < 2; b++) {
      mbarInit(&fullMix[b], NTM);
      mbarInit(&emptyBar[b], NTF);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned chan = blockIdx.x / P.tilesPerChannel;
  unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  // taps -> hs2[pp][j] = (h[j*D + 2pp], h[j*D + 2pp + 1]) by threads [0, nthreads)
  auto loadTaps = [&](unsigned c, unsigned nthreads) {
    const float* h = P.h + (size_t)NPLANE * c * P.hStride;  // NPLANE = 2: interleaved (re, im)
    const unsigned nh = D * P.Jpad;
    for (unsigned i = tid; i < NPLANE * tapPlaneFloats; i += nthreads) {
      const unsigned pl = (NPLANE == 2 && i >= tapPlaneFloats) ? 1u : 0u;
      const unsigned q = i - pl * tapPlaneFloats;
      const unsigned pp = q / (2u * P.Jpad);
      const unsigned rem = q - pp * 2u * P.Jpad;
      const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
      hs[i] = (q < nh && ti < P.T) ? __ldg(h + NPLANE * ti + pl) : 0.0f;
    }
  };
  if (chan < P.numChannels) loadTaps(chan, NTF + NTM);
  __syncthreads();
  auto advance = [&](unsigned& c, unsigned& tl) {
    c += P.strideChan;
    tl += P.strideTile;
    if (tl >= P.tilesPerChannel) {
      tl -= P.tilesPerChannel;
      c += 1;
    }
  };

  if (tid >= NTF) {
    // ===================== producer warps: raw copy + rearrangement, ahead of the filter warps =====================
    const unsigned lt = tid - NTF;
    const float* xr = reinterpret_cast<const float*>(P.x);
    const unsigned rawBytes = P.rawFloats * 4u;
    // a tile is bulk-copied when every float it reads lies inside the caller-guaranteed extent
    auto tileIsFast = [&](unsigned tl) -> bool {
      return (unsigned long long)tl * BOUT * D + P.rawFloats <= P.nIn;
    };
    auto issueRaw = [&](unsigned c, unsigned tl, unsigned rb) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbarExpectTx(&rawBar[rb], rawBytes);
      bulkLoad1d(raw + (size_t)rb * P.rawFloats, xr + (size_t)c * P.xStride + (size_t)tl * BOUT * D, rawBytes,
                 &rawBar[rb]);
    };
    // prefetch cursor: the bulk copy of tile k + NRAW - 1 is issued while tile k is rearranged
    unsigned pfChan = chan, pfTile = tile;
    for (unsigned j = 0; j + 1 < (unsigned)NRAW; j++) {
      if (lt == 0 && pfChan < P.numChannels && tileIsFast(pfTile)) issueRaw(pfChan, pfTile, j);
      if (pfChan < P.numChannels) advance(pfChan, pfTile);
    }
    for (unsigned k = 0; chan < P.numChannels; k++) {
      const unsigned rb = k % NRAW;
      const unsigned b = (NWIN == 2) ? (k & 1u) : 0u;
      // ring slot (k - 1) % NRAW was last read one iteration ago (tile k-1); the barrier that ends every iteration
      // orders those reads before this copy
      if (lt == 0 && pfChan < P.numChannels && tileIsFast(pfTile)) issueRaw(pfChan, pfTile, (k + NRAW - 1) % NRAW);
      if (pfChan < P.numChannels) advance(pfChan, pfTile);
      if (k >= (unsigned)NWIN) mbarWait(&emptyBar[b], ((k / NWIN) - 1u) & 1u);  // filters have left this window
      float* rawb = raw + (size_t)rb * P.rawFloats;
      if (tileIsFast(tile)) {
        mbarWait(&rawBar[rb], (k / NRAW) & 1u);
      } else {
        if (lt == 0) mbarArrive(&rawBar[rb]);  // keep the phase of the unused barrier in step with k
        const float* xc = xr + (size_t)chan * P.xStride;
        const unsigned long long g0 = (unsigned long long)tile * BOUT * D;
        for (unsigned i = lt; i < P.rawFloats; i += NTM) {
          const unsigned long long g = g0 + i;
          rawb[i] = (g < P.nIn) ? __ldg(xc + g) : 0.0f;
        }
        mixBarrier<NTM, 1>();
      }
      // rearrange: one work item = one row m' of the pair stream: floats [D*m', D*m' + D + Dr) of the raw window
      // -> Dr float4 (branch pairs) of plane m' & 7, row group m' >> 3
      unsigned char* buf = bufBase + b * bufBytes;
      if (DT == 2) {
        // decimation 1: rows are (x0, x1, x1, x2) at a stride of two floats; four rows per work item from two
        // 16-byte loads.  Rows 4q..4q+3 share a row group and sit in planes 4(q & 1) .. 4(q & 1) + 3.
        for (unsigned q = lt; q < (rowsStaged >> 2); q += NTM) {
          const float4 a = *reinterpret_cast<const float4*>(rawb + 8u * q);
          const float4 c = *reinterpret_cast<const float4*>(rawb + 8u * q + 4u);
          const float e = rawb[8u * q + 8u];
          unsigned char* dst = buf + ((q & 1u) * 4u) * planeBytes + (q >> 1) * 16u;
          *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, a.y, a.z);
          *reinterpret_cast<float4*>(dst + planeBytes) = make_float4(a.z, a.w, a.w, c.x);
          *reinterpret_cast<float4*>(dst + 2u * planeBytes) = make_float4(c.x, c.y, c.y, c.z);
          *reinterpret_cast<float4*>(dst + 3u * planeBytes) = make_float4(c.z, c.w, c.w, e);
        }
      } else
      for (unsigned m = lt; m < rowsStaged; m += NTM) {
        const float* r = rawb + (size_t)m * D;
        unsigned char* rowBase = buf + (m & 7u) * planeBytes;
        const unsigned mh = m >> 3;
        if (DT != 0) {
          // compile-time row: all loads first (8-byte aligned pairs), then the stores
          constexpr unsigned NF = (DT ? DT : 2) + (DT ? DT : 2) / 2;  // floats read
          float v[NF + 1];
#pragma unroll
          for (unsigned i = 0; i < NF; i += 2) {
            if (i + 1 < NF) {
              const float2 t2 = *reinterpret_cast<const float2*>(r + i);
              v[i] = t2.x;
              v[i + 1] = t2.y;
            } else if (i < NF) {
              v[i] = r[i];
            }
          }
#pragma unroll
          for (unsigned pp = 0; pp < (DT ? DT : 2) / 2; pp++) {
            *reinterpret_cast<float4*>(rowBase + tmaPairOffset<DT>(mh, pp, planeBytes, P)) =
                make_float4(v[2 * pp], v[2 * pp + Dr], v[2 * pp + 1], v[2 * pp + 1 + Dr]);
          }
        } else {
          for (unsigned pp = 0; pp < Dr; pp++) {
            *reinterpret_cast<float4*>(rowBase + tmaPairOffset<DT>(mh, pp, planeBytes, P)) =
                make_float4(r[2 * pp], r[2 * pp + Dr], r[2 * pp + 1], r[2 * pp + 1 + Dr]);
          }
        }
      }
      mbarArrive(&fullMix[b]);  // release: the window is visible to whoever acquires the barrier
      mixBarrier<NTM, 1>();     // ring slot rb may be overwritten by the copy issued in the next iteration
      advance(chan, tile);
    }
    return;
  }

  // ============================== filter warps ==============================
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  const unsigned numPairs = D >> 1;
  constexpr unsigned NSUB = PSPLIT / NPLANE;  // groups per tap plane: they split the branch pairs
  const unsigned plane = (NPLANE == 2) ? (grp & 1u) : 0u;
  const unsigned sub = (NPLANE == 2) ? (grp >> 1) : grp;
  const unsigned ppBegin = (sub * numPairs) / NSUB;
  const unsigned ppEnd = ((sub + 1) * numPairs) / NSUB;
  const float* hsPlane = hs + plane * tapPlaneFloats;
  const unsigned long long pairsWhole = P.nOutReal >> 1;  // output pairs with both halves inside the output
  unsigned tapsChan = chan;
  for (unsigned k = 0; chan < P.numChannels; k++, advance(chan, tile)) {
    if (P.hStride != 0 && chan != tapsChan) {
      // per-channel tap sets: every filter thread has left the previous channel's taps before they are replaced
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");
      loadTaps(chan, NTF);
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");
      tapsChan = chan;
    }
    const unsigned b = (NWIN == 2) ? (k & 1u) : 0u;
    const unsigned char* buf = bufBase + b * bufBytes;
    const unsigned long long o0 = (unsigned long long)tile * BOUT;
    mbarWait(&fullMix[b], (k / NWIN) & 1u);
    float2 acc[kTmaR];
#pragma unroll
    for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
    if (ppBegin < ppEnd) {
      if (P.Jpad == 8u) {
        firComputePairsShort<DT>(acc, buf, hsPlane, t, ppBegin, ppEnd, planeBytes, P);
      } else {
        firComputePairs<DT>(acc, buf, hsPlane, t, ppBegin, ppEnd, P.Jpad, planeBytes, P);
      }
    }
    mbarArrive(&emptyBar[b]);  // this thread has no more reads of the window
    if (PSPLIT > 1) {
      float4* red = scratch + (size_t)(k & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
      if (grp > 0) {
#pragma unroll
        for (int q = 0; q < kTmaR / 2; q++) {
          red[((grp - 1) * (kTmaR / 2) + q) * TG + t] =
              make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");
      if (grp == 0) {
#pragma unroll
        for (int g = 1; g < PSPLIT; g++) {
          if (NPLANE == 2 && (g & 1)) continue;  // imaginary-plane groups are collected below
#pragma unroll
          for (int q = 0; q < kTmaR / 2; q++) {
            const float4 v = red[((g - 1) * (kTmaR / 2) + q) * TG + t];
            acc[2 * q].x += v.x;
            acc[2 * q].y += v.y;
            acc[2 * q + 1].x += v.z;
            acc[2 * q + 1].y += v.w;
          }
        }
      }
    }
    if (NPLANE == 2) {
      if (grp == 0) {
        // imaginary plane: (im[2n], im[2n+1]) per output pair, summed over that plane's groups in a fixed order
        const float4* red = scratch + (size_t)(k & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
        float2 im[kTmaR];
#pragma unroll
        for (int r = 0; r < kTmaR; r++) im[r] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int g = 1; g < PSPLIT; g += 2) {
#pragma unroll
          for (int q = 0; q < kTmaR / 2; q++) {
            const float4 v = red[((g - 1) * (kTmaR / 2) + q) * TG + t];
            im[2 * q].x += v.x;
            im[2 * q].y += v.y;
            im[2 * q + 1].x += v.z;
            im[2 * q + 1].y += v.w;
          }
        }
        const unsigned long long ob = o0 + (unsigned long long)t * kTmaR;  // first output pair of this thread
        float2* y = P.y + (size_t)chan * P.yStride;
        if (P.y16 && ob + kTmaR <= pairsWhole) {
#pragma unroll
          for (int r = 0; r < kTmaR; r++) {
            *reinterpret_cast<float4*>(y + 2 * (ob + r)) = make_float4(acc[r].x, im[r].x, acc[r].y, im[r].y);
          }
        } else {
#pragma unroll
          for (int r = 0; r < kTmaR; r++) {
            const unsigned long long o = 2 * (ob + r);
            if (o < P.nOutReal) y[o] = make_float2(acc[r].x, im[r].x);
            if (o + 1 < P.nOutReal) y[o + 1] = make_float2(acc[r].y, im[r].y);
          }
        }
      }
      continue;
    }
    if (grp == 0) {
      const unsigned long long ob = o0 + (unsigned long long)t * kTmaR;  // first output pair of this thread
      float* y = reinterpret_cast<float*>(P.y) + (size_t)chan * P.yStride;
      if (P.y16 && ob + kTmaR <= pairsWhole) {
#pragma unroll
        for (int r = 0; r < kTmaR; r += 2) {
          *reinterpret_cast<float4*>(y + 2 * (ob + r)) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
        }
      } else {
#pragma unroll
        for (int r = 0; r < kTmaR; r++) {
          const unsigned long long o = 2 * (ob + r);
          if (o < P.nOutReal) y[o] = acc[r].x;
          if (o + 1 < P.nOutReal) y[o + 1] = acc[r].y;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Complex taps x complex input (gsdrFirCC; replaces ref: src/fir.cu:26-71 for <cuComplex, cuComplex, cuComplex>)
// as two real-tap filters over the same window:  out = sum x*hr + j * sum x*hi.  The branch groups split into two
// tap planes (even groups: real parts, odd groups: imaginary parts); a group of the imaginary plane publishes its
// partial sum already multiplied by j, (-im, re), so the fixed-order sum of the partial sums is the result.
// Same TMA staging, layout and inner loop as firTmaKernel; the taps take twice the shared memory.
// PSPLIT is even; the PSPLIT / 2 groups of a plane split the branch pairs.
// ---------------------------------------------------------------------------------------------------------
template <int TG, int PSPLIT, int DT, int NBUF, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT, MINB)
    firTmaCcKernel(const __grid_constant__ CUtensorMap map, const TmaParams P) {
  static_assert(NBUF == 1 || NBUF == 2, "one or two window buffers");
  static_assert(PSPLIT >= 2 && PSPLIT % 2 == 0, "two tap planes");
  constexpr unsigned NT = TG * PSPLIT;
  constexpr unsigned BOUT = kTmaR * TG;
  constexpr unsigned NSUB = PSPLIT / 2;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long fullBar[2];
  const unsigned D = DT ? (unsigned)DT : P.D;
  const unsigned rowBytes = 8u * D;
  const unsigned segBytes = DT ? tmaSegBytes(DT ? DT : 2) : P.segBytes;
  const unsigned numSegs = rowBytes / segBytes;
  const unsigned planeBytes = DT ? tmaPlaneRows(TG, kTmaJpadCap, DT ? DT : 2) * segBytes : P.planeBytes;
  const unsigned bufBytes = numSegs * 8u * planeBytes;
  unsigned char* bufBase = smemRaw + ((1024u - (smemU32(smemRaw) & 1023u)) & 1023u);
  float4* scratch = reinterpret_cast<float4*>(bufBase + NBUF * bufBytes);
  float* hs = reinterpret_cast<float*>(scratch + 2u * (PSPLIT - 1) * (kTmaR / 2) * TG);  // 2 x ([D/2][Jpad][2] + 32)
  const unsigned tapPlaneFloats = D * P.Jpad + 32u;

  const unsigned tid = threadIdx.x;
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  const unsigned plane = grp & 1u, sub = grp >> 1;
  const unsigned numPairs = D >> 1;
  const unsigned ppBegin = (sub * numPairs) / NSUB;
  const unsigned ppEnd = ((sub + 1) * numPairs) / NSUB;
  const unsigned rowsStaged = BOUT + P.Jpad;

  if (tid == 0) {
    mbarInit(&fullBar[0], 1);
    mbarInit(&fullBar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned chan = blockIdx.x / P.tilesPerChannel;
  unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  auto advance = [&](unsigned& c, unsigned& tl) {
    c += P.strideChan;
    tl += P.strideTile;
    if (tl >= P.tilesPerChannel) {
      tl -= P.tilesPerChannel;
      c += 1;
    }
  };
  auto tileIsFast = [&](unsigned tl) -> bool { return tl * BOUT + rowsStaged <= P.tmaRows; };
  auto issueTile = [&](unsigned c, unsigned tl, unsigned b) {
    unsigned char* buf = bufBase + b * bufBytes;
    if (tileIsFast(tl)) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&fullBar[b], bufBytes);
        for (unsigned sg = 0; sg < numSegs; sg++) {
          tmaLoad4(buf + sg * 8u * planeBytes, &map, &fullBar[b], (int)(sg * (segBytes / 4u)), (int)(tl * (BOUT / 8)), 0,
                   (int)c);
        }
      }
    } else {
      tmaStageSlow<NT, DT>(buf, P.x + (size_t)c * P.xStride, (unsigned long long)tl * BOUT * D, rowsStaged, planeBytes, P);
    }
  };

  unsigned phaseBits = 0;
  if (NBUF == 2) {
    if (chan < P.numChannels) issueTile(chan, tile, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  unsigned tapsChan = 0xffffffffu;

  for (unsigned it = 0; chan < P.numChannels; it++) {
    const unsigned b = (NBUF == 2) ? (it & 1u) : 0u;
    const unsigned char* buf = bufBase + b * bufBytes;
    const unsigned long long o0 = (unsigned long long)tile * BOUT;
    unsigned nextChan = chan, nextTile = tile;
    advance(nextChan, nextTile);
    if (NBUF == 2) {
      if (nextChan < P.numChannels) issueTile(nextChan, nextTile, b ^ 1u);
    } else {
      issueTile(chan, tile, 0);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");

    // taps: plane 0 = real parts, plane 1 = imaginary parts, each laid out like firTmaKernel's hs2
    const bool tapsReloaded = chan != tapsChan && (tapsChan == 0xffffffffu || P.hStride != 0);
    if (tapsReloaded) {
      const float* h = P.h + 2u * (size_t)chan * P.hStride;  // interleaved (re, im)
      const unsigned nh = D * P.Jpad;
      for (unsigned i = tid; i < 2u * tapPlaneFloats; i += NT) {
        const unsigned pl = i >= tapPlaneFloats ? 1u : 0u;
        const unsigned k = i - pl * tapPlaneFloats;
        const unsigned pp = k / (2u * P.Jpad);
        const unsigned rem = k - pp * 2u * P.Jpad;
        const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
        hs[i] = (k < nh && ti < P.T) ? __ldg(h + 2u * ti + pl) : 0.0f;
      }
    }
    tapsChan = chan;

    const bool fast = tileIsFast(tile);
    if (fast) {
      mbarWait(&fullBar[b], (phaseBits >> b) & 1u);
      phaseBits ^= 1u << b;
    }
    if (!fast || tapsReloaded) {
      if (NBUF == 2) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      }
      __syncthreads();
    }

    float2 acc[kTmaR];
#pragma unroll
    for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
    if (ppBegin < ppEnd) {
      firComputePairs<DT>(acc, buf, hs + plane * tapPlaneFloats, t, ppBegin, ppEnd, P.Jpad, planeBytes, P);
    }
    if (grp > 0) {
      float4* red = scratch + (size_t)(it & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
#pragma unroll
      for (int k = 0; k < kTmaR / 2; k++) {
        red[((grp - 1) * (kTmaR / 2) + k) * TG + t] =
            plane ? make_float4(-acc[2 * k].y, acc[2 * k].x, -acc[2 * k + 1].y, acc[2 * k + 1].x)
                  : make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
      }
    }
    __syncthreads();
    if (grp == 0) {
      const float4* red = scratch + (size_t)(it & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
#pragma unroll
      for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
        for (int k = 0; k < kTmaR / 2; k++) {
          const float4 v = red[((g - 1) * (kTmaR / 2) + k) * TG + t];
          acc[2 * k].x += v.x;
          acc[2 * k].y += v.y;
          acc[2 * k + 1].x += v.z;
          acc[2 * k + 1].y += v.w;
        }
      }
      const unsigned long long ob = o0 + (unsigned long long)t * kTmaR;
      float2* y = P.y + (size_t)chan * P.yStride;
      if (P.y16 && ob + kTmaR <= P.nOut) {
#pragma unroll
        for (int r = 0; r < kTmaR; r += 2) {
          *reinterpret_cast<float4*>(y + ob + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
        }
      } else {
#pragma unroll
        for (int r = 0; r < kTmaR; r++) {
          if (ob + r < P.nOut) y[ob + r] = acc[r];
        }
      }
    }
    chan = nextChan;
    tile = nextTile;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// int8 IQ input (the conversion of ref: src/conversion.cu:20-27 fused into the staging):
//     x[k] = (conv(I[k]), conv(Q[k])),  conv(v) = max(-1, v / 127)        (interleaved int8 I, Q in HBM: 2 bytes / sample)
//     out[n] = sum_i x[nD + i] * [phasor(first + nD + i)] * h[i]
// Producer warps bulk-copy the raw bytes of a tile's window (a quarter of the float traffic), convert them, optionally
// apply the exact NCO (anchor per row x per-CTA rotation table, as in tmaMixWindow) and write the plane layout; filter
// warps run firComputePairs.  conv(v) = max(v, -127) / 127 exactly, and the division by 127 is folded into the taps
// (staged as h[i] / 127): one rounding per tap instead of one per sample — within the FP32 tolerance of
// convert-then-filter, not bit-identical to it.  Hand-over as in firTmaRealKernel (rawBar / fullMix / empty).
// ---------------------------------------------------------------------------------------------------------
struct Int8Params : TmaParams {
  unsigned rawBytes;  // bytes of input one tile reads, a multiple of 16
  float tapScale;     // 1 / 127
};

// the four int8 of w (already XORed with 0x80808080: offset binary) -> floats, clamped to >= -127
__device__ __forceinline__ float4 int8x4ToFloat(unsigned wBiased) {
  float4 r;
  // byte k into the mantissa of 2^23: 0x4B000000 | u  ==  8388608 + u;  u - 128 = the signed value
  r.x = __uint_as_float(__byte_perm(wBiased, 0x4B000000u, 0x7650)) - 8388736.0f;
  r.y = __uint_as_float(__byte_perm(wBiased, 0x4B000000u, 0x7651)) - 8388736.0f;
  r.z = __uint_as_float(__byte_perm(wBiased, 0x4B000000u, 0x7652)) - 8388736.0f;
  r.w = __uint_as_float(__byte_perm(wBiased, 0x4B000000u, 0x7653)) - 8388736.0f;
  r.x = fmaxf(r.x, -127.0f);
  r.y = fmaxf(r.y, -127.0f);
  r.z = fmaxf(r.z, -127.0f);
  r.w = fmaxf(r.w, -127.0f);
  return r;
}

template <int TG, int PSPLIT, int DT, int MIXW, bool NCO, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT + 32 * MIXW, MINB) firTmaInt8Kernel(const Int8Params P) {
  constexpr unsigned NTF = TG * PSPLIT;  // filter threads
  constexpr unsigned NTM = 32 * MIXW;    // producer threads
  constexpr unsigned BOUT = kTmaR * TG;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long rawBar[2], fullMix[2], emptyBar[2];
  const unsigned D = DT ? (unsigned)DT : P.D;
  const unsigned rowBytes = 8u * D;
  const unsigned segBytes = DT ? tmaSegBytes(DT ? DT : 2) : P.segBytes;
  const unsigned numSegs = rowBytes / segBytes;
  const unsigned planeBytes =
      DT ? tmaPlaneRows(TG, kTmaJpadCap, DT ? DT : 2) * segBytes + kRealPlanePad : P.planeBytes;
  const unsigned bufBytes = numSegs * 8u * planeBytes;
  unsigned char* bufBase = smemRaw;
  float4* scratch = reinterpret_cast<float4*>(bufBase + 2u * bufBytes);
  float* hs = reinterpret_cast<float*>(scratch + 2u * (PSPLIT - 1) * (kTmaR / 2) * TG);
  float2* ncoA = reinterpret_cast<float2*>(hs + (size_t)D * P.Jpad + 32u);  // row anchors of the window being built
  float2* ncoR = ncoA + (BOUT + P.Jpad);                                     // exp(j*p*step), p < D
  unsigned char* raw = reinterpret_cast<unsigned char*>(ncoR + D + (D & 1u));  // 2 x rawBytes, 16-byte aligned

  const unsigned tid = threadIdx.x;
  const unsigned rowsStaged = BOUT + P.Jpad;
  if (tid == 0) {
    for (int b = 0; b < 2; b++) {
      mbarInit(&rawBar[b], 1);
      mbarInit(&fullMix[b], NTM);
      mbarInit(&emptyBar[b], NTF);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (NCO) {
    for (unsigned p = tid; p < D; p += NTF + NTM) ncoR[p] = ncoExactPhasor((unsigned long long)p, P.ncoStep);
  }
  {
    const unsigned nh = D * P.Jpad;
    for (unsigned i = tid; i < nh + 32u; i += NTF + NTM) {
      const unsigned pp = i / (2u * P.Jpad);
      const unsigned rem = i - pp * 2u * P.Jpad;
      const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
      hs[i] = (i < nh && ti < P.T) ? __ldg(P.h + ti) * P.tapScale : 0.0f;
    }
  }
  __syncthreads();

  unsigned chan = blockIdx.x / P.tilesPerChannel;
  unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  auto advance = [&](unsigned& c, unsigned& tl) {
    c += P.strideChan;
    tl += P.strideTile;
    if (tl >= P.tilesPerChannel) {
      tl -= P.tilesPerChannel;
      c += 1;
    }
  };

  if (tid >= NTF) {
    // ===================== producer warps =====================
    const unsigned lt = tid - NTF;
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(P.x);  // 2 bytes per sample
    auto tileIsFast = [&](unsigned tl) -> bool {
      return (unsigned long long)tl * BOUT * D + (P.rawBytes >> 1) <= P.nIn;
    };
    auto issueRaw = [&](unsigned c, unsigned tl, unsigned rb) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbarExpectTx(&rawBar[rb], P.rawBytes);
      bulkLoad1d(raw + (size_t)rb * P.rawBytes, xb + 2u * ((size_t)c * P.xStride + (size_t)tl * BOUT * D), P.rawBytes,
                 &rawBar[rb]);
    };
    if (lt == 0 && chan < P.numChannels && tileIsFast(tile)) issueRaw(chan, tile, 0);
    for (unsigned k = 0; chan < P.numChannels; k++) {
      unsigned nextChan = chan, nextTile = tile;
      advance(nextChan, nextTile);
      const unsigned b = k & 1u;
      if (lt == 0 && nextChan < P.numChannels && tileIsFast(nextTile)) issueRaw(nextChan, nextTile, b ^ 1u);
      if (k >= 2) mbarWait(&emptyBar[b], ((k >> 1) - 1u) & 1u);
      unsigned char* rawb = raw + (size_t)b * P.rawBytes;
      const unsigned long long in0 = (unsigned long long)tile * BOUT * D;
      if (tileIsFast(tile)) {
        mbarWait(&rawBar[b], (k >> 1) & 1u);
      } else {
        if (lt == 0) mbarArrive(&rawBar[b]);
        const unsigned short* xc = reinterpret_cast<const unsigned short*>(xb) + (size_t)chan * P.xStride;
        unsigned short* r16 = reinterpret_cast<unsigned short*>(rawb);
        for (unsigned i = lt; i < (P.rawBytes >> 1); i += NTM) {
          const unsigned long long g = in0 + i;
          r16[i] = (g < P.nIn) ? __ldg(xc + g) : (unsigned short)0;
        }
        mixBarrier<NTM, 1>();
      }
      unsigned char* buf = bufBase + b * bufBytes;
      if (NCO) {
        const unsigned mhCount = rowsStaged >> 3;
        for (unsigned m = lt; m < rowsStaged; m += NTM) {
          ncoA[(m & 7u) * mhCount + (m >> 3)] = ncoExactPhasor(P.ncoFirst + in0 + (unsigned long long)m * D, P.ncoStep);
        }
        mixBarrier<NTM, 1>();
      }
      // work item = (row m, branch pair pp): 4 raw bytes -> two converted (and mixed) samples -> one float4
      {
        const unsigned pairsPerRow = D >> 1;
        const unsigned total = rowsStaged * pairsPerRow;
        unsigned m = lt / pairsPerRow, pp = lt - m * pairsPerRow;
        const unsigned dm = NTM / pairsPerRow, dpp = NTM - dm * pairsPerRow;
        const unsigned mhCount = rowsStaged >> 3;
        for (unsigned e = lt; e < total; e += NTM) {
          const unsigned w = *reinterpret_cast<const unsigned*>(rawb + 4u * e) ^ 0x80808080u;
          float4 v = int8x4ToFloat(w);
          if (NCO) {
            const float2 an = ncoA[(m & 7u) * mhCount + (m >> 3)];
            const float4 rr = *reinterpret_cast<const float4*>(ncoR + 2u * pp);
            const float2 w0 = cmulf(an, make_float2(rr.x, rr.y));
            const float2 w1 = cmulf(an, make_float2(rr.z, rr.w));
            const float2 a = cmulf(make_float2(v.x, v.y), w0);
            const float2 c = cmulf(make_float2(v.z, v.w), w1);
            v = make_float4(a.x, a.y, c.x, c.y);
          }
          *reinterpret_cast<float4*>(buf + tmaSampleOffset<DT>(m, 2u * pp, planeBytes, P)) = v;
          m += dm;
          pp += dpp;
          if (pp >= pairsPerRow) {
            pp -= pairsPerRow;
            m += 1;
          }
        }
      }
      mbarArrive(&fullMix[b]);
      mixBarrier<NTM, 1>();  // raw buffer b and the anchors may be overwritten
      chan = nextChan;
      tile = nextTile;
    }
    return;
  }

  // ============================== filter warps ==============================
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  const unsigned numPairs = D >> 1;
  const unsigned ppBegin = (grp * numPairs) / PSPLIT;
  const unsigned ppEnd = ((grp + 1) * numPairs) / PSPLIT;
  for (unsigned k = 0; chan < P.numChannels; k++, advance(chan, tile)) {
    const unsigned b = k & 1u;
    const unsigned char* buf = bufBase + b * bufBytes;
    const unsigned long long o0 = (unsigned long long)tile * BOUT;
    mbarWait(&fullMix[b], (k >> 1) & 1u);
    float2 acc[kTmaR];
#pragma unroll
    for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
    if (ppBegin < ppEnd) firComputePairs<DT>(acc, buf, hs, t, ppBegin, ppEnd, P.Jpad, planeBytes, P);
    mbarArrive(&emptyBar[b]);
    if (PSPLIT > 1) {
      float4* red = scratch + (size_t)(k & 1u) * (PSPLIT - 1) * (kTmaR / 2) * TG;
      if (grp > 0) {
#pragma unroll
        for (int q = 0; q < kTmaR / 2; q++) {
          red[((grp - 1) * (kTmaR / 2) + q) * TG + t] =
              make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");
      if (grp == 0) {
#pragma unroll
        for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
          for (int q = 0; q < kTmaR / 2; q++) {
            const float4 v = red[((g - 1) * (kTmaR / 2) + q) * TG + t];
            acc[2 * q].x += v.x;
            acc[2 * q].y += v.y;
            acc[2 * q + 1].x += v.z;
            acc[2 * q + 1].y += v.w;
          }
        }
      }
    }
    if (grp == 0) {
      const unsigned long long ob = o0 + (unsigned long long)t * kTmaR;
      float2* y = P.y + (size_t)chan * P.yStride;
      if (P.y16 && ob + kTmaR <= P.nOut) {
#pragma unroll
        for (int r = 0; r < kTmaR; r += 2) {
          *reinterpret_cast<float4*>(y + ob + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
        }
      } else {
#pragma unroll
        for (int r = 0; r < kTmaR; r++) {
          if (ob + r < P.nOut) y[ob + r] = acc[r];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// One input, K frequency shifts (the idea of the reference's dead k_Fm4x, ref: src/fm.cu:71-179):
//     out_k[n] = sum_i x[nD + i] * exp(j * phase_k(first + nD + i)) * h[i],   k < K <= 16
// The window of a tile is fetched from HBM ONCE (tensor copy into a raw buffer) and mixed K times, each time into the
// work buffer the FIR then reads, so the wideband input crosses HBM once instead of K times.  Per shift the arithmetic
// is exactly gsdrAdjustFrequencyFirFC's (row anchor x per-shift rotation table, firComputePairs), hence the outputs are
// bit-identical to K separate calls through firTmaKernel with the same tile shape.
// What it buys is measured in profiles/r02/channelizer.jsonl: little where the FIR is issue-bound (32 taps per input
// sample), up to the HBM ratio for short filters.
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned kChanMaxShifts = 16;
struct ChanParams : TmaParams {
  unsigned numShifts;
  unsigned long long yShiftStride;  // elements between the output arrays of consecutive shifts
  unsigned long long steps[kChanMaxShifts];
};

template <int TG, int PSPLIT, int DT, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT, MINB)
    firTmaChannelizerKernel(const __grid_constant__ CUtensorMap map, const ChanParams P) {
  static_assert(DT != 0 && DT <= 16, "compile-time decimation, rows of one segment");
  static_assert(MixStatic<TG * PSPLIT, DT>::ok, "needs the register-resident mixer");
  constexpr unsigned NT = TG * PSPLIT;
  constexpr unsigned BOUT = kTmaR * TG;
  constexpr unsigned D = DT;
  constexpr unsigned planeBytes = tmaPlaneRows(TG, kTmaJpadCap, DT) * tmaSegBytes(DT);
  constexpr unsigned bufBytes = 8u * planeBytes;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long fullBar;
  unsigned char* raw = smemRaw + ((1024u - (smemU32(smemRaw) & 1023u)) & 1023u);
  unsigned char* work = raw + bufBytes;
  float4* scratch = reinterpret_cast<float4*>(work + bufBytes);
  float* hs = reinterpret_cast<float*>(scratch + (PSPLIT - 1) * (kTmaR / 2) * TG);
  float2* ncoA = reinterpret_cast<float2*>(hs + (size_t)D * P.Jpad + 32u);
  float2* ncoR = ncoA + (BOUT + P.Jpad);  // [numShifts][D]

  const unsigned tid = threadIdx.x;
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  constexpr unsigned numPairs = D >> 1;
  const unsigned ppBegin = (grp * numPairs) / PSPLIT;
  const unsigned ppEnd = ((grp + 1) * numPairs) / PSPLIT;
  const unsigned rowsStaged = BOUT + P.Jpad;
  const unsigned mhCount = rowsStaged >> 3;

  if (tid == 0) {
    mbarInit(&fullBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (unsigned i = tid; i < P.numShifts * D; i += NT) {
    ncoR[i] = ncoExactPhasor((unsigned long long)(i % D), P.steps[i / D]);
  }
  {
    const unsigned nh = D * P.Jpad;
    for (unsigned i = tid; i < nh + 32u; i += NT) {
      const unsigned pp = i / (2u * P.Jpad);
      const unsigned rem = i - pp * 2u * P.Jpad;
      const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
      hs[i] = (i < nh && ti < P.T) ? __ldg(P.h + ti) : 0.0f;
    }
  }
  __syncthreads();

  unsigned phase = 0;
  for (unsigned tile = blockIdx.x; tile < P.tilesPerChannel; tile += gridDim.x) {
    const unsigned long long o0 = (unsigned long long)tile * BOUT;
    const unsigned long long in0 = o0 * D;
    if (tile * BOUT + rowsStaged <= P.tmaRows) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&fullBar, bufBytes);
        tmaLoad4(raw, &map, &fullBar, 0, (int)(tile * (BOUT / 8)), 0, 0);
      }
      mbarWait(&fullBar, phase);
      phase ^= 1u;
    } else {
      tmaStageSlow<NT, DT>(raw, P.x, in0, rowsStaged, planeBytes, P);
      cpAsyncCommitWaitAll();
      __syncthreads();
    }
    for (unsigned k = 0; k < P.numShifts; k++) {
      const unsigned long long step = P.steps[k];
      for (unsigned m = tid; m < rowsStaged; m += NT) {
        ncoA[(m & 7u) * mhCount + (m >> 3)] = ncoExactPhasor(P.ncoFirst + in0 + (unsigned long long)m * D, step);
      }
      __syncthreads();  // anchors written; every thread has left the work buffer (FIR of the previous shift)
      tmaMixRowsStatic<NT, DT>(raw, mhCount, planeBytes, ncoA, ncoR + k * D, P, tid, (int)bufBytes);
      __syncthreads();
      float2 acc[kTmaR];
#pragma unroll
      for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
      if (ppBegin < ppEnd) firComputePairs<DT>(acc, work, hs, t, ppBegin, ppEnd, P.Jpad, planeBytes, P);
      if (PSPLIT > 1) {
        if (grp > 0) {
#pragma unroll
          for (int q = 0; q < kTmaR / 2; q++) {
            scratch[((grp - 1) * (kTmaR / 2) + q) * TG + t] =
                make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
          }
        }
        __syncthreads();
        if (grp == 0) {
#pragma unroll
          for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
            for (int q = 0; q < kTmaR / 2; q++) {
              const float4 w = scratch[((g - 1) * (kTmaR / 2) + q) * TG + t];
              acc[2 * q].x += w.x;
              acc[2 * q].y += w.y;
              acc[2 * q + 1].x += w.z;
              acc[2 * q + 1].y += w.w;
            }
          }
        }
      }
      if (grp == 0) {
        const unsigned long long ob = o0 + (unsigned long long)t * kTmaR;
        float2* y = P.y + (size_t)k * P.yShiftStride;
        if (P.y16 && ob + kTmaR <= P.nOut) {
#pragma unroll
          for (int r = 0; r < kTmaR; r += 2) {
            *reinterpret_cast<float4*>(y + ob + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
          }
        } else {
#pragma unroll
          for (int r = 0; r < kTmaR; r++) {
            if (ob + r < P.nOut) y[ob + r] = acc[r];
          }
        }
      }
    }
    __syncthreads();  // the last shift's mix has read the raw window: the next tile's copy may overwrite it
  }
}

// ---------------------------------------------------------------------------------------------------------
// Wide rows (D a multiple of 16 above 16: rows of more than 128 bytes, BASELINE config 3), with or without the
// exact NCO.  firTmaKernel / firTmaNcoSpecKernel hold a whole window (all 128-byte row segments) per buffer: at
// D = 32 that is 74 KB per 256-output tile, so a double-buffered CTA is alone on its SM with FOUR filter warps — one
// per scheduler, nobody to hide its latencies (FIR-only 0.757 ms against 0.645 ms with eight warps per SM;
// profiles/r02/sweep_fc_d32.jsonl) — and the single-buffered variant with two CTAs per SM cannot overlap the copy
// (0.76 ms).  Here the pipeline stage is ONE SEGMENT of a tile: a buffer holds the 8 branch pairs of one 128-byte
// segment for 512 outputs (8 planes x 72 row groups x 128 B = 72 KB), two buffers alternate, and the EIGHT filter
// warps (64 threads x 4 branch groups) keep their accumulators in registers across the segments of a tile; partial
// sums are exchanged and outputs stored after the last one.  Halo rows drop from 12.5 % to 6 % of a tile, the next
// segment is always in flight (TMA, one tensor copy per stage), and with the NCO the mixer warps work one stage
// ahead exactly as in firTmaNcoSpecKernel (row anchors are computed once per tile, at its first segment).
//   fullRaw[b]  TMA bytes of buffer b have landed            (tx count, armed by producer thread 0)
//   fullMix[b]  buffer b is mixed (NCO only)                  (every producer thread arrives)
//   empty[b]    the filter warps are done reading buffer b    (every filter thread arrives)
// One tap set for all channels (hStride == 0); compile-time decimation only.
// ---------------------------------------------------------------------------------------------------------
template <int NT, int DT>
__device__ __forceinline__ void tmaStageSlowSegment(unsigned char* stageBuf, unsigned sg, const float2* src,
                                                    unsigned long long in0, unsigned rows, unsigned planeBytes,
                                                    unsigned long long nIn, const TmaParams& P, unsigned lt) {
  constexpr unsigned D = DT;
  const unsigned total = rows * 16u;  // 16 samples (8 branch pairs) of every row
  for (unsigned s = lt; s < total; s += NT) {
    const unsigned m = s >> 4, p = 16u * sg + (s & 15u);
    const unsigned long long g = in0 + (unsigned long long)m * D + p;
    const bool valid = g < nIn;
    // tmaSampleOffset addresses the whole-window layout: take the segment's base out again
    cpAsync8z(stageBuf + tmaSampleOffset<DT>(m, p, planeBytes, P) - sg * 8u * planeBytes, src + (valid ? g : 0ull), valid);
  }
}

// exact-NCO mix of one landed segment (phases 16*sg .. 16*sg + 15 of every row); arithmetic as in tmaMixRowsStatic
template <int NT, int DT>
__device__ __forceinline__ void tmaMixSegmentStatic(unsigned char* stageBuf, unsigned sg, unsigned mhCount,
                                                    unsigned planeBytes, const float2* ncoA, const float2* ncoR,
                                                    const TmaParams& P, unsigned lt) {
  constexpr unsigned CH = 4, CHUNKS = 2, G = NT / 8;
  static_assert(NT % 16 == 0, "two chunk lanes per residue");
  constexpr unsigned V_LANES = G / CHUNKS;
  const unsigned s = lt & 7u;
  const unsigned rest = lt >> 3;
  const unsigned ck = rest % CHUNKS, vLane = rest / CHUNKS;
  const unsigned sw = tmaSwizzle<DT>(s, P);
  const unsigned vTotal = 8u * ((mhCount + 7u) >> 3);
  const unsigned pp0 = 8u * sg + ck * CH;  // first branch pair of this thread's chunk
  float2 r[2 * CH];
  unsigned off[CH];
#pragma unroll
  for (unsigned j = 0; j < CH; j++) {
    const float4 rr = *reinterpret_cast<const float4*>(ncoR + 2u * (pp0 + j));
    r[2 * j] = make_float2(rr.x, rr.y);
    r[2 * j + 1] = make_float2(rr.z, rr.w);
    off[j] = ((ck * CH + j) ^ sw) << 4;
  }
  unsigned char* base = stageBuf + s * 128u;
  const float2* anchors = ncoA + s;
  for (unsigned v = vLane; v < vTotal; v += V_LANES) {
    const unsigned ml = v & 7u, u = v >> 3;
    if (s + 8u * u >= mhCount) continue;
    const float2 an = anchors[ml * mhCount + 8u * u];
    unsigned char* row = base + ml * planeBytes + u * 1024u;
    float4 x[CH];
#pragma unroll
    for (unsigned j = 0; j < CH; j++) x[j] = *reinterpret_cast<const float4*>(row + off[j]);
#pragma unroll
    for (unsigned j = 0; j < CH; j++) {
      const float2 w0 = cmulf(an, r[2 * j]);
      const float2 w1 = cmulf(an, r[2 * j + 1]);
      const float2 a = cmulf(make_float2(x[j].x, x[j].y), w0);
      const float2 c = cmulf(make_float2(x[j].z, x[j].w), w1);
      *reinterpret_cast<float4*>(row + off[j]) = make_float4(a.x, a.y, c.x, c.y);
    }
  }
}

template <int MODE, int TG, int PSPLIT, int DT, int MIXW, int MINB>
__global__ void __launch_bounds__(TG* PSPLIT + 32 * MIXW, MINB)
    firTmaWideKernel(const __grid_constant__ CUtensorMap map, const TmaParams P) {
  static_assert(DT >= 32 && DT % 16 == 0, "rows of at least two 128-byte segments, compile-time decimation");
  static_assert(MODE == kPolyFC || MODE == kPolyNcoExact, "plain FIR or the exact NCO");
  static_assert(8 % PSPLIT == 0, "the branch groups split the 8 pairs of a segment");
  constexpr unsigned NTF = TG * PSPLIT;  // filter threads
  constexpr unsigned NTM = 32 * MIXW;    // producer threads
  constexpr unsigned BOUT = kTmaR * TG;
  constexpr unsigned D = DT;
  constexpr unsigned numSegs = D / 16;
  constexpr unsigned planeBytes = tmaPlaneRows(TG, kTmaJpadCap, DT) * 128u;
  constexpr unsigned stageBytes = 8u * planeBytes;
  constexpr bool kNco = MODE == kPolyNcoExact;
  // Stage buffers.  Plain FIR: two (the copy of the next segment runs under the FIR of this one).  NCO: three —
  // landing, being mixed, being filtered — because copy (1.6 us for 72 KB at an SM's share of HBM) PLUS mix no longer
  // fit under one FIR stage (3 us); the partial-sum exchange lives in the stage buffer the tile has just finished
  // with, so three buffers fit the 227 KB.
  constexpr unsigned NBUF = kNco ? 3u : 2u;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long fullRaw[NBUF], fullMix[NBUF], emptyBar[NBUF];
  unsigned char* bufBase = smemRaw + ((1024u - (smemU32(smemRaw) & 1023u)) & 1023u);
  float* hs = reinterpret_cast<float*>(bufBase + NBUF * stageBytes);
  float2* ncoA = reinterpret_cast<float2*>(hs + (size_t)D * P.Jpad + 32u);
  float2* ncoR = ncoA + (BOUT + P.Jpad);

  const unsigned tid = threadIdx.x;
  const unsigned rowsStaged = BOUT + P.Jpad;
  if (tid == 0) {
    for (unsigned b = 0; b < NBUF; b++) {
      mbarInit(&fullRaw[b], 1);
      mbarInit(&fullMix[b], NTM);
      mbarInit(&emptyBar[b], NTF);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kNco) {
    for (unsigned p = tid; p < D; p += NTF + NTM) ncoR[p] = ncoExactPhasor((unsigned long long)p, P.ncoStep);
    for (unsigned p = tid; p < 8u; p += NTF + NTM) ncoFineOf(ncoR, D)[p] = ncoExactPhasor((unsigned long long)p * D, P.ncoStep);
  }
  {
    const unsigned nh = D * P.Jpad;
    for (unsigned i = tid; i < nh + 32u; i += NTF + NTM) {
      const unsigned pp = i / (2u * P.Jpad);
      const unsigned rem = i - pp * 2u * P.Jpad;
      const unsigned ti = (rem >> 1) * D + 2u * pp + (rem & 1u);
      hs[i] = (i < nh && ti < P.T) ? __ldg(P.h + ti) : 0.0f;
    }
  }
  __syncthreads();

  unsigned chan = blockIdx.x / P.tilesPerChannel;
  unsigned tile = blockIdx.x - chan * P.tilesPerChannel;
  auto advance = [&](unsigned& c, unsigned& tl) {
    c += P.strideChan;
    tl += P.strideTile;
    if (tl >= P.tilesPerChannel) {
      tl -= P.tilesPerChannel;
      c += 1;
    }
  };
  auto tileIsFast = [&](unsigned tl) -> bool { return tl * BOUT + rowsStaged <= P.tmaRows; };

  if (tid >= NTF) {
    // ===================== producer warps =====================
    const unsigned lt = tid - NTF;
    // fetch of one stage: wait until the filter warps have left the buffer's previous use, then one tensor copy
    // (or, for the last tile of a channel, guarded cp.async by all producer threads)
    auto fetch = [&](unsigned c, unsigned tl, unsigned sg, unsigned b, unsigned use) {
      unsigned char* buf = bufBase + b * stageBytes;
      if (use > 0) mbarWait(&emptyBar[b], (use - 1u) & 1u);
      if (tileIsFast(tl)) {
        if (lt == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbarExpectTx(&fullRaw[b], stageBytes);
          tmaLoad4(buf, &map, &fullRaw[b], (int)(sg * 32u), (int)(tl * (BOUT / 8)), 0, (int)c);
        }
      } else {
        tmaStageSlowSegment<NTM, DT>(buf, sg, P.x + (size_t)c * P.xStride, (unsigned long long)tl * BOUT * D, rowsStaged,
                                     planeBytes, P.nIn, P, lt);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        mixBarrier<NTM, 1>();
        if (lt == 0) mbarArrive(&fullRaw[b]);  // completes the phase the consumers of this stage wait for
      }
    };
    if (!kNco) {
      unsigned b = 0, use = 0;
      for (; chan < P.numChannels; advance(chan, tile)) {
        for (unsigned sg = 0; sg < numSegs; sg++) {
          fetch(chan, tile, sg, b, use);
          if (++b == NBUF) b = 0, use++;
        }
      }
      return;
    }
    // NCO: the copy of stage v + 1 is issued before stage v is mixed
    unsigned nChan = chan, nTile = tile, nSg = 0, nb = 0, nUse = 0;  // cursor of the next stage to fetch
    auto fetchNext = [&]() {
      if (nChan >= P.numChannels) return;
      fetch(nChan, nTile, nSg, nb, nUse);
      if (++nb == NBUF) nb = 0, nUse++;
      if (++nSg == numSegs) {
        nSg = 0;
        advance(nChan, nTile);
      }
    };
    fetchNext();
    unsigned b = 0, use = 0;
    const unsigned mhCount = rowsStaged >> 3;
    for (; chan < P.numChannels; advance(chan, tile)) {
      const unsigned long long in0 = (unsigned long long)tile * BOUT * D;
      for (unsigned sg = 0; sg < numSegs; sg++) {
        fetchNext();
        unsigned char* buf = bufBase + b * stageBytes;
        mbarWait(&fullRaw[b], use & 1u);
        if (sg == 0) {  // row anchors of the tile, stored [ml][mh]
          ncoRowAnchors<NTM, 1>(ncoA, ncoCoarseOf(ncoR, D), ncoFineOf(ncoR, D), rowsStaged, mhCount,
                                P.ncoRow0 + (unsigned long long)tile * BOUT, D, P.ncoRho, P.ncoStep, lt);
          mixBarrier<NTM, 1>();
        }
        tmaMixSegmentStatic<NTM, DT>(buf, sg, mhCount, planeBytes, ncoA, ncoR, P, lt);
        mbarArrive(&fullMix[b]);  // release: the mixed segment is visible to whoever acquires the barrier
        if (sg + 1 == numSegs) mixBarrier<NTM, 1>();  // ncoA may be overwritten by the next tile's anchors
        if (++b == NBUF) b = 0, use++;
      }
    }
    return;
  }

  // ============================== filter warps ==============================
  const unsigned grp = tid / TG;
  const unsigned t = tid - grp * TG;
  constexpr unsigned pairsPerGroup = 8 / PSPLIT;
  unsigned b = 0, use = 0;
  for (; chan < P.numChannels; advance(chan, tile)) {
    const unsigned long long o0 = (unsigned long long)tile * BOUT;
    float2 acc[kTmaR];
#pragma unroll
    for (int r = 0; r < kTmaR; r++) acc[r] = make_float2(0.0f, 0.0f);
    unsigned lastB = 0;
    for (unsigned sg = 0; sg < numSegs; sg++) {
      // firComputePairs addresses the whole-window layout (segment sg starts 8 * sg planes in): shift the base back
      const unsigned char* buf = bufBase + b * stageBytes - sg * stageBytes;
      mbarWait(kNco ? &fullMix[b] : &fullRaw[b], use & 1u);
      const unsigned ppBegin = 8u * sg + grp * pairsPerGroup;
      firComputePairs<DT>(acc, buf, hs, t, ppBegin, ppBegin + pairsPerGroup, P.Jpad, planeBytes, P);
      lastB = b;
      if (sg + 1 < numSegs || PSPLIT == 1) mbarArrive(&emptyBar[b]);  // no more reads of the stage by this thread
      if (++b == NBUF) b = 0, use++;
    }
    if (PSPLIT > 1) {
      // partial sums of the branch groups go through the stage buffer the tile has just finished with
      float4* red = reinterpret_cast<float4*>(bufBase + lastB * stageBytes);
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");  // every filter thread has left the window
      if (grp > 0) {
#pragma unroll
        for (int q = 0; q < kTmaR / 2; q++) {
          red[((grp - 1) * (kTmaR / 2) + q) * TG + t] =
              make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NTF) : "memory");
      if (grp == 0) {
#pragma unroll
        for (int g = 1; g < PSPLIT; g++) {
#pragma unroll
          for (int q = 0; q < kTmaR / 2; q++) {
            const float4 w = red[((g - 1) * (kTmaR / 2) + q) * TG + t];
            acc[2 * q].x += w.x;
            acc[2 * q].y += w.y;
            acc[2 * q + 1].x += w.z;
            acc[2 * q + 1].y += w.w;
          }
        }
      }
      mbarArrive(&emptyBar[lastB]);  // the buffer may now be refilled
    }
    if (grp == 0) {
      tmaStoreTile(P, chan, o0 + (unsigned long long)t * kTmaR, acc);
    }
  }
}

}  // namespace gsdr_b200
