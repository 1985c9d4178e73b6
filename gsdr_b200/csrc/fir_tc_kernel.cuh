// fir_tc_kernel.cuh — decimating FIR, complex input x real taps (gsdrFirFC), on the 5th-generation tensor cores.
// Replaces ref: src/fir.cu:49-71 (k_FirDecimate<cuComplex,cuComplex,float>) where it was measured faster than the
// FFMA2 kernel of fir_tma_kernel.cuh, which is bound by FP32 issue slots: decimation 8 with 129..264 taps (BASELINE
// config 2), decimation 4 with 65..260 taps (config 4) and decimation 16 with 257..528 taps.  The selection rule and every measurement behind it: gsdr_fir.cu tcTilesPerChannel, DESIGN.md §4.3b.
//
// Formulation (banded Toeplitz GEMM).  A window of S = 32 (decimation 4: 64) consecutive outputs starting at output o0 reads the
// K = (S-1)*D + T consecutive samples starting at sample o0*D:
//     y[o0 + s] = sum_k x[o0*D + k] * h[k - s*D]                 (h = 0 outside [0, T))
// i.e. D[row][s] = sum_k A[row][k] * B[k][s] with one ROW per (window, component, part): re and im are two real
// problems sharing B, and "part" is the FP16 head or remainder of the samples (below).
//
// Precision: FP16 operands, error-compensated.  tools/tc_probe.cu / tc_probe_f16.cu measured what the hardware gives:
// a kind::tf32 MMA (K = 8) costs 26 + 0.46 N cycles whatever the accumulator pattern — 40.5 at N = 32 — while
// kind::f16 (K = 16) runs at N/2 + 2 (17.3 at N = 32): per multiply-accumulate FP16 is ~5 x cheaper, which pays for
// splitting.  Samples (per segment of S*D samples and per component) and taps (per call) are scaled by powers of two
// into [0.5, 1) and split,
//     x*sx = xh + xl,   h*sh = hh + hl,   xh = top 11 significant bits (exact in FP16), xl = FP16(x*sx - xh),
// and ALL FOUR partial products reach the FP32 accumulators (rows of both sample parts, two MMAs per k-step for the
// two tap parts; the epilogue adds the head row and the remainder row), so nothing is dropped; what remains is the
// FP16 rounding of the remainders: <= 2^-21 relative to a value near the segment's maximum, 2^-25 of that maximum for
// small values — FP32-grade against BASELINE's tolerance 1e-5 * sum|h| * max|x| (tests/test_tc_gpu.py holds it to the
// FFMA2 kernel's own error).  A window spans two segments with two scales: two accumulators, combined by the epilogue,
// which undoes all scales exactly (powers of two).
//
// Operands.
//   The copy warp brings the tile's raw complex samples into shared memory: each byte is fetched from HBM once per
//     tile, by one bulk copy per segment, every segment on its own mbarrier.  The eight producer warps rewrite each
//     segment IN PLACE as it lands (the conversion overlaps the copies still in flight) as four planes of S*D FP16
//     values: [re head | re remainder | im head | im remainder] (8 bytes per sample either way).
//   A comes from TENSOR MEMORY: row (= TMEM lane) 32*w + l holds window 8*w + (l & 7), component (l >> 3) & 1, part
//     l >> 4; per stage of 32 samples it loads 64 contiguous bytes of its plane and writes 16 columns with one
//     tcgen05.st; a ring of four TMEM stages decouples the producers from the MMAs (the two warpgroups take
//     alternate stages).
//   B never exists as a matrix: B[k][s] depends on k - s*D only, so per tap part TWO tables are kept in shared
//     memory, T_tb[u] = (h[(aMax-u)*D + 8*tb + e])_{e<8} for the first and second group of eight k of a k-step; the
//     K-major, un-swizzled descriptor of a k-step starts 16*(aMax - a) bytes into table 0 (rows s = 0..31 of the
//     operand are the next 32 entries: 16-byte row pitch, SBO = 128; LBO = table pitch).  A few KB serve every step.
//   D (2 x 128 x 32 FP32) lives in TMEM; the producer warps read it back with tcgen05.ld when the tile's last MMA has
//     been committed.
// A tile is 32 windows (1024 outputs; 2048 at decimation 4) of one channel; CTAs are persistent over tiles, three per
// SM (two at decimation 4, whose CTAs hold 256 TMEM columns; one at decimation 16, whose samples fill the shared memory; shared memory: 33 segments + 16 bytes of padding each,
// so the 8 windows a warp reads hit different banks) — while one CTA computes, another's bulk copies are in flight.  Segments that reach past the
// caller-guaranteed input are staged by the copy warp with guarded loads and zero fill, and the epilogue masks outputs
// >= numOutputs.
//
// What a result depends on: the tile's own samples and the output's position in the tile (k-step alignment, segment
// scales) — nothing else; the tile's last segment is masked down to the T - D samples its one reader uses, so the
// next tile's samples never set a scale.  Calls whose first outputs differ by a multiple of the tile therefore agree
// bit for bit on the outputs they share: gsdrShardPlanTime and the host pipeline cut shards / chunks on that grid.
//
// Non-finite samples: an Inf/NaN keeps its segment's scale (of its component) at 1 and reaches every output of the
// windows that read the segment (0 * Inf in the band's zeros); the reference confines it to the outputs whose taps
// overlap it (tests/test_tc_gpu.py pins the reach).
#pragma once

#include <cuda_fp16.h>

#include "fir_tma_kernel.cuh"

namespace gsdr_b200 {

struct TcParams {
  const float2* x;
  const float* h;
  float2* y;
  unsigned long long nOut, nIn;         // per channel: outputs, and the (nOut - 1) * D + T samples the caller guarantees
  unsigned long long xStride, yStride;  // channel strides, elements
  unsigned tilesPerChannel, totalTiles;
  unsigned T;
  unsigned numStages;   // stages of 32 samples (two k-steps of 16) per tile: ceil(((S-1)*D + T) / 32)
  unsigned aMax;        // (16 * (2 * numStages - 1)) / D: tap-row offset of the last k-step
  unsigned tablePitch;  // bytes between the two tables of a tap part: (aMax + S) * 16
};

constexpr int kTcWindows = 32;    // windows per tile
// outputs per window = MMA N: 32, or 64 for decimation 4 (a segment of S*D = 256 samples either way)
__host__ __device__ constexpr int tcWindowOutputs(int D) { return D == 4 ? 64 : 32; }
// The ring of TMEM stages (16 columns each) between the producers and the MMAs:
//   D = 8 (S = 32): four stages next to the two accumulators of 32 columns in ONE allocation of 128 columns, three CTAs per
//           SM.  (Measured alternatives: six stages, two of them in a second allocation of 32 columns: 3 % slower;
//           eight stages in a second allocation of 128 columns with two CTAs per SM: 5 % slower.)
//   D = 4 (S = 64): the two accumulators of 64 columns fill the first allocation; eight stages in a second allocation of 128
//           columns, two CTAs per SM (0.1586 ms at decimation 4, 127 taps — with two stages in a second allocation
//           of 32 columns and three CTAs per SM: 0.174 ms).
//   D = 16 (one CTA per SM: its samples fill the shared memory): sixteen stages in a second allocation of 256 columns.
__host__ __device__ constexpr int tcRing(int D) { return D == 4 ? 8 : D == 16 ? 16 : 4; }
__host__ __device__ constexpr unsigned tcTmemCols2(int D) { return D == 4 ? 128u : D == 16 ? 256u : 0u; }
constexpr int kTcProducers = 256;  // warps 0-7: two warpgroups of producers + epilogue (warp w and w + 4 share the
                                   // TMEM lanes 32 * (w & 3) ..: the groups take alternate stages)
constexpr int kTcThreads = kTcProducers + 64;  // warp 8: MMA issue, warp 9: bulk copies
constexpr unsigned kTcTmemCols = 128;

template <int D>
struct TcGeom {
  static_assert(D == 4 || D == 8 || D == 16, "k-steps of 16 samples must start on a tap row");
  static constexpr int S = tcWindowOutputs(D);
  static constexpr int tileOut = S * kTcWindows;
  static constexpr int ring = tcRing(D);
  static constexpr unsigned SD = S * D;                // samples per segment = window stride
  static constexpr unsigned segBytes = SD * 8;         // raw: SD complex FP32; converted: four planes of SD FP16
  static constexpr unsigned planeBytes = SD * 2;
  static constexpr unsigned segPitch = segBytes + 16;  // 16 bytes of padding: consecutive windows, different banks
  static constexpr unsigned numSegs = kTcWindows + 1;
  static constexpr unsigned rawBytes = numSegs * segPitch;
  static constexpr unsigned maxTaps = SD + D;          // (S-1)*D + T <= 2 * SD: a window fits two segments
};

__host__ __device__ inline unsigned tcTableBytes(unsigned tablePitch) { return 4u * tablePitch; }  // {head, rem} x 2

__device__ __forceinline__ unsigned tcElectOne() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tcMmaF16(unsigned dTmem, unsigned aTmem, unsigned long long bDesc, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(dTmem),
      "r"(aTmem), "l"(bDesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcCommit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemU32(bar))
               : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tcStore16(unsigned taddr, const uint4 (&v)[4]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w),
      "r"(v[2].x), "r"(v[2].y), "r"(v[2].z), "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w)
      : "memory");
}
__device__ __forceinline__ void tcLoad16(unsigned taddr, unsigned (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// kind::f16 with FP16 A and B (format 0), FP32 accumulate, A and B K-major, M = 128, N = 32
// (bit layout: cute/arch/mma_sm100_desc.hpp)
__host__ __device__ constexpr unsigned tcIdesc(int S) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((unsigned)(S >> 3) << 17) | ((128u >> 4) << 24);
}
constexpr unsigned kTcHeadMask = 0xFFFFE000u;  // sign, exponent, 10 mantissa bits: an FP16 value when in range

// exponent n of the power of two 2^n that brings a value with |bits| `absBits` into [0.5, 1); 0 for Inf and NaN
__device__ __forceinline__ int tcScaleExponent(unsigned absBits) {
  const int e = (int)(absBits >> 23);  // biased exponent; 0 for zero / subnormal: 2^126 keeps those below 1
  if (e == 255) return 0;
  const int n = 126 - e;
  return n < -125 ? -125 : n;
}
__device__ __forceinline__ float tcPow2(int n) { return __uint_as_float((unsigned)(n + 127) << 23); }  // -126 <= n <= 127

// The kernel is capped at 64 registers (three CTAs of ten warps per SM) and ptxas would rather recompute thread ids and
// shared-window addresses inside the stage loop (S2R / S2UR + address arithmetic on the critical path of every stage)
// than keep them: values passed through here are opaque to it and stay in their register.
__device__ __forceinline__ unsigned tcKeep(unsigned v) {
  unsigned r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ void tcBarWait(unsigned barAddr, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(barAddr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint4 tcLoadShared16(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tcBarArrive(unsigned barAddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(barAddr) : "memory");
}

__device__ __forceinline__ unsigned tcPackHalf2(float lo, float hi) {
  const __half2 p = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const unsigned*>(&p);
}

template <int D, int MINB>
__global__ void __launch_bounds__(kTcThreads, MINB) firTcKernel(const TcParams P) {
  using G = TcGeom<D>;
  constexpr int S = G::S, kTcRing = G::ring;
  constexpr unsigned kTcIdesc = tcIdesc(S);
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long segFull[G::numSegs], rawEmpty, dFull, dEmpty, aFull[kTcRing], aEmpty[kTcRing];
  __shared__ int segExp[2][G::numSegs];  // per component (re, im) and segment: the exponent it was scaled by
  __shared__ unsigned tmemBaseSlot, tmemBaseSlot2;
  __shared__ unsigned redMax[kTcThreads / 32];
  unsigned char* raw = smemRaw;                  // numSegs x segPitch
  unsigned char* tables = smemRaw + G::rawBytes;  // [head | remainder][2][aMax + S] x 16 bytes (8 FP16 taps)

  const unsigned tid = tcKeep(threadIdx.x), warp = tid >> 5, lane = tid & 31u;
  constexpr unsigned kMmaWarp = kTcProducers / 32, kNumWarps = kTcThreads / 32;
  if (tid == 0) {
    for (unsigned i = 0; i < G::numSegs; i++) mbarInit(&segFull[i], 1);
    mbarInit(&rawEmpty, kTcProducers);
    mbarInit(&dFull, 1);
    mbarInit(&dEmpty, kTcProducers);
    for (int i = 0; i < kTcRing; i++) {
      mbarInit(&aFull[i], 128);
      mbarInit(&aEmpty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBaseSlot)),
                 "r"(kTcTmemCols)
                 : "memory");
    if (tcTmemCols2(D) > 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBaseSlot2)),
                   "r"(tcTmemCols2(D))
                   : "memory");
    }
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // tap scale: the largest |h| into [0.5, 1)
  {
    unsigned m = 0;
    for (unsigned i = tid; i < P.T; i += kTcThreads) m = max(m, __float_as_uint(__ldg(P.h + i)) & 0x7fffffffu);
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) redMax[warp] = m;
  }
  __syncthreads();
  unsigned hm = 0;
  for (unsigned w = 0; w < kNumWarps; w++) hm = max(hm, redMax[w]);
  const int tapExp = tcScaleExponent(hm);
  const float tapScale = tcPow2(tapExp);
  // tap tables: entry u of table tb holds taps (aMax - u)*D + 8*tb + e, e < 8 (0 outside [0, T)), scaled, as FP16 head
  // and FP16 remainder
  {
    const unsigned entries = P.tablePitch >> 4;
    const unsigned total = 2u * entries * 8u;
    __half* head = reinterpret_cast<__half*>(tables);
    __half* rem = reinterpret_cast<__half*>(tables + 2u * P.tablePitch);
    for (unsigned i = tid; i < total; i += kTcThreads) {
      const unsigned e = i & 7u, u = (i >> 3) % entries, tb = (i >> 3) / entries;
      const long long ti = ((long long)P.aMax - (long long)u) * D + 8 * tb + e;
      const float v = (ti >= 0 && ti < (long long)P.T) ? __ldg(P.h + ti) * tapScale : 0.0f;
      const float hd = __uint_as_float(__float_as_uint(v) & kTcHeadMask);
      head[i] = __float2half_rn(hd);
      rem[i] = __float2half_rn(v - hd);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tables are read by the tensor core
  tcFenceBefore();
  __syncthreads();
  tcFenceAfter();
  const unsigned tmem = tmemBaseSlot;
  const unsigned ringCols = tcTmemCols2(D) > 0 ? tmemBaseSlot2 : tmem + 2u * S;  // first column of the A ring
  // two accumulators of S columns (a window's samples in its first / second segment: the segments have their own
  // scales); A ring: kTcRing x 16 columns
  const unsigned colD = 0;
  const bool useSecond = P.numStages * 32u > G::SD;
  const unsigned numStages = P.numStages;

  if (warp < kMmaWarp) {
    // ===================== producers (then epilogue) =====================
    // lane -> (part, component, window): the eight windows of a warp sit in eight different bank groups (16 bytes of
    // segment padding), the four rows of a window read four different planes
    const unsigned group = warp >> 2, quad = warp & 3u;  // warpgroup; TMEM lane quadrant
    const unsigned part = lane >> 4;                     // 0: head rows, 1: remainder rows
    const unsigned comp = (lane >> 3) & 1u;              // 0: re, 1: im
    const unsigned q = 8u * quad + (lane & 7u);          // window of the tile
    const unsigned laneAddr = tcKeep((32u * quad) << 16);
    const unsigned aFullAddr = tcKeep(smemU32(&aFull[0])), aEmptyAddr = tcKeep(smemU32(&aEmpty[0]));
    const unsigned ringBase = tcKeep(ringCols + laneAddr);
    const unsigned rowAddr = tcKeep(smemU32(raw + q * G::segPitch + (2u * comp + part) * G::planeBytes));
    const unsigned numStagesK = tcKeep(numStages);
    const unsigned lastSegValid = P.T > (unsigned)D ? P.T - (unsigned)D : 0u;  // (S-1)*D + T - S*D
    unsigned g = 0;                 // stage counter over all tiles of this CTA: stage g belongs to warpgroup g & 1
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < P.totalTiles; tile += gridDim.x, it++) {
      // ---- every segment in place, as it lands: SD complex FP32 -> planes [re head | re rem | im head | im rem] of
      //      SD FP16, scaled by the power of two that brings the segment's largest |component| into [0.5, 1) ----
      // one warp per segment (taking them from a shared counter instead, 33 over 8 warps, measured 1 % slower); a
      // lane holds chunk 32*p + lane of every pass p before the first plane word is written
      for (unsigned sg = warp; sg < G::numSegs; sg += kMmaWarp) {
        unsigned char* seg = raw + sg * G::segPitch;
        constexpr int kPasses = G::SD / 64;
        float4 c[kPasses];
        mbarWait(&segFull[sg], it & 1u);
        // the largest |component| (fmaxf drops NaNs: a NaN stays a NaN under any scale; an Inf gives scale 1)
#pragma unroll
        for (int p = 0; p < kPasses; p++) c[p] = *reinterpret_cast<const float4*>(seg + 16u * (32u * p + lane));
        if (sg == G::numSegs - 1u) {
          // the tile's last segment serves the last window only, which reads its first T - D samples: the others (the
          // next tile's, or nothing at the end of a call) must not set the scale — that is what makes a tile's
          // results a function of the tile's own samples, and aligned time shards bit-identical to the whole call
#pragma unroll
          for (int p = 0; p < kPasses; p++) {
            const unsigned i0 = 2u * (32u * p + lane);
            if (i0 >= lastSegValid) c[p].x = c[p].y = 0.0f;
            if (i0 + 1u >= lastSegValid) c[p].z = c[p].w = 0.0f;
          }
        }
        // re and im are separate rows of the GEMM, so each gets its own scale: a component much smaller than the other
        // keeps its own relative precision (and an Inf in one does not reach the other)
        float mRe = 0.0f, mIm = 0.0f;
#pragma unroll
        for (int p = 0; p < kPasses; p++) {
          mRe = fmaxf(mRe, fmaxf(fabsf(c[p].x), fabsf(c[p].z)));
          mIm = fmaxf(mIm, fmaxf(fabsf(c[p].y), fabsf(c[p].w)));
        }
        // (the reductions also order every lane's loads before the stores below)
        const int reExp = tcScaleExponent(__reduce_max_sync(0xffffffffu, __float_as_uint(mRe)));
        const int imExp = tcScaleExponent(__reduce_max_sync(0xffffffffu, __float_as_uint(mIm)));
        const float sRe = tcPow2(reExp), sIm = tcPow2(imExp);
        if (lane == 0) {
          segExp[0][sg] = reExp;
          segExp[1][sg] = imExp;
        }
#pragma unroll
        for (int p = 0; p < kPasses; p++) {
          const float a = c[p].x * sRe, b = c[p].z * sRe;  // re of samples 2i, 2i+1
          const float e = c[p].y * sIm, f = c[p].w * sIm;  // im
          const float ah = __uint_as_float(__float_as_uint(a) & kTcHeadMask);
          const float bh = __uint_as_float(__float_as_uint(b) & kTcHeadMask);
          const float eh = __uint_as_float(__float_as_uint(e) & kTcHeadMask);
          const float fh = __uint_as_float(__float_as_uint(f) & kTcHeadMask);
          unsigned* w = reinterpret_cast<unsigned*>(seg) + 32u * p + lane;
          w[0] = tcPackHalf2(ah, bh);
          w[G::planeBytes / 4] = tcPackHalf2(a - ah, b - bh);
          w[2 * G::planeBytes / 4] = tcPackHalf2(eh, fh);
          w[3 * G::planeBytes / 4] = tcPackHalf2(e - eh, f - fh);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kTcProducers) : "memory");  // planes and segExp complete
      // ---- stages: 32 samples of this row's plane -> 16 TMEM columns; the warpgroups take alternate stages ----
      // a warp's own stages are st0, st0 + 2, ...; the loads of its next stage are in flight while the tcgen05.st of
      // the current one completes
      auto stageSrc = [&](unsigned st) {  // shared-memory address of the row's 64 bytes of stage st
        const unsigned k = 32u * st;
        const unsigned over = k >= G::SD ? 1u : 0u;  // second segment of the window (stages never straddle)
        return rowAddr + over * G::segPitch + (k - over * G::SD) * 2u;
      };
      const unsigned st0 = (g ^ group) & 1u;
      unsigned slot = (g + st0) % kTcRing, ringPass = (g + st0) / kTcRing;
      uint4 av[4];
      if (st0 < numStagesK) {
        const unsigned src = stageSrc(st0);
#pragma unroll
        for (int j = 0; j < 4; j++) av[j] = tcLoadShared16(src + 16u * j);
      }
      for (unsigned st = st0; st < numStagesK; st += 2) {
        if (ringPass > 0) {
          tcBarWait(aEmptyAddr + 8u * slot, (ringPass - 1u) & 1u);  // the MMAs that read this slot have completed
          tcFenceAfter();
        }
        tcStore16(ringBase + 16u * slot, av);
        if (st + 2 < numStagesK) {
          const unsigned src = stageSrc(st + 2);
#pragma unroll
          for (int j = 0; j < 4; j++) av[j] = tcLoadShared16(src + 16u * j);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcFenceBefore();
        tcBarArrive(aFullAddr + 8u * slot);
        slot += 2;
        if (slot >= kTcRing) {
          slot -= kTcRing;
          ringPass++;
        }
      }
      g += numStages;
      // (read before the arrive: the next tile's conversion rewrites segExp as soon as its copies have landed)
      const int e1 = -(segExp[comp][q] + tapExp), e2 = -(segExp[comp][q + 1] + tapExp);
      mbarArrive(&rawEmpty);  // this thread has no more reads of the tile's samples
      // ---- epilogue: head row + remainder row -> y, undoing both scales (exact powers of two, in two factors to
      //      stay in range); warpgroup 0 takes the accumulator's columns 0..15, warpgroup 1 columns 16..31 ----
      mbarWait(&dFull, it & 1u);
      tcFenceAfter();
      // v[i] = D1[i] / scale(first segment) + D2[i] / scale(second segment), the tap scale undone in the same factors
      const float f1a = tcPow2(e1 / 2), f1b = tcPow2(e1 - e1 / 2), f2a = tcPow2(e2 / 2), f2b = tcPow2(e2 - e2 / 2);
      const unsigned chan = tile / P.tilesPerChannel;
      const unsigned tl = tile - chan * P.tilesPerChannel;
      // every warpgroup takes half of the accumulators' columns, 16 at a time
#pragma unroll
      for (int cp = 0; cp < S / 32; cp++) {
        const unsigned col0 = (unsigned)(S / 2) * group + 16u * cp;  // first of this pass's 16 outputs of the window
        unsigned d[16];
        float v[16];
        tcLoad16(tmem + laneAddr + colD + col0, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (__uint_as_float(d[i]) * f1a) * f1b;
        if (useSecond) {
          tcLoad16(tmem + laneAddr + colD + S + col0, d);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; i++) v[i] = fmaf(__uint_as_float(d[i]) * f2a, f2b, v[i]);
        }
        if (cp == S / 32 - 1) {
          tcFenceBefore();
          mbarArrive(&dEmpty);  // the accumulators may be overwritten by the next tile
        }
        // the four lanes of a window each end up with four complete outputs: col0 + 8*part + 4*comp + (0..3).
        // 1) the two parts trade the half of their columns the other one sums
        float sum[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float give = part ? v[i] : v[8 + i];
          const float keep = part ? v[8 + i] : v[i];
          sum[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
        }
        // 2) the two components trade the half of those outputs the other one stores
        float2 out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float give = comp ? sum[i] : sum[4 + i];
          const float keep = comp ? sum[4 + i] : sum[i];
          const float got = __shfl_xor_sync(0xffffffffu, give, 8);
          out[i] = comp ? make_float2(got, keep) : make_float2(keep, got);
        }
        const unsigned long long o0 = (unsigned long long)tl * G::tileOut + q * S + col0 + 8u * part + 4u * comp;
        float2* y = P.y + (size_t)chan * P.yStride + o0;
        const unsigned valid = o0 >= P.nOut ? 0u : (P.nOut - o0 >= 4ull ? 4u : (unsigned)(P.nOut - o0));
        if (valid == 4u && (reinterpret_cast<uintptr_t>(y) & 15u) == 0) {
#pragma unroll
          for (int i = 0; i < 2; i++) {
            reinterpret_cast<float4*>(y)[i] =
                make_float4(out[2 * i].x, out[2 * i].y, out[2 * i + 1].x, out[2 * i + 1].y);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            if ((unsigned)i < valid) y[i] = out[i];
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issue =====================
    // This warp shares its scheduler with seven others, so every instruction between a stage becoming ready and its
    // MMAs being issued is latency on the ring: descriptors, TMEM addresses and barrier addresses advance by constant
    // steps instead of being recomputed (the k-step's tap rows move down by 16/D table entries per step).
    const unsigned leader = tcElectOne();
    const unsigned long long descHigh =
        ((unsigned long long)((128u >> 4) | (1u << 14)) << 32) | ((unsigned long long)((P.tablePitch >> 4) & 0x3FFFu) << 16);
    const unsigned long long descHead0 = descHigh | (unsigned long long)((smemU32(tables) + 16u * P.aMax) >> 4);
    const unsigned remDelta = (2u * P.tablePitch) >> 4;  // head table -> remainder table, in 16-byte units
    constexpr unsigned kStepDelta = 16u / D;             // table entries (= 16-byte units) per k-step
    constexpr unsigned kSecondFrom = G::SD / 32u;        // first stage of the window's second segment
    const unsigned aFullAddr = tcKeep(smemU32(&aFull[0])), aEmptyAddr = tcKeep(smemU32(&aEmpty[0]));
    const unsigned ring0 = tcKeep(ringCols);
    unsigned slot = 0, ringPass = 0, it = 0;
    for (unsigned tile = blockIdx.x; tile < P.totalTiles; tile += gridDim.x, it++) {
      if (it > 0) {
        __nanosleep(500);  // the next tile's first stage is further away than that (copies, conversion)
        mbarWait(&dEmpty, (it - 1u) & 1u);  // the epilogue has read the previous tile's accumulators
        tcFenceAfter();
      }
      unsigned long long bHead = descHead0;
      for (unsigned st = 0; st < numStages; st++) {
        tcBarWait(aFullAddr + 8u * slot, ringPass & 1u);  // (a suspended wait costs 4 % here)
        tcFenceAfter();
        if (leader) {
          const unsigned acc = tmem + colD + (st >= kSecondFrom ? (unsigned)S : 0u);
          const unsigned first = (st == 0u || st == kSecondFrom) ? 0u : 1u;
          const unsigned aCols = ring0 + 16u * slot;
          tcMmaF16(acc, aCols, bHead, kTcIdesc, first);
          tcMmaF16(acc, aCols, bHead + remDelta, kTcIdesc, 1u);
          tcMmaF16(acc, aCols + 8u, bHead - kStepDelta, kTcIdesc, 1u);
          tcMmaF16(acc, aCols + 8u, bHead - kStepDelta + remDelta, kTcIdesc, 1u);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           aEmptyAddr + 8u * slot)
                       : "memory");
          if (st + 1 == numStages) tcCommit(&dFull);
        }
        bHead -= 2u * kStepDelta;
        __syncwarp();
        if (++slot == kTcRing) {
          slot = 0;
          ringPass++;
        }
      }
    }
  } else {
    // ===================== copy warp: 33 segments per tile =====================
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < P.totalTiles; tile += gridDim.x, it++) {
      const unsigned chan = tile / P.tilesPerChannel;
      const unsigned tl = tile - chan * P.tilesPerChannel;
      const float2* src = P.x + (size_t)chan * P.xStride;
      const unsigned long long s0 = (unsigned long long)tl * G::tileOut * D;  // first sample of the tile
      // the producers need well over a microsecond for a tile: sleeping through the first part of the wait leaves
      // the issue slots of the spin to them (0.1451 -> 0.1385 ms on config 2 together with the MMA warp's sleep)
      if (it > 0) __nanosleep(800);
      if (it > 0) mbarWait(&rawEmpty, (it - 1u) & 1u);  // every producer has left the previous tile's samples
      // segments inside the caller-guaranteed extent (a prefix): one bulk copy each, every one on its own barrier so
      // that its conversion starts when it lands; the others: guarded loads, zero fill
      const unsigned long long room = P.nIn > s0 ? (P.nIn - s0) / G::SD : 0ull;
      const unsigned fast = room < G::numSegs ? (unsigned)room : G::numSegs;
      // (one lane issues all the copies, in segment order: 32 lanes issuing one each was measured 7 % slower — the
      // segments then land out of order and the conversion, which takes them in order, starts late)
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (unsigned sg = 0; sg < fast; sg++) {
          mbarExpectTx(&segFull[sg], G::segBytes);  // also the one arrival the barrier waits for
          bulkLoad1d(raw + sg * G::segPitch, src + s0 + (size_t)sg * G::SD, G::segBytes, &segFull[sg]);
        }
      }
      for (unsigned sg = fast; sg < G::numSegs; sg++) {
        const unsigned long long b = s0 + (unsigned long long)sg * G::SD;
        float2* dst = reinterpret_cast<float2*>(raw + sg * G::segPitch);
        for (unsigned i = lane; i < G::SD; i += 32) {
          dst[i] = (b + i < P.nIn) ? __ldg(src + b + i) : make_float2(0.0f, 0.0f);
        }
        __syncwarp();
        if (lane == 0) mbarArrive(&segFull[sg]);
      }
      __syncwarp();
    }
  }
  tcFenceBefore();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
    if (tcTmemCols2(D) > 0) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBaseSlot2), "r"(tcTmemCols2(D))
                   : "memory");
    }
  }
}

}  // namespace gsdr_b200
