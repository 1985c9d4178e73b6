// fir_tc_kernel.cuh — decimating FIR, complex input x real taps (gsdrFirFC), on the 5th-generation tensor cores.
// Replaces ref: src/fir.cu:49-71 (k_FirDecimate<cuComplex,cuComplex,float>) for shapes with many taps per output
// (tapCount / decimation around 32: BASELINE configs 2 and 4), where the FFMA2 kernel of fir_tma_kernel.cuh is bound
// by FP32 issue slots and this one by HBM.
//
// Formulation (banded Toeplitz GEMM, error-compensated TF32).  A window of S = 32 consecutive outputs starting at
// output o0 reads the K = (S-1)*D + T consecutive samples starting at sample o0*D:
//     y[o0 + s] = sum_k x[o0*D + k] * h[k - s*D]                 (h = 0 outside [0, T))
// i.e. D[row][s] = sum_k A[row][k] * B[k][s] with one ROW per (window, component): re and im are two real problems
// sharing B.  FP32 accuracy comes from splitting both operands, x = xh + xl and h = hh + hl (xh, hh = the 11
// significant bits tcgen05 keeps of an FP32 operand — it truncates —, xl, hl the exact remainders), and summing
// all four partial products in the FP32 accumulator:
//   * the M = 128 rows of one MMA are 32 windows x 2 components x {hi part, lo part} of the samples,
//   * two MMAs per k-step, one against the hi taps and one against the lo taps, accumulate into the same tile,
//   * the epilogue adds the hi-part row and the lo-part row of each (window, component).
// tools/tc_probe.cu measures the pieces: truncation, descriptors, and an error of ~1e-6 relative for the split.
//
// Operands.
//   A comes from TENSOR MEMORY: four producer warps read the raw complex samples of the tile from shared memory
//     (each byte is fetched from HBM exactly once per tile by 33 bulk copies; the two windows that overlap a sample
//     both read it from shared memory), pick their component, form the hi / lo part and write 16 columns (two
//     k-steps of 8) per stage with tcgen05.st; a 4-stage ring of TMEM columns decouples them from the MMA.
//     Row (= TMEM lane) 32*w + l holds window 8*w + (l & 7), component (l >> 3) & 1, part l >> 4.
//   B never exists as a matrix: B[k][s] depends on k - s*D only, so for every residue of k modulo D (in groups of
//     four consecutive k = one 16-byte core-matrix row) ONE table T[u] = (h[(aMax-u)*D + off .. +3]) is kept in
//     shared memory, and the K-major, un-swizzled descriptor of a k-step simply starts 16*(aMax - a) bytes into the
//     table: rows s = 0..31 of the operand are the next 32 entries (16-byte row pitch, SBO = 128), the second
//     k-group of the step is the next table (LBO = table pitch).  A few KB of taps serve every step of every tile.
//   D (128 x 32 FP32) lives in TMEM; the producer warps read it back with tcgen05.ld when the tile's last MMA has
//     been committed.
// Useful MACs / issued MACs = T / K (about one half) x 3/4 (the lo*lo product is computed but not needed), and the
// kernel still has 1.5x headroom over the HBM time of config 2, which is what bounds it.
//
// A tile is 1024 outputs (32 windows) of one channel; CTAs are persistent over tiles, 2-4 CTAs per SM (shared
// memory: 33 sample segments of S*D samples + 16 bytes of padding each, so the 8 windows a warp reads hit
// different banks) — while one CTA computes, the other's bulk copies are in flight.  Segments that reach past the
// caller-guaranteed input (the last tile or two of a channel) are staged by the copy warp with guarded loads and zero
// fill instead of a bulk copy, and the epilogue masks outputs >= numOutputs: every output of a call goes through the
// same arithmetic, whatever its position — which is what makes time shards reproduce the unsharded bits.
//
// Non-finite samples: a window's Inf/NaN reaches all 32 outputs of its window row (0 * Inf in the band's zeros);
// the reference would confine it to the outputs whose taps overlap it (documented in include/gsdr/fir.h).
#pragma once

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
  unsigned numStages;   // 16-sample stages per tile: ceil(((S-1)*D + T) / 16)
  unsigned aMax;        // largest (8*j) / D over the tile's k-steps j
  unsigned tablePitch;  // bytes between the tables of consecutive k-groups: (aMax + S) * 16
};

constexpr int kTcS = 32;          // outputs per window = MMA N
constexpr int kTcWindows = 32;    // windows per tile
constexpr int kTcTileOut = kTcS * kTcWindows;
constexpr int kTcRing = 4;        // TMEM stages of 16 columns
constexpr int kTcThreads = 192;   // warps 0-3: producers + epilogue, warp 4: MMA issue, warp 5: bulk copies
constexpr unsigned kTcTmemCols = 128;

template <int D>
struct TcGeom {
  static_assert(D == 4 || D == 8 || D == 16, "segment must fit shared memory; k-groups must not straddle rows");
  static constexpr unsigned SD = kTcS * D;             // samples per segment = window stride
  static constexpr unsigned segBytes = SD * 8;
  static constexpr unsigned segPitch = segBytes + 16;  // 16 bytes of padding: consecutive windows, different banks
  static constexpr unsigned numSegs = kTcWindows + 1;
  static constexpr unsigned rawBytes = numSegs * segPitch;
  static constexpr unsigned numTables = (D >= 8 ? D : 8) / 4;
  static constexpr unsigned maxTaps = SD + D;          // the window must fit two segments
};

__host__ __device__ inline unsigned tcTableBytes(unsigned D, unsigned tablePitch) {
  return 2u * ((D >= 8 ? D : 8) / 4) * tablePitch;  // hi and lo parts
}

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
__device__ __forceinline__ void tcMma(unsigned dTmem, unsigned aTmem, unsigned long long bDesc, unsigned idesc,
                                      unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
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

__device__ __forceinline__ void tcStore16(unsigned taddr, const unsigned (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tcLoad32(unsigned taddr, unsigned (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 32 (bit layout: cute/arch/mma_sm100_desc.hpp)
constexpr unsigned kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kTcS >> 3) << 17) | ((128u >> 4) << 24);

template <int D, int MINB>
__global__ void __launch_bounds__(kTcThreads, MINB) firTcKernel(const TcParams P) {
  using G = TcGeom<D>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ __align__(8) unsigned long long rawFull, rawEmpty, dFull, dEmpty, aFull[kTcRing], aEmpty[kTcRing];
  __shared__ unsigned tmemBaseSlot;
  unsigned char* raw = smemRaw;                  // numSegs x segPitch
  unsigned char* tables = smemRaw + G::rawBytes;  // [hi | lo][numTables][aMax + S] x 16 bytes

  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  if (tid == 0) {
    mbarInit(&rawFull, 1);
    mbarInit(&rawEmpty, 128);
    mbarInit(&dFull, 1);
    mbarInit(&dEmpty, 128);
    for (int i = 0; i < kTcRing; i++) {
      mbarInit(&aFull[i], 128);
      mbarInit(&aEmpty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32(&tmemBaseSlot)),
                 "r"(kTcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // tap tables: entry u of table tb holds h[(aMax - u)*D + 4*tb + e], e < 4 (0 outside [0, T)); the hi part is the
  // tap itself (the tensor core keeps its upper 19 bits), the lo part the exact remainder
  {
    const unsigned entries = P.tablePitch >> 4;
    const unsigned total = G::numTables * entries * 4u;
    float* hi = reinterpret_cast<float*>(tables);
    float* lo = reinterpret_cast<float*>(tables + G::numTables * P.tablePitch);
    for (unsigned i = tid; i < total; i += kTcThreads) {
      const unsigned e = i & 3u, u = (i >> 2) % entries, tb = (i >> 2) / entries;
      const long long ti = ((long long)P.aMax - (long long)u) * D + 4 * tb + e;
      const float v = (ti >= 0 && ti < (long long)P.T) ? __ldg(P.h + ti) : 0.0f;
      hi[i] = v;
      lo[i] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tables are read by the tensor core
  tcFenceBefore();
  __syncthreads();
  tcFenceAfter();
  const unsigned tmem = tmemBaseSlot;
  const unsigned colD = 0, colA = kTcS;  // accumulator: 32 columns; A ring: kTcRing x 16 columns
  const unsigned numStages = P.numStages;

  if (warp < 4) {
    // ===================== producers (then epilogue) =====================
    // lane -> (part, component, window): the four lanes that share a window read the same shared-memory words
    // (broadcast), and the eight windows of a warp sit in eight different bank groups (16 bytes of segment padding)
    const unsigned part = lane >> 4;                       // 0: hi part rows, 1: lo part rows
    const unsigned comp = (lane >> 3) & 1u;                // 0: re, 1: im
    const unsigned q = 8u * warp + (lane & 7u);            // window of the tile
    const unsigned mask = part ? 0xFFFFE000u : 0u;         // value = x - (x & mask): x itself, or its low part
    const unsigned laneAddr = (32u * warp) << 16;
    unsigned g = 0;  // running stage counter over all tiles of this CTA
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < P.totalTiles; tile += gridDim.x, it++) {
      mbarWait(&rawFull, it & 1u);
      const unsigned char* win = raw + q * G::segPitch + comp * 4u;
      for (unsigned st = 0; st < numStages; st++, g++) {
        const unsigned slot = g % kTcRing, n = g / kTcRing;
        const unsigned k = 16u * st;
        // samples k .. k+15 of this window: segment q (k < SD) or q + 1
        const unsigned char* src = win + (k >= G::SD ? G::segPitch + (k - G::SD) * 8u : k * 8u);
        unsigned v[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          // 16 bytes = samples 2i, 2i+1 (re, im, re, im); comp shifted the base by 4 bytes: .x/.z are ours
          const float a = *reinterpret_cast<const float*>(src + 16 * i);
          const float b = *reinterpret_cast<const float*>(src + 16 * i + 8);
          v[2 * i] = __float_as_uint(a - __uint_as_float(__float_as_uint(a) & mask));
          v[2 * i + 1] = __float_as_uint(b - __uint_as_float(__float_as_uint(b) & mask));
        }
        if (n > 0) {
          mbarWait(&aEmpty[slot], (n - 1u) & 1u);  // the MMAs that read this ring slot have completed
          tcFenceAfter();
        }
        tcStore16(tmem + laneAddr + colA + 16u * slot, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcFenceBefore();
        mbarArrive(&aFull[slot]);
      }
      mbarArrive(&rawEmpty);  // this thread has no more reads of the raw tile
      // ---- epilogue: D rows (hi part + lo part) -> y ----
      mbarWait(&dFull, it & 1u);
      tcFenceAfter();
      unsigned d[32];
      tcLoad32(tmem + laneAddr + colD, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tcFenceBefore();
      mbarArrive(&dEmpty);  // the accumulator may be overwritten by the next tile
      const unsigned chan = tile / P.tilesPerChannel;
      const unsigned tl = tile - chan * P.tilesPerChannel;
      const unsigned long long o0 = (unsigned long long)tl * kTcTileOut + q * kTcS;  // first output of the window
      float* y = reinterpret_cast<float*>(P.y + (size_t)chan * P.yStride + o0) + comp;
      const unsigned valid = o0 >= P.nOut ? 0u : (P.nOut - o0 >= (unsigned long long)kTcS ? (unsigned)kTcS : (unsigned)(P.nOut - o0));
#pragma unroll
      for (int s = 0; s < 32; s++) {
        const float mine = __uint_as_float(d[s]);
        const float sum = mine + __shfl_xor_sync(0xffffffffu, mine, 16);
        // lanes of the hi part store outputs 0..15 of the window, lanes of the lo part outputs 16..31
        if ((unsigned)(s >> 4) == part && (unsigned)s < valid) y[2 * s] = sum;
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issue =====================
    const unsigned leader = tcElectOne();
    const unsigned tabHi = smemU32(tables), tabLo = tabHi + G::numTables * P.tablePitch;
    const unsigned long long descHigh =
        ((unsigned long long)((128u >> 4) | (1u << 14)) << 32) | ((unsigned long long)((P.tablePitch >> 4) & 0x3FFFu) << 16);
    unsigned g = 0, it = 0;
    for (unsigned tile = blockIdx.x; tile < P.totalTiles; tile += gridDim.x, it++) {
      if (it > 0) {
        mbarWait(&dEmpty, (it - 1u) & 1u);  // the epilogue has read the previous tile's accumulator
        tcFenceAfter();
      }
      for (unsigned st = 0; st < numStages; st++, g++) {
        const unsigned slot = g % kTcRing, n = g / kTcRing;
        mbarWait(&aFull[slot], n & 1u);
        tcFenceAfter();
        if (leader) {
#pragma unroll
          for (unsigned half = 0; half < 2; half++) {
            const unsigned j = 2u * st + half;           // k-step: k = 8j .. 8j+7
            const unsigned a = (8u * j) / D, tb0 = ((8u * j) % D) / 4u;
            const unsigned off = tb0 * P.tablePitch + 16u * (P.aMax - a);
            const unsigned aCols = tmem + colA + 16u * slot + 8u * half;
            tcMma(tmem + colD, aCols, descHigh | (((tabHi + off) >> 4) & 0x3FFFu), kTcIdesc, j > 0 ? 1u : 0u);
            tcMma(tmem + colD, aCols, descHigh | (((tabLo + off) >> 4) & 0x3FFFu), kTcIdesc, 1u);
          }
          tcCommit(&aEmpty[slot]);
          if (st + 1 == numStages) tcCommit(&dFull);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== copy warp: 33 segments per tile =====================
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < P.totalTiles; tile += gridDim.x, it++) {
      const unsigned chan = tile / P.tilesPerChannel;
      const unsigned tl = tile - chan * P.tilesPerChannel;
      const float2* src = P.x + (size_t)chan * P.xStride;
      const unsigned long long s0 = (unsigned long long)tl * kTcTileOut * D;  // first sample of the tile
      if (it > 0) mbarWait(&rawEmpty, (it - 1u) & 1u);  // every producer has left the previous tile's samples
      // segments inside the caller-guaranteed extent: one bulk copy each; the others: guarded loads, zero fill
      unsigned fast = 0;
      for (unsigned sg = 0; sg < G::numSegs; sg++) {
        const unsigned long long b = s0 + (unsigned long long)sg * G::SD;
        if (b + G::SD <= P.nIn) {
          fast++;
        } else {
          float2* dst = reinterpret_cast<float2*>(raw + sg * G::segPitch);
          for (unsigned i = lane; i < G::SD; i += 32) {
            dst[i] = (b + i < P.nIn) ? __ldg(src + b + i) : make_float2(0.0f, 0.0f);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&rawFull, fast * G::segBytes);  // also the one arrival the barrier waits for
        for (unsigned sg = 0; sg < fast; sg++) {     // in-bounds segments form a prefix
          bulkLoad1d(raw + sg * G::segPitch, src + s0 + (size_t)sg * G::SD, G::segBytes, &rawFull);
        }
      }
      __syncwarp();
    }
  }
  tcFenceBefore();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
  }
}

}  // namespace gsdr_b200
