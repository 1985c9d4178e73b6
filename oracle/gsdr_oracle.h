/*
 * gsdr_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the arithmetic performed by kernrj/gsdr's FIR and
 * adjustFrequency CUDA kernels.  It exists so the sm_100a kernels in
 * gsdr_b200/csrc can be checked; nothing in the product path may include, link
 * or call it (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl cpu legs do).
 *
 * PARITY STATUS: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md §8c), so the oracle is pinned instead against outputs of the
 * reference's CUDA kernels themselves (oracle/_ref/libgsdr_ref.so, built by
 * oracle/build_ref.sh from /root/reference/src/fir.cu) captured on a B200 and
 * committed under tests/golden/ (see tests/golden/README.md).
 *
 * Every function cites the reference lines it follows as "ref: path:line"
 * (paths relative to /root/reference).
 */
#ifndef GSDR_ORACLE_H_
#define GSDR_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-compatible with cuComplex (float2: x = real, y = imaginary). */
typedef struct {
  float re;
  float im;
} oracle_c32;

/* ---- decimating FIR, order-faithful float arithmetic (ref: src/fir.cu:49-71) ---- */
void gsdr_oracle_fir_ff(size_t decimation, const float* taps, size_t tapCount, const float* input, float* output,
                        size_t numOutputs);
void gsdr_oracle_fir_fc(size_t decimation, const float* taps, size_t tapCount, const oracle_c32* input,
                        oracle_c32* output, size_t numOutputs);
void gsdr_oracle_fir_cc(size_t decimation, const oracle_c32* taps, size_t tapCount, const oracle_c32* input,
                        oracle_c32* output, size_t numOutputs);
void gsdr_oracle_fir_cf(size_t decimation, const oracle_c32* taps, size_t tapCount, const float* input,
                        oracle_c32* output, size_t numOutputs);

/* ---- same sums accumulated in double: the "truth" the float tolerance is measured from ---- */
void gsdr_oracle_fir_ff_f64(size_t decimation, const float* taps, size_t tapCount, const float* input, double* output,
                            size_t numOutputs);
/* output: interleaved (re, im) doubles */
void gsdr_oracle_fir_fc_f64(size_t decimation, const float* taps, size_t tapCount, const oracle_c32* input,
                            double* output, size_t numOutputs);
void gsdr_oracle_fir_cc_f64(size_t decimation, const oracle_c32* taps, size_t tapCount, const oracle_c32* input,
                            double* output, size_t numOutputs);
void gsdr_oracle_fir_cf_f64(size_t decimation, const oracle_c32* taps, size_t tapCount, const float* input,
                            double* output, size_t numOutputs);

/* ---- multi-threaded wrappers (pthreads over contiguous output blocks); same bits as the scalar ones ---- */
void gsdr_oracle_fir_fc_mt(size_t decimation, const float* taps, size_t tapCount, const oracle_c32* input,
                           oracle_c32* output, size_t numOutputs, int numThreads);
void gsdr_oracle_fir_ff_mt(size_t decimation, const float* taps, size_t tapCount, const float* input, float* output,
                           size_t numOutputs, int numThreads);

/* ---- NCO (adjustFrequency) ---- */
enum {
  GSDR_ORACLE_NCO_LITERAL = 0, /* the reference's per-tap phase arithmetic, bugs included */
  GSDR_ORACLE_NCO_EXACT = 1    /* e^{j 2 pi f n / fs} with a 64-bit fixed-point phase accumulator */
};

/* Host-side pre-reduction of the sample index done by the reference's callers (ref: src/fm.cu:202, src/am.cu:67). */
uint32_t gsdr_oracle_reduce_first_sample_index(size_t firstSampleIndex, float sampleRate);

/* thetaDivPi for one sample (ref: src/adjustFrequency.cu:23,35-43). */
float gsdr_oracle_nco_literal_theta_div_pi(float frequencyShift, uint32_t sampleIndex, float sampleRate);

/* 64-bit phase increment (cycles * 2^64 per sample) used by the EXACT mode. */
uint64_t gsdr_oracle_nco_exact_phase_step(float frequencyShift, float sampleRate);
/* Phase of absolute sample n in EXACT mode, as the signed 32-bit fraction the kernels feed to sincospi
 * (thetaDivPi = value * 2^-31). */
int32_t gsdr_oracle_nco_exact_phase_q31(uint64_t phaseStep, uint64_t sampleIndex);

/* Fused mix + decimating FIR, one complex output per D inputs
 * (ref: src/adjustFrequency.cu:25-56 called as in src/fm.cu:46-56). */
void gsdr_oracle_adjust_frequency_fir_fc(int ncoMode, float sampleRate, float frequencyShift, size_t firstSampleIndex,
                                         size_t decimation, const float* taps, size_t tapCount,
                                         const oracle_c32* input, oracle_c32* output, size_t numOutputs);
/* double-accumulate truth for the same (phasor evaluated in double from the same phase argument) */
void gsdr_oracle_adjust_frequency_fir_fc_f64(int ncoMode, float sampleRate, float frequencyShift,
                                             size_t firstSampleIndex, size_t decimation, const float* taps,
                                             size_t tapCount, const oracle_c32* input, double* output,
                                             size_t numOutputs);

/* ---- quadrature demod (ref: src/quad_demod.cu:23-37) ---- */
void gsdr_oracle_quad_fm_demod(const oracle_c32* input, float* output, float gain, size_t numOutputs);

#ifdef __cplusplus
}
#endif
#endif /* GSDR_ORACLE_H_ */
