/*
 * gsdr_oracle.c — TEST INFRASTRUCTURE ONLY (see gsdr_oracle.h).
 *
 * Scalar C restatement of what the reference's CUDA kernels compute, following
 * the floating-point expression shapes nvcc 12.9 emits for them at -O3 for
 * sm_100 (checked with cuobjdump, see DESIGN.md "Oracle"):
 *   - FF/FC/CF: one fused multiply-add per component per tap, taps ascending,
 *     single accumulator starting at +0      (ref: src/fir.cu:64-70,
 *     src/cuComplexOperatorOverloads.cuh:29-33,57-62)
 *   - CC: cuCmulf contracted to FMUL+FFMA, then a separate FADD into the
 *     accumulator                             (ref: src/fir.cu:64-70,
 *     src/cuComplexOperatorOverloads.cuh:25-27,57-62)
 *
 * Build with -ffp-contract=off: every fmaf below is deliberate and every
 * a*b+c that is NOT written as fmaf must stay two roundings.
 */
#include "gsdr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* FIR: out[n] = sum_{i<T} in[n*D + i] * taps[i]   (correlation; the caller     */
/* pre-reverses taps — the parameter is called tapsReversed, ref: src/fir.cu:52) */
/* ------------------------------------------------------------------------- */

void gsdr_oracle_fir_ff(size_t D, const float* taps, size_t T, const float* in, float* out, size_t nOut) {
  for (size_t n = 0; n < nOut; n++) { /* one CUDA thread per n, ref: src/fir.cu:57-62 */
    const float* x = in + n * D;      /* ref: src/fir.cu:58,65 */
    float acc = 0.0f;                 /* zero<float>(), ref: src/cuComplexOperatorOverloads.cuh:64-67 */
    for (size_t i = 0; i < T; i++) {  /* ref: src/fir.cu:68-70 */
      acc = fmaf(x[i], taps[i], acc);
    }
    out[n] = acc;
  }
}

void gsdr_oracle_fir_fc(size_t D, const float* taps, size_t T, const oracle_c32* in, oracle_c32* out, size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const oracle_c32* x = in + n * D;
    float re = 0.0f, im = 0.0f; /* zero<cuComplex>(), ref: src/cuComplexOperatorOverloads.cuh:69-72 */
    for (size_t i = 0; i < T; i++) {
      /* (c * r) then +=, ref: src/cuComplexOperatorOverloads.cuh:29-31,57-62 — contracted per component */
      re = fmaf(x[i].re, taps[i], re);
      im = fmaf(x[i].im, taps[i], im);
    }
    out[n].re = re;
    out[n].im = im;
  }
}

void gsdr_oracle_fir_cf(size_t D, const oracle_c32* taps, size_t T, const float* in, oracle_c32* out, size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const float* x = in + n * D;
    float re = 0.0f, im = 0.0f;
    for (size_t i = 0; i < T; i++) {
      /* (r * c) == (c * r), ref: src/cuComplexOperatorOverloads.cuh:33 */
      re = fmaf(taps[i].re, x[i], re);
      im = fmaf(taps[i].im, x[i], im);
    }
    out[n].re = re;
    out[n].im = im;
  }
}

void gsdr_oracle_fir_cc(size_t D, const oracle_c32* taps, size_t T, const oracle_c32* in, oracle_c32* out,
                        size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const oracle_c32* x = in + n * D;
    float re = 0.0f, im = 0.0f;
    for (size_t i = 0; i < T; i++) {
      /* cuCmulf(x, h) = (x.re*h.re - x.im*h.im, x.re*h.im + x.im*h.re), ref:
       * src/cuComplexOperatorOverloads.cuh:25-27.  nvcc 12.9 SASS for k_Fir/k_FirDecimate<float2,float2,float2>:
       *   FMUL t1 = x.im*h.im ; FMUL t2 = x.re*h.im ; FFMA p.re = x.re*h.re - t1 ; FFMA p.im = x.im*h.re + t2 ;
       *   FADD acc.re += p.re ; FADD acc.im += p.im */
      const float t1 = x[i].im * taps[i].im;
      const float t2 = x[i].re * taps[i].im;
      const float pre = fmaf(x[i].re, taps[i].re, -t1);
      const float pim = fmaf(x[i].im, taps[i].re, t2);
      re = re + pre;
      im = im + pim;
    }
    out[n].re = re;
    out[n].im = im;
  }
}

/* ---- double-accumulate "truth" --------------------------------------------------------------- */

void gsdr_oracle_fir_ff_f64(size_t D, const float* taps, size_t T, const float* in, double* out, size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const float* x = in + n * D;
    double acc = 0.0;
    for (size_t i = 0; i < T; i++) acc += (double)x[i] * (double)taps[i];
    out[n] = acc;
  }
}

void gsdr_oracle_fir_fc_f64(size_t D, const float* taps, size_t T, const oracle_c32* in, double* out, size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const oracle_c32* x = in + n * D;
    double re = 0.0, im = 0.0;
    for (size_t i = 0; i < T; i++) {
      re += (double)x[i].re * (double)taps[i];
      im += (double)x[i].im * (double)taps[i];
    }
    out[2 * n] = re;
    out[2 * n + 1] = im;
  }
}

void gsdr_oracle_fir_cf_f64(size_t D, const oracle_c32* taps, size_t T, const float* in, double* out, size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const float* x = in + n * D;
    double re = 0.0, im = 0.0;
    for (size_t i = 0; i < T; i++) {
      re += (double)x[i] * (double)taps[i].re;
      im += (double)x[i] * (double)taps[i].im;
    }
    out[2 * n] = re;
    out[2 * n + 1] = im;
  }
}

void gsdr_oracle_fir_cc_f64(size_t D, const oracle_c32* taps, size_t T, const oracle_c32* in, double* out,
                            size_t nOut) {
  for (size_t n = 0; n < nOut; n++) {
    const oracle_c32* x = in + n * D;
    double re = 0.0, im = 0.0;
    for (size_t i = 0; i < T; i++) {
      const double xr = x[i].re, xi = x[i].im, hr = taps[i].re, hi = taps[i].im;
      re += xr * hr - xi * hi;
      im += xr * hi + xi * hr;
    }
    out[2 * n] = re;
    out[2 * n + 1] = im;
  }
}

/* ---- threaded wrappers ----------------------------------------------------------------------- */

typedef struct {
  int kind; /* 0 = fc, 1 = ff */
  size_t D, T, first, count;
  const float* taps;
  const void* in;
  void* out;
} fir_job;

static void* fir_job_run(void* p) {
  fir_job* j = (fir_job*)p;
  if (j->count == 0) return NULL;
  if (j->kind == 0) {
    gsdr_oracle_fir_fc(j->D, j->taps, j->T, (const oracle_c32*)j->in + j->first * j->D, (oracle_c32*)j->out + j->first,
                       j->count);
  } else {
    gsdr_oracle_fir_ff(j->D, j->taps, j->T, (const float*)j->in + j->first * j->D, (float*)j->out + j->first,
                       j->count);
  }
  return NULL;
}

static void fir_mt(int kind, size_t D, const float* taps, size_t T, const void* in, void* out, size_t nOut,
                   int numThreads) {
  if (numThreads < 1) numThreads = 1;
  fir_job* jobs = (fir_job*)calloc((size_t)numThreads, sizeof(fir_job));
  pthread_t* tids = (pthread_t*)calloc((size_t)numThreads, sizeof(pthread_t));
  for (int t = 0; t < numThreads; t++) {
    const size_t a = (size_t)(((unsigned __int128)nOut * (unsigned)t) / (unsigned)numThreads);
    const size_t b = (size_t)(((unsigned __int128)nOut * (unsigned)(t + 1)) / (unsigned)numThreads);
    jobs[t] = (fir_job){kind, D, T, a, b - a, taps, in, out};
    if (t + 1 < numThreads) pthread_create(&tids[t], NULL, fir_job_run, &jobs[t]);
  }
  fir_job_run(&jobs[numThreads - 1]);
  for (int t = 0; t + 1 < numThreads; t++) pthread_join(tids[t], NULL);
  free(jobs);
  free(tids);
}

void gsdr_oracle_fir_fc_mt(size_t D, const float* taps, size_t T, const oracle_c32* in, oracle_c32* out, size_t nOut,
                           int numThreads) {
  fir_mt(0, D, taps, T, in, out, nOut, numThreads);
}

void gsdr_oracle_fir_ff_mt(size_t D, const float* taps, size_t T, const float* in, float* out, size_t nOut,
                           int numThreads) {
  fir_mt(1, D, taps, T, in, out, nOut, numThreads);
}

/* ------------------------------------------------------------------------- */
/* NCO                                                                        */
/* ------------------------------------------------------------------------- */

uint32_t gsdr_oracle_reduce_first_sample_index(size_t firstSampleIndex, float sampleRate) {
  /* ref: src/fm.cu:202 — (uint32_t)fmodf((float)firstSampleIndex, rfSampleRate) */
  return (uint32_t)fmodf((float)firstSampleIndex, sampleRate);
}

float gsdr_oracle_nco_literal_theta_div_pi(float frequencyShift, uint32_t sampleIndex, float sampleRate) {
  /* ref: src/adjustFrequency.cu:35 — __frcp_rn(f): correctly rounded reciprocal == IEEE 1.0f / f */
  const float period = 1.0f / frequencyShift;
  /* ref: src/adjustFrequency.cu:23,37 — modNorm: fmodf(__uint2float_rn(n), maxVal) / maxVal */
  const float timeSeconds = fmodf((float)sampleIndex, sampleRate) / sampleRate;
  /* ref: src/adjustFrequency.cu:40 — named thetaDiv2Pi there, but no division by the period is done */
  const float thetaDiv2Pi = fmodf(timeSeconds, period);
  /* ref: src/adjustFrequency.cu:43 — scalbnf(x, 1) */
  return thetaDiv2Pi * 2.0f;
}

uint64_t gsdr_oracle_nco_exact_phase_step(float frequencyShift, float sampleRate) {
  /* cycles per sample, wrapped to [-0.5, 0.5), times 2^64, round to nearest. */
  const double r = (double)frequencyShift / (double)sampleRate;
  const double frac = r - floor(r + 0.5); /* [-0.5, 0.5) */
  /* frac * 2^64 is an exact power-of-two scaling; its magnitude is <= 2^63 - 2^10, so llrint cannot overflow
   * (frac == -0.5 maps to INT64_MIN == 2^63 as a uint64, i.e. half a cycle per sample). */
  return (uint64_t)(int64_t)llrint(frac * 18446744073709551616.0);
}

int32_t gsdr_oracle_nco_exact_phase_q31(uint64_t phaseStep, uint64_t sampleIndex) {
  const uint64_t phase = phaseStep * sampleIndex; /* mod 2^64: exact, associative, shardable */
  return (int32_t)(uint32_t)(phase >> 32);
}

static void nco_phasor_f32(int mode, float f, float fs, uint32_t idx32, uint64_t step, uint64_t idx64, float* c,
                           float* s) {
  double v;
  if (mode == GSDR_ORACLE_NCO_LITERAL) {
    v = (double)gsdr_oracle_nco_literal_theta_div_pi(f, idx32, fs);
  } else {
    v = (double)((float)gsdr_oracle_nco_exact_phase_q31(step, idx64) * 4.656612873077392578125e-10f); /* 2^-31 */
  }
  /* ref: src/adjustFrequency.cu:50 — sincospif(v): sin(pi v), cos(pi v); CUDA's is <= 2 ulp, use the
   * correctly rounded double result here. */
  *s = (float)sin(M_PI * v);
  *c = (float)cos(M_PI * v);
}

void gsdr_oracle_adjust_frequency_fir_fc(int mode, float fs, float f, size_t firstSampleIndex, size_t D,
                                         const float* taps, size_t T, const oracle_c32* in, oracle_c32* out,
                                         size_t nOut) {
  const uint32_t first32 = gsdr_oracle_reduce_first_sample_index(firstSampleIndex, fs);
  const uint64_t step = gsdr_oracle_nco_exact_phase_step(f, fs);
  for (size_t n = 0; n < nOut; n++) {
    /* ref: src/fm.cu:43-56 — initialInputIndex = decimation*outputIndex; first + initialInputIndex (uint32 wrap) */
    uint32_t idx32 = first32 + (uint32_t)(D * n);
    const uint64_t idx64 = (uint64_t)firstSampleIndex + (uint64_t)D * n;
    const oracle_c32* x = in + n * D;
    float re = 0.0f, im = 0.0f;
    for (size_t i = 0; i < T; i++, idx32++) { /* ref: src/adjustFrequency.cu:36 */
      float c, s;
      nco_phasor_f32(mode, f, fs, idx32, step, idx64 + i, &c, &s);
      /* ref: src/adjustFrequency.cu:51 — inVal * cosVal via cuCmulf; nvcc 12.9 SASS:
       *   FMUL a = in.im*sin ; FFMA m.re = in.re*cos - a ; FMUL b = in.im*cos ; FFMA m.im = in.re*sin + b */
      const float a = x[i].im * s;
      const float mre = fmaf(x[i].re, c, -a);
      const float b = x[i].im * c;
      const float mim = fmaf(x[i].re, s, b);
      /* ref: src/adjustFrequency.cu:52-54 — (m * tap) then += : FFMA acc = tap*m + acc */
      re = fmaf(taps[i], mre, re);
      im = fmaf(taps[i], mim, im);
    }
    out[n].re = re;
    out[n].im = im;
  }
}

void gsdr_oracle_adjust_frequency_fir_fc_f64(int mode, float fs, float f, size_t firstSampleIndex, size_t D,
                                             const float* taps, size_t T, const oracle_c32* in, double* out,
                                             size_t nOut) {
  const uint32_t first32 = gsdr_oracle_reduce_first_sample_index(firstSampleIndex, fs);
  const uint64_t step = gsdr_oracle_nco_exact_phase_step(f, fs);
  for (size_t n = 0; n < nOut; n++) {
    uint32_t idx32 = first32 + (uint32_t)(D * n);
    const uint64_t idx64 = (uint64_t)firstSampleIndex + (uint64_t)D * n;
    const oracle_c32* x = in + n * D;
    double re = 0.0, im = 0.0;
    for (size_t i = 0; i < T; i++, idx32++) {
      double v;
      if (mode == GSDR_ORACLE_NCO_LITERAL) {
        v = (double)gsdr_oracle_nco_literal_theta_div_pi(f, idx32, fs);
      } else {
        v = (double)((float)gsdr_oracle_nco_exact_phase_q31(step, idx64 + i) * 4.656612873077392578125e-10f);
      }
      const double c = cos(M_PI * v), s = sin(M_PI * v);
      const double xr = x[i].re, xi = x[i].im;
      re += (xr * c - xi * s) * (double)taps[i];
      im += (xr * s + xi * c) * (double)taps[i];
    }
    out[2 * n] = re;
    out[2 * n + 1] = im;
  }
}

/* ------------------------------------------------------------------------- */
/* Quadrature FM demod                                                         */
/* ------------------------------------------------------------------------- */

void gsdr_oracle_quad_fm_demod(const oracle_c32* in, float* out, float gain, size_t nOut) {
  for (size_t i = 0; i < nOut; i++) {
    /* ref: src/quad_demod.cu:30 — m = input[i+1] * conj(input[i]) via cuCmulf(x = in[i+1], y = conj(in[i])) */
    const float xr = in[i + 1].re, xi = in[i + 1].im;
    const float yr = in[i].re, yi = -in[i].im;
    const float t1 = xi * yi;
    const float t2 = xr * yi;
    const float mre = fmaf(xr, yr, -t1);
    const float mim = fmaf(xi, yr, t2);
    /* ref: src/quad_demod.cu:31 */
    out[i] = gain * atan2f(mim, mre);
  }
}
