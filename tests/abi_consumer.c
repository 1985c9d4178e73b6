/* abi_consumer.c — a plain-C user of <gsdr/fir.h>, written the way a caller of kernrj/gsdr writes it
 * (ref: include/gsdr/fir.h:30-38: decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream;
 * all pointers are device memory).  Built by tests/test_abi.py with gcc and linked against libgsdr_b200.so.
 * Exit codes: 0 = ran on a GPU and the impulse response came back, 3 = no usable CUDA device (the call still has to
 * return a cudaError_t), anything else = failure. */
#include <gsdr/fir.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x)                                                                      \
  do {                                                                                \
    cudaError_t st_ = (x);                                                            \
    if (st_ != cudaSuccess) {                                                         \
      printf("%s -> %d (%s)\n", #x, (int)st_, cudaGetErrorName(st_));                 \
      return st_ == cudaErrorNoDevice || st_ == cudaErrorInsufficientDriver ? 3 : 1;  \
    }                                                                                 \
  } while (0)

int main(void) {
  enum { D = 8, T = 255, NOUT = 1000, NIN = (NOUT - 1) * D + T };
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    /* no device: the entry point must still come back with an error code */
    cudaError_t st = gsdrFirFC(D, NULL, T, NULL, NULL, NOUT, 0, NULL);
    printf("gsdrFirFC without a device -> %d\n", (int)st);
    return st != cudaSuccess ? 3 : 1;
  }
  float* taps = (float*)malloc(sizeof(float) * T);
  cuComplex* x = (cuComplex*)calloc(NIN, sizeof(cuComplex));
  cuComplex* y = (cuComplex*)malloc(sizeof(cuComplex) * NOUT);
  for (int i = 0; i < T; i++) taps[i] = (float)(i + 1);
  x[5 * D + 7].x = 1.0f; /* an impulse at input sample 47: out[n] = taps[47 - 8 n] for n = 0..5 */
  x[5 * D + 7].y = -2.0f;
  float* dTaps;
  cuComplex *dX, *dY;
  cudaStream_t stream;
  CHECK(cudaMalloc((void**)&dTaps, sizeof(float) * T));
  CHECK(cudaMalloc((void**)&dX, sizeof(cuComplex) * NIN));
  CHECK(cudaMalloc((void**)&dY, sizeof(cuComplex) * NOUT));
  CHECK(cudaStreamCreate(&stream));
  CHECK(cudaMemcpyAsync(dTaps, taps, sizeof(float) * T, cudaMemcpyHostToDevice, stream));
  CHECK(cudaMemcpyAsync(dX, x, sizeof(cuComplex) * NIN, cudaMemcpyHostToDevice, stream));
  CHECK(gsdrFirFC(D, dTaps, T, dX, dY, NOUT, 0, stream));
  CHECK(cudaMemcpyAsync(y, dY, sizeof(cuComplex) * NOUT, cudaMemcpyDeviceToHost, stream));
  CHECK(cudaStreamSynchronize(stream));
  for (int n = 0; n < NOUT; n++) {
    const int k = 47 - D * n;
    const float want = (k >= 0 && k < T) ? taps[k] : 0.0f;
    if (y[n].x != want || y[n].y != -2.0f * want) {
      printf("gsdrFirFC: out[%d] = (%g, %g), expected (%g, %g)\n", n, y[n].x, y[n].y, want, -2.0f * want);
      return 1;
    }
  }
  printf("gsdrFirFC impulse response ok (%d outputs)\n", NOUT);
  return 0;
}
