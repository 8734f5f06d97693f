// fft_generic.cuh — batched Stockham auto-sort FFT through global memory, one launch
// per radix stage, for any length N = 2^a 3^b 5^c 7^d, in binary32 or binary64.
//
// Two users:
//   * the FUNcube auto-tune path (FUNcubeBPSKDemod.java:406-464) transforms in
//     binary64 (JTransforms DoubleFFT_1D.complexForward / complexInverse(.., true),
//     :422-423,459); N = rate/10 does not fit a CTA's shared memory in binary64;
//   * fft.java lengths without a compiled single-CTA plan (fft_plans.h), e.g. the top
//     of BASELINE config 3's sweep (32768, 65536).
// Every stage reads and writes the whole batch (L2 / HBM), so this is the slow,
// general path: 2*16 (or 2*8) bytes per sample per stage.
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace jsdr {
namespace fftg {

template <typename T> struct Cx;
template <> struct Cx<float> { typedef float2 type; };
template <> struct Cx<double> { typedef double2 type; };

struct StagePlan {
    int nstages;
    int radix[24];
};

// radices of N (4 first, then 2, 3, 5, 7); nstages = 0 if N has another prime factor
StagePlan make_plan(int n);

// Transform `batch` blocks of n complex values.  `a` holds the input and is used as
// ping-pong scratch together with `b`; returns the buffer that holds the result
// (natural order).  sign = -1 forward (e^{-i..}), +1 inverse (unscaled).
template <typename T>
typename Cx<T>::type *run(jsdr_ctx *ctx, cudaStream_t st, typename Cx<T>::type *a, typename Cx<T>::type *b, int n,
                         long long batch, int sign, int *rc);

}  // namespace fftg
}  // namespace jsdr
