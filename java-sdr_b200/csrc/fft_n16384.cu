// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(16384, 512, 1, 32, 32, 16, 1)
JSDR_FFT_DEFINE_SPLIT(16384, 512, 1, 32, 32, 16, 1)
