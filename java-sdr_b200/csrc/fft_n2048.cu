// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(2048, 128, 1, 16, 16, 8, 1)
