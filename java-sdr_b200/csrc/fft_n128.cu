// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(128, 64, 8, 8, 16, 1, 1)
