// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(128, 256, 16, 16, 8, 1, 1)
