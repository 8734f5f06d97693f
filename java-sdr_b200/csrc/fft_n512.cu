// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(512, 256, 8, 32, 16, 1, 1)
