// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(4096, 256, 1, 16, 16, 16, 1)
