// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(4096, 64, 1, 64, 64, 1, 1)
