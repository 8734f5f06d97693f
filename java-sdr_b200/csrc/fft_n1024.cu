// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(1024, 128, 4, 32, 32, 1, 1)
