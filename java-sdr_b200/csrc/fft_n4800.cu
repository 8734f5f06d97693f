// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(4800, 320, 1, 20, 16, 15, 1)
