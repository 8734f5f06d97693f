// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(4410, 224, 1, 10, 21, 21, 1)
