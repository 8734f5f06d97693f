// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(19200, 640, 1, 16, 16, 15, 5)
