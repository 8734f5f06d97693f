// generated per-length instantiation (see fft_plans.h)
#include "fft_inst.cuh"
JSDR_FFT_DEFINE(9600, 480, 1, 20, 20, 24, 1)
