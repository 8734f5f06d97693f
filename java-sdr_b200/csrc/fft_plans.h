// fft_plans.h — the table of FFT lengths with a compiled plan.
// X(N, T, G, R0, R1, R2, R3): N = R0*R1*R2*R3, T threads per CTA, G blocks per CTA.
// Power-of-two radices go first so that butterfly strides stay 16-aligned; the
// odd radices of N = rate/10 (fft.java:67, JavaAudio.java:59) come last.
#pragma once
#define JSDR_FFT_PLANS(X)            \
    X(128, 64, 8, 8, 16, 1, 1)     \
    X(256, 128, 8, 16, 16, 1, 1)    \
    X(512, 64, 4, 16, 32, 1, 1)     \
    X(1024, 128, 4, 32, 32, 1, 1)    \
    X(2048, 128, 1, 16, 16, 8, 1)    \
    X(4096, 64, 1, 64, 64, 1, 1)    \
    X(8192, 256, 1, 32, 16, 16, 1)   \
    X(16384, 512, 1, 32, 32, 16, 1)  \
    X(4410, 224, 1, 10, 21, 21, 1)   \
    X(4800, 320, 1, 20, 16, 15, 1)   \
    X(9600, 320, 1, 32, 20, 15, 1)   \
    X(19200, 640, 1, 16, 16, 15, 5)
