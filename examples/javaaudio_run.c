/*
 * javaaudio_run.c — a plain C99 client of libjsdrcuda.so that does what JavaAudio.run does
 * (JavaAudio.java:224-304): read one block of s16le IQ at a time into a buffer it reuses, hand
 * the raw bytes to the handlers, print what they publish.  Nothing but include/jsdrcuda.h and
 * libc: this is the shape of the calls a java.lang.foreign or JNI binding makes
 * (INTEGRATION.md), and tests/test_c_client.py builds it with `gcc -std=c99 -pedantic -Werror`
 * to keep the header honest C.
 *
 *   cc -std=c99 -I include examples/javaaudio_run.c -L java-sdr_b200 -ljsdrcuda -o javaaudio_run
 *   LD_LIBRARY_PATH=java-sdr_b200 ./javaaudio_run tests/golden/sine4410.raw 44100 4096
 *
 * Arguments: file (s16le interleaved I,Q), sample rate, complex samples per block
 * (blen / size of the AudioDescriptor; default rate/10 as JavaAudio.java:59), then optional
 * tuner frequencies in Hz (default 12000, FUNcubeBPSKDemod.java:195).  One line per block:
 *   block <k> peak_hz <int> peak_db <float> peak_bin <int> ds <outputs per tuner> bits <n0> <n1> ...
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jsdrcuda.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != JSDR_OK) {                                                        \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, jsdr_last_error()); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s file.raw rate [samples_per_block [tuning_hz ...]]\n", argv[0]);
        return 2;
    }
    const int rate = atoi(argv[2]);
    const int n = argc > 3 ? atoi(argv[3]) : rate / 10;
    double tuning[16];
    int ntuners = 0;
    for (int a = 4; a < argc && ntuners < 16; a++) tuning[ntuners++] = atof(argv[a]);
    if (ntuners == 0) tuning[ntuners++] = 12000.0;
    if (rate <= 0 || n <= 0) {
        fprintf(stderr, "bad rate or block length\n");
        return 2;
    }
    FILE *in = fopen(argv[1], "rb");
    if (!in) {
        perror(argv[1]);
        return 2;
    }
    if (jsdr_abi_version() != JSDR_ABI_VERSION) {
        fprintf(stderr, "libjsdrcuda.so has ABI %d, header %d\n", jsdr_abi_version(), JSDR_ABI_VERSION);
        return 1;
    }

    jsdr_ctx *ctx = NULL;
    jsdr_fft *fft = NULL;
    jsdr_bpsk *bank = NULL;
    CHECK(jsdr_ctx_create(0, &ctx));
    CHECK(jsdr_fft_create(ctx, n, rate, 1, &fft));                       /* fft.setup, fft.java:63-77 */
    /* the tuners share the one stream (jsdr.java:479-483); rates below 9600 have no decimator */
    if (rate >= 9600 && jsdr_bpsk_create(ctx, rate, ntuners, tuning, n, &bank) != JSDR_OK) {
        fprintf(stderr, "no tuner bank: %s\n", jsdr_last_error());
        bank = NULL;
    }

    /* pinned buffers, reused every block like JavaAudio's `buf` (JavaAudio.java:224) */
    void *raw = NULL, *psd = NULL, *peak = NULL;
    CHECK(jsdr_host_alloc(ctx, (size_t)n * 4, &raw));
    CHECK(jsdr_host_alloc(ctx, ((size_t)n + 2) * sizeof(float), &psd));
    CHECK(jsdr_host_alloc(ctx, sizeof(int32_t), &peak));
    const int max_bits = n / 8 + 16;
    int8_t *bits = malloc((size_t)ntuners * max_bits);
    int64_t *bit_at = malloc((size_t)ntuners * max_bits * sizeof(int64_t));
    int32_t *nbits = calloc((size_t)ntuners, sizeof(int32_t));
    if (!bits || !bit_at || !nbits) return 1;

    long long total_bits = 0;
    int blocks = 0;
    while (fread(raw, 4, (size_t)n, in) == (size_t)n) {                    /* whole blocks only */
        /* IRawHandler.receive(byte[]) of the spectrum handler: fft.java:190-228 on the device */
        CHECK(jsdr_fft_receive_s16(fft, (const int16_t *)raw, 1, 0, 0, (float *)psd, (int32_t *)peak, JSDR_MEM_HOST));
        const float *p = (const float *)psd;
        printf("block %d peak_hz %d peak_db %.4f peak_bin %d", blocks, (int)p[n], (double)p[n + 1], (int)*(int32_t *)peak);
        if (bank) {
            int32_t nds = 0;
            CHECK(jsdr_bpsk_receive_s16(bank, (const int16_t *)raw, n, 0 /* shared stream */, 0, 0, JSDR_MEM_HOST));
            CHECK(jsdr_bpsk_last_counts(bank, &nds));
            CHECK(jsdr_bpsk_read_bits(bank, bits, bit_at, nbits, max_bits, JSDR_MEM_HOST));
            printf(" ds %d bits", (int)nds);
            for (int t = 0; t < ntuners; t++) {
                printf(" %d", (int)nbits[t]);
                total_bits += nbits[t];
            }
        }
        printf("\n");
        blocks++;
    }
    printf("done blocks %d bits %lld\n", blocks, total_bits);

    fclose(in);
    free(bits);
    free(bit_at);
    free(nbits);
    if (bank) CHECK(jsdr_bpsk_destroy(bank));
    CHECK(jsdr_fft_destroy(fft));
    CHECK(jsdr_host_free(ctx, raw));
    CHECK(jsdr_host_free(ctx, psd));
    CHECK(jsdr_host_free(ctx, peak));
    CHECK(jsdr_ctx_destroy(ctx));
    return 0;
}
