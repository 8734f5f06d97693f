// fec.cu — the frame stage behind the FUNcube bit stream:
//   sync correlator   FUNcubeBPSKDemod.java:553-574  (65-point correlation against
//                     SYNC_VECTOR at stride 80 over the last 5200 bits, >= 45 starts a decode)
//   FECDecode         FECDecoder.java:703-852        (de-interleave 80x65, Viterbi K=7 r=1/2,
//                     de-scramble, 2 x RS(160,128) = CCSDS (255,223) shortened by 95,
//                     re-encode and count channel errors)
// All integer work, bit-exact.  The reference shifts a 5200-byte array per bit (:553); here
// the last 5199 bits of every channel stay on the device and every new bit's correlation
// is an independent thread.  A detected frame is decoded by one warp: the 64-state
// add-compare-select runs one butterfly per lane with the path metrics exchanged by
// shuffles, the decisions of a step are two ballots, lanes 0/1 run the two RS decoders.
//
// Tables: GF(256), scrambler, convolutional-code symbols, RS generator and the sync
// vector are generated from their defining polynomials (FECDecoder.java:40-57,105-181,
// 544-546, 600-605).  The Viterbi metric table (FECDecoder.java:67-100) is a literal in
// the reference, not a formula: it is an INPUT of jsdr_bpsk_enable_fec (the Java shim
// passes the reference's own array).
#include <string.h>

#include <algorithm>
#include <vector>

#include "handles.h"

namespace jsdr {
namespace fec {

constexpr int NN = 255, KK = 223, NROOTS = 32, FCR = 112, PRIM = 11, IPRIM = 116, A0 = 255;
constexpr int RSBLOCKS = 2, RSPAD = 95;
constexpr int NBITS = (256 + NROOTS * RSBLOCKS) * 8 + 6;      // 2566
constexpr int ROWS = 80, COLS = 65;                           // interleaver (FECDecoder "ROWS"=80, "COLUMNS"=65)
constexpr int SYMS = 5200, SYNC_LEN = 65, HIST = SYMS - 1;
constexpr int CPOLYA = 0x4f, CPOLYB = 0x6d, SYNC_POLY = 0x48;

struct Tables {
    int16_t mettab[2][256];
    uint8_t partab[256];
    uint8_t syms[128];
    uint8_t scrambler[320];
    uint8_t alpha_to[256], index_of[256];
    uint8_t rs_poly[16];
    int8_t sync[SYNC_LEN];        // +1 / -1 (FUNcubeBPSKDemod.SYNC_VECTOR, :79-81)
};

__constant__ Tables c_tab;

static int parity8(int x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1; }
__host__ __device__ static inline int mod255(int x)
{
    while (x >= 255) { x -= 255; x = (x >> 8) + (x & 255); }
    return x;
}

static void build_tables(Tables &t, const int16_t *mettab)
{
    memcpy(t.mettab, mettab, sizeof(t.mettab));
    for (int i = 0; i < 256; i++) t.partab[i] = (uint8_t)parity8(i);
    // symbol pair of encoder state s; the second symbol is inverted (FECDecoder.java:105-114, 564)
    for (int s = 0; s < 128; s++) t.syms[s] = (uint8_t)((parity8(s & CPOLYA) << 1) | (1 - parity8(s & CPOLYB)));
    int sr = 1;                                               // GF(256), x^8+x^7+x^2+x+1 (:145-181)
    for (int i = 0; i < 255; i++) {
        t.alpha_to[i] = (uint8_t)sr;
        t.index_of[sr] = (uint8_t)i;
        sr <<= 1;
        if (sr & 0x100) sr ^= 0x187;
    }
    t.alpha_to[255] = 0;
    t.index_of[0] = A0;
    int st = 0xff;                                            // CCSDS randomiser, x^8+x^7+x^5+x^3+1 (:118-139)
    for (int i = 0; i < 320; i++) {
        int byte = 0;
        for (int b = 0; b < 8; b++) {
            byte = (byte << 1) | ((st >> 7) & 1);
            const int fb = ((st >> 7) ^ (st >> 4) ^ (st >> 2) ^ st) & 1;
            st = ((st << 1) | fb) & 0xff;
        }
        t.scrambler[i] = (uint8_t)byte;
    }
    int g[NROOTS + 1];                                        // RS generator, roots alpha^(PRIM*(FCR+i)) (:544-546)
    memset(g, 0, sizeof(g));
    g[0] = 1;
    int root = FCR * PRIM;
    for (int i = 0; i < NROOTS; i++, root += PRIM) {
        g[i + 1] = 1;
        for (int j = i; j > 0; j--) {
            if (g[j] != 0) g[j] = g[j - 1] ^ t.alpha_to[mod255(t.index_of[g[j]] + root)];
            else g[j] = g[j - 1];
        }
        g[0] = t.alpha_to[mod255(t.index_of[g[0]] + root)];
    }
    for (int j = 0; j < 16; j++) t.rs_poly[j] = t.index_of[g[j + 1]];
    int s7 = 0x7f;                                            // sync LFSR (:600-605)
    for (int i = 0; i < SYNC_LEN; i++) {
        t.sync[i] = (s7 & 64) ? 1 : -1;
        s7 = ((s7 << 1) | t.partab[s7 & SYNC_POLY]) & 0xff;
    }
}

// ------------------------------------------------------------------ sync correlator
struct FrameMeta {
    int chan;
    int corr;
    long long bit_index;      // cntBit of the bit that completed the frame
    int start;                // position of the frame's first bit in the channel's (history ++ new bits) sequence
    int errors;               // FECDecode's return: channel errors, or -1
};

// the bit sequence of one channel as the correlator sees it: 5199 old bits, then this call's
__device__ __forceinline__ int seq_bit(const int8_t *__restrict__ hist, const int8_t *__restrict__ bits, int pos)
{
    return pos < HIST ? hist[pos] : bits[pos - HIST];
}

// :556-562 for every new bit of every channel: one thread each
__global__ void __launch_bounds__(256) k_sync(const int8_t *__restrict__ hist, const int8_t *__restrict__ bits,
                                              const int32_t *__restrict__ nbits, int max_bits, int nchan,
                                              const long long *__restrict__ cnt_bit_after, int ts_stride_ll,
                                              FrameMeta *__restrict__ frames, int *__restrict__ nframes, int max_frames,
                                              long long *__restrict__ cnt_fec)
{
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = min(nbits[ch], max_bits);
    if (i >= nb) return;
    const int8_t *h = hist + (size_t)ch * HIST;
    const int8_t *b = bits + (size_t)ch * max_bits;
    int corr = 0;
#pragma unroll 5
    for (int n = 0; n < SYNC_LEN; n++) corr += seq_bit(h, b, i + 80 * n) * c_tab.sync[n];
    if (corr >= 45) {
        atomicAdd((unsigned long long *)&cnt_fec[ch], 1ull);
        const int slot = atomicAdd(nframes, 1);
        if (slot < max_frames) {
            FrameMeta m;
            m.chan = ch;
            m.corr = corr;
            m.bit_index = cnt_bit_after[(size_t)ch * ts_stride_ll] - nb + i;
            m.start = i;
            m.errors = -2;
            frames[slot] = m;
        }
    }
}

// new history = the last 5199 bits of (history ++ new bits)
__global__ void __launch_bounds__(256) k_sync_shift(const int8_t *__restrict__ hist_in, int8_t *__restrict__ hist_out,
                                                    const int8_t *__restrict__ bits, const int32_t *__restrict__ nbits,
                                                    int max_bits)
{
    const int ch = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= HIST) return;
    const int nb = min(nbits[ch], max_bits);
    hist_out[(size_t)ch * HIST + k] = (int8_t)seq_bit(hist_in + (size_t)ch * HIST, bits + (size_t)ch * max_bits, k + nb);
}

// ------------------------------------------------------------------ FECDecode, one warp per frame
struct DecodeSmem {
    uint8_t raw[SYMS];                       // dmFECBits: 0xc0 / 0x40 (:564-566)
    uint8_t symbols[NBITS * 2 + 65 + 3];     // de-interleaved
    unsigned dec[NBITS][2];                  // decisions of states 2b (even) and 2b+1 (odd), bit b
    uint8_t vit[(NBITS - 6) / 8];
    uint8_t rs[RSBLOCKS][NN];
    uint8_t out[256];
    uint8_t reenc[SYMS];
    int rserr[RSBLOCKS];
};

// decode_rs_8 (FECDecoder.java:325-519) with no erasures, one thread
__device__ int decode_rs_8(uint8_t *data)
{
    int lambda[NROOTS + 1], s[NROOTS], b[NROOTS + 1], t[NROOTS + 1], omega[NROOTS + 1];
    int root[NROOTS], reg[NROOTS + 1], loc[NROOTS];
    int deg_lambda, el, deg_omega, count, r, syn_error;
    const uint8_t *AT = c_tab.alpha_to, *IO = c_tab.index_of;
    for (int i = 0; i <= NROOTS; i++) lambda[i] = 0;
    for (int i = 0; i < NROOTS; i++) s[i] = data[0];
    for (int j = 1; j < NN; j++)
        for (int i = 0; i < NROOTS; i++) {
            if (s[i] == 0) s[i] = data[j];
            else s[i] = data[j] ^ AT[mod255(IO[s[i]] + (FCR + i) * PRIM)];
        }
    syn_error = 0;
    for (int i = 0; i < NROOTS; i++) { syn_error |= s[i]; s[i] = IO[s[i]]; }
    if (!syn_error) return 0;
    lambda[0] = 1;
    for (int i = 0; i < NROOTS + 1; i++) b[i] = IO[lambda[i]];
    r = 0; el = 0;
    while (++r <= NROOTS) {
        int discr_r = 0;
        for (int i = 0; i < r; i++)
            if (lambda[i] != 0 && s[r - i - 1] != A0) discr_r ^= AT[mod255(IO[lambda[i]] + s[r - i - 1])];
        discr_r = IO[discr_r];
        if (discr_r == A0) {
            for (int i = NROOTS; i > 0; i--) b[i] = b[i - 1];
            b[0] = A0;
        } else {
            t[0] = lambda[0];
            for (int i = 0; i < NROOTS; i++) {
                if (b[i] != A0) t[i + 1] = lambda[i + 1] ^ AT[mod255(discr_r + b[i])];
                else t[i + 1] = lambda[i + 1];
            }
            if (2 * el <= r - 1) {
                el = r - el;
                for (int i = 0; i <= NROOTS; i++) b[i] = (lambda[i] == 0) ? A0 : mod255(IO[lambda[i]] - discr_r + NN);
            } else {
                for (int i = NROOTS; i > 0; i--) b[i] = b[i - 1];
                b[0] = A0;
            }
            for (int i = 0; i <= NROOTS; i++) lambda[i] = t[i];
        }
    }
    deg_lambda = 0;
    for (int i = 0; i < NROOTS + 1; i++) {
        lambda[i] = IO[lambda[i]];
        if (lambda[i] != A0) deg_lambda = i;
    }
    for (int i = 1; i <= NROOTS; i++) reg[i] = lambda[i];
    count = 0;
    for (int i = 1, k = IPRIM - 1; i <= NN; i++, k = mod255(k + IPRIM)) {
        int q = 1;
        for (int j = deg_lambda; j > 0; j--)
            if (reg[j] != A0) { reg[j] = mod255(reg[j] + j); q ^= AT[reg[j]]; }
        if (q != 0) continue;
        root[count] = i;
        loc[count] = k;
        if (++count == deg_lambda) break;
    }
    if (deg_lambda != count) return -1;
    deg_omega = 0;
    for (int i = 0; i < NROOTS; i++) {
        int tmp = 0;
        int j = (deg_lambda < i) ? deg_lambda : i;
        for (; j >= 0; j--)
            if (s[i - j] != A0 && lambda[j] != A0) tmp ^= AT[mod255(s[i - j] + lambda[j])];
        if (tmp != 0) deg_omega = i;
        omega[i] = IO[tmp];
    }
    omega[NROOTS] = A0;
    for (int j = count - 1; j >= 0; j--) {
        int num1 = 0;
        for (int i = deg_omega; i >= 0; i--)
            if (omega[i] != A0) num1 ^= AT[mod255(omega[i] + i * root[j])];
        const int num2 = AT[mod255(root[j] * (FCR - 1) + NN)];
        int den = 0;
        const int lim = deg_lambda < NROOTS - 1 ? deg_lambda : NROOTS - 1;
        for (int i = lim & ~1; i >= 0; i -= 2)
            if (lambda[i + 1] != A0) den ^= AT[mod255(lambda[i + 1] + i * root[j])];
        if (den == 0) return -1;
        if (num1 != 0) data[loc[j]] ^= AT[mod255(IO[num1] + IO[num2] + NN - IO[den])];
    }
    return count;
}

// encode_FEC40 (FECDecoder.java:527-688) into sym[5200] (0/1), one thread
__device__ void encode_fec40(const uint8_t *data, uint8_t *sym)
{
    const uint8_t *AT = c_tab.alpha_to, *IO = c_tab.index_of, *PT = c_tab.partab;
    int rs_block[RSBLOCKS][NROOTS];
    for (int r = 0; r < RSBLOCKS; r++)
        for (int i = 0; i < NROOTS; i++) rs_block[r][i] = 0;
    int nbytes = 0, bindex = COLS, conv_sr = 0;
    auto put = [&](int c) {                                   // interleave_symbol (:549-556)
        const int col = bindex / COLS, row = bindex % COLS;
        if (c) sym[row * ROWS + col] = 1;
        bindex++;
    };
    auto conv = [&](int c, int cnt) {                         // encode_and_interleave (:558-567)
        while (cnt-- != 0) {
            conv_sr = ((conv_sr << 1) | ((c >> 7) & 1)) & 0xff;
            c = (c << 1) & 0xff;
            put(PT[conv_sr & CPOLYA]);
            put(1 - PT[conv_sr & CPOLYB]);
        }
    };
    for (int i = 0; i < SYMS; i++) sym[i] = 0;
    int sr = 0x7f;                                            // sync column (:600-605)
    for (int i = 0; i < SYNC_LEN; i++) {
        if (sr & 64) sym[ROWS * i] = 1;
        sr = ((sr << 1) | PT[sr & SYNC_POLY]) & 0xff;
    }
    for (int i = 0; i < 256; i++) {                           // :614-655
        const int c = data[i];
        const int rsi = nbytes & 1;
        const int feedback = IO[c ^ rs_block[rsi][0]];
        if (feedback != A0) {
            for (int j = 0; j < 15; j++) {
                const int t = AT[mod255(feedback + c_tab.rs_poly[j])];
                rs_block[rsi][j + 1] ^= t;
                rs_block[rsi][31 - j] ^= t;
            }
            rs_block[rsi][16] ^= AT[mod255(feedback + c_tab.rs_poly[15])];
        }
        for (int k = 0; k < 31; k++) rs_block[rsi][k] = rs_block[rsi][k + 1];
        rs_block[rsi][31] = (feedback != A0) ? AT[feedback] : 0;
        conv(c ^ c_tab.scrambler[nbytes], 8);
        nbytes++;
    }
    for (int i = 0; i < 64; i++) {                            // :662-671
        const int c = rs_block[nbytes & 1][(nbytes - 256) >> 1];
        conv(c ^ c_tab.scrambler[nbytes], 8);
        if (++nbytes == 320) conv(0, 6);
    }
}

__global__ void __launch_bounds__(32) k_fec_decode(const int8_t *__restrict__ hist, const int8_t *__restrict__ bits,
                                                   int max_bits, FrameMeta *__restrict__ frames,
                                                   const int *__restrict__ nframes, int max_frames,
                                                   uint8_t *__restrict__ data_out, long long *__restrict__ cnt_dec)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecodeSmem &S = *reinterpret_cast<DecodeSmem *>(smem_raw);
    const int slot = blockIdx.x, lane = threadIdx.x;
    if (slot >= min(*nframes, max_frames)) return;
    const FrameMeta m = frames[slot];
    const int8_t *h = hist + (size_t)m.chan * HIST;
    const int8_t *b = bits + (size_t)m.chan * max_bits;
    // :564-566 dmFECBits[n] = dmFECCorr[n]==1 ? 0xc0 : 0x40
    for (int n = lane; n < SYMS; n += 32) S.raw[n] = (seq_bit(h, b, m.start + n) == 1) ? 0xc0 : 0x40;
    for (int n = lane; n < (int)sizeof(S.symbols); n += 32) S.symbols[n] = 0;
    __syncwarp();
    // :707-723 de-interleave, skipping the sync column: symbols[(col-1)*65 + row] = raw[row*80 + col]
    for (int n = lane; n < (ROWS - 1) * COLS; n += 32) {
        const int col = 1 + n / COLS, row = n % COLS;
        S.symbols[n] = S.raw[row * ROWS + col];
    }
    __syncwarp();

    // ---- viterbi27 (:203-278).  Lane b owns butterfly b: predecessors b and b+32, successors 2b, 2b+1.
    {
        int lo = (lane == 0) ? 0 : -999999, hi = -999999;     // cmetric[b], cmetric[b+32]
        const int sym_e = c_tab.syms[2 * lane], sym_o = c_tab.syms[2 * lane + 1];
        for (int bit = 0; bit < NBITS; bit++) {
            const int s0 = S.symbols[2 * bit], s1 = S.symbols[2 * bit + 1];
            int mets[4];
#pragma unroll
            for (int i = 0; i < 4; i++) mets[i] = c_tab.mettab[(i >> 1) & 1][s0] + c_tab.mettab[i & 1][s1];
            int b1 = mets[sym_e];
            const int b2 = mets[sym_o];
            int m0 = lo + b1;                                  // nmetric[2b] candidates
            int m1 = hi + b2;
            b1 -= b2;
            const bool d_e = m1 > m0;
            const int n_e = d_e ? m1 : m0;
            m0 -= b1;                                          // nmetric[2b+1] candidates
            m1 += b1;
            const bool d_o = m1 > m0;
            const int n_o = d_o ? m1 : m0;
            const unsigned be = __ballot_sync(0xffffffffu, d_e), bo = __ballot_sync(0xffffffffu, d_o);
            if (lane == 0) {
                S.dec[bit][0] = be;
                S.dec[bit][1] = bo;
            }
            // next step: cmetric[b] = nmetric[b] (lane b>>1, even/odd), cmetric[b+32] = nmetric[b+32] (lane 16 + (b>>1))
            const int src_lo = lane >> 1, src_hi = 16 + (lane >> 1);
            const int e_lo = __shfl_sync(0xffffffffu, n_e, src_lo), o_lo = __shfl_sync(0xffffffffu, n_o, src_lo);
            const int e_hi = __shfl_sync(0xffffffffu, n_e, src_hi), o_hi = __shfl_sync(0xffffffffu, n_o, src_hi);
            lo = (lane & 1) ? o_lo : e_lo;
            hi = (lane & 1) ? o_hi : e_hi;
        }
    }
    __syncwarp();
    for (int i = lane; i < (NBITS - 6) / 8; i += 32) S.vit[i] = 0;
    __syncwarp();
    if (lane == 0) {                                           // trace back from state 0 (:262-277)
        int state = 0;
        for (int i = NBITS - 7, bit = NBITS - 1; i >= 0; i--, bit--) {
            // decision of `state` at step `bit`: state = 2b + e
            const unsigned w = S.dec[bit][state & 1];
            if ((w >> (state >> 1)) & 1u) {
                state |= 64;
                S.vit[i >> 3] |= (uint8_t)(0x80 >> (i & 7));
            }
            state >>= 1;
        }
    }
    __syncwarp();
    // ---- de-scramble into the two RS code blocks (:762-771)
    for (int n = lane; n < RSBLOCKS * NN; n += 32) (&S.rs[0][0])[n] = 0;
    __syncwarp();
    for (int n = lane; n < (NN - RSPAD) * RSBLOCKS; n += 32) {
        const int col = RSPAD + n / RSBLOCKS, row = n % RSBLOCKS;
        S.rs[row][col] = S.vit[n] ^ c_tab.scrambler[n];
    }
    __syncwarp();
    if (lane < RSBLOCKS) S.rserr[lane] = decode_rs_8(S.rs[lane]);      // :774-777
    __syncwarp();
    int rc = (S.rserr[0] == -1 || S.rserr[1] == -1) ? -1 : 0;
    if (rc == 0) {
        for (int j = lane; j < 256; j += 32) S.out[j] = S.rs[j % RSBLOCKS][RSPAD + j / RSBLOCKS];   // :783-789
        __syncwarp();
        if (lane == 0) encode_fec40(S.out, S.reenc);                   // :831-847
        __syncwarp();
        int errors = 0;
        for (int i = lane; i < SYMS; i += 32) errors += (S.reenc[i] != (S.raw[i] >> 7));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) errors += __shfl_xor_sync(0xffffffffu, errors, o);
        rc = errors;
        for (int j = lane; j < 256; j += 32) data_out[(size_t)slot * 256 + j] = S.out[j];
        if (lane == 0) atomicAdd((unsigned long long *)&cnt_dec[m.chan], 1ull);
    } else {
        for (int j = lane; j < 256; j += 32) data_out[(size_t)slot * 256 + j] = 0;
    }
    if (lane == 0) frames[slot].errors = rc;
}

}  // namespace fec
}  // namespace jsdr

// =========================================================================== host
using namespace jsdr;

struct jsdr_fec_state {
    int max_frames = 0;
    int8_t *d_hist[2] = {nullptr, nullptr};   // [nchan][5199]
    int cur = 0;
    fec::FrameMeta *d_frames = nullptr;
    int *d_nframes = nullptr;
    uint8_t *d_data = nullptr;                // [max_frames][256]
    long long *d_cnt = nullptr;               // [2][nchan] cntFEC, cntDec (:567,571)
};

extern "C" int jsdr_bpsk_enable_fec(jsdr_bpsk *b, const int16_t *mettab, int max_frames)
{
    JSDR_REQUIRE(b && mettab && max_frames > 0, JSDR_EINVAL, "bad argument");
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    fec::Tables t;
    fec::build_tables(t, mettab);
    JSDR_CUDA(cudaMemcpyToSymbol(fec::c_tab, &t, sizeof(t)));
    if (b->fec) return JSDR_OK;                                 // tables refreshed, state kept
    jsdr_fec_state *f = new jsdr_fec_state();
    f->max_frames = max_frames;
    const size_t nc = (size_t)b->nchan;
    cudaError_t e = cudaMalloc(&f->d_hist[0], nc * fec::HIST);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_hist[1], nc * fec::HIST);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_frames, sizeof(fec::FrameMeta) * (size_t)max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_nframes, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_data, (size_t)max_frames * 256);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_cnt, sizeof(long long) * 2 * nc);
    if (e != cudaSuccess) {
        set_error("jsdr_bpsk_enable_fec: cudaMalloc: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return JSDR_ENOMEM;
    }
    JSDR_CUDA(cudaMemsetAsync(f->d_hist[0], 0, nc * fec::HIST, ctx->stream));     // dmFECCorr starts as zeros (:503)
    JSDR_CUDA(cudaMemsetAsync(f->d_nframes, 0, sizeof(int), ctx->stream));
    JSDR_CUDA(cudaMemsetAsync(f->d_cnt, 0, sizeof(long long) * 2 * nc, ctx->stream));
    JSDR_CUDA(cudaStreamSynchronize(ctx->stream));            // the stage itself runs on the auxiliary stream
    b->fec = f;
    return JSDR_OK;
}

// called by bpsk_receive after the bit-timing stage; like that stage it runs on the auxiliary stream
int jsdr_fec_after_bits(jsdr_bpsk *b)
{
    jsdr_fec_state *f = b->fec;
    jsdr_ctx *ctx = b->ctx;
    const int nchan = b->nchan, max_bits = b->max_bits;
    JSDR_CUDA(cudaMemsetAsync(f->d_nframes, 0, sizeof(int), ctx->aux));
    if (max_bits > 0) {
        dim3 grid((max_bits + 255) / 256, nchan);
        // cntBit lives at the end of each TimingState
        const long long *cnt_bit = reinterpret_cast<const long long *>(
            reinterpret_cast<const char *>(b->d_ts) + offsetof(jsdr::bpsk::TimingState, cntBit));
        fec::k_sync<<<grid, 256, 0, ctx->aux>>>(f->d_hist[f->cur], b->d_bits, b->d_nbits, max_bits, nchan, cnt_bit,
                                                   (int)(sizeof(jsdr::bpsk::TimingState) / sizeof(long long)), f->d_frames,
                                                   f->d_nframes, f->max_frames, f->d_cnt);
        JSDR_TRY(launched(ctx, "k_sync"));
    }
    const size_t smem = sizeof(fec::DecodeSmem);
    static PerDeviceFlag attr_done;
    if (!attr_done.test_and_set(ctx->device))
        JSDR_CUDA(cudaFuncSetAttribute(fec::k_fec_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fec::k_fec_decode<<<f->max_frames, 32, smem, ctx->aux>>>(f->d_hist[f->cur], b->d_bits, max_bits, f->d_frames,
                                                               f->d_nframes, f->max_frames, f->d_data, f->d_cnt + nchan);
    JSDR_TRY(launched(ctx, "k_fec_decode"));
    dim3 g2((fec::HIST + 255) / 256, nchan);
    fec::k_sync_shift<<<g2, 256, 0, ctx->aux>>>(f->d_hist[f->cur], f->d_hist[f->cur ^ 1], b->d_bits, b->d_nbits, max_bits);
    JSDR_TRY(launched(ctx, "k_sync_shift"));
    f->cur ^= 1;
    return JSDR_OK;
}

void jsdr_fec_destroy(jsdr_bpsk *b)
{
    jsdr_fec_state *f = b->fec;
    if (!f) return;
    void *ptrs[] = {f->d_hist[0], f->d_hist[1], f->d_frames, f->d_nframes, f->d_data, f->d_cnt};
    for (void *p : ptrs) cudaFree(p);
    delete f;
    b->fec = nullptr;
}

extern "C" int jsdr_bpsk_read_frames(jsdr_bpsk *b, int32_t *nframes, int32_t *chan, int64_t *bit_index,
                                     int32_t *errors, uint8_t *data, int max_frames)
{
    JSDR_REQUIRE(b && nframes, JSDR_EINVAL, "null argument");
    JSDR_REQUIRE(b->fec, JSDR_ESTATE, "jsdr_bpsk_enable_fec has not been called");
    jsdr_fec_state *f = b->fec;
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    int n = 0;
    JSDR_CUDA(cudaMemcpyAsync(&n, f->d_nframes, sizeof(int), cudaMemcpyDeviceToHost, ctx->aux));
    JSDR_CUDA(cudaStreamSynchronize(ctx->aux));
    const int have = std::min(n, f->max_frames);
    std::vector<fec::FrameMeta> meta(have);
    std::vector<uint8_t> dat((size_t)have * 256);
    if (have > 0) {
        JSDR_CUDA(cudaMemcpyAsync(meta.data(), f->d_frames, sizeof(fec::FrameMeta) * have, cudaMemcpyDeviceToHost, ctx->aux));
        JSDR_CUDA(cudaMemcpyAsync(dat.data(), f->d_data, (size_t)have * 256, cudaMemcpyDeviceToHost, ctx->aux));
        JSDR_CUDA(cudaStreamSynchronize(ctx->aux));
    }
    // detection order on the device is not deterministic: report by (channel, bit)
    std::vector<int> order(have);
    for (int i = 0; i < have; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        if (meta[x].chan != meta[y].chan) return meta[x].chan < meta[y].chan;
        return meta[x].bit_index < meta[y].bit_index;
    });
    const int give = std::min(have, std::max(max_frames, 0));
    for (int i = 0; i < give; i++) {
        const fec::FrameMeta &m = meta[order[i]];
        if (chan) chan[i] = m.chan;
        if (bit_index) bit_index[i] = m.bit_index;
        if (errors) errors[i] = m.errors;
        if (data) memcpy(data + (size_t)i * 256, dat.data() + (size_t)order[i] * 256, 256);
    }
    *nframes = n;                                               // detections this call (may exceed what fitted)
    return JSDR_OK;
}

extern "C" int jsdr_bpsk_read_fec_counters(jsdr_bpsk *b, int64_t *cnt_fec, int64_t *cnt_dec)
{
    JSDR_REQUIRE(b && cnt_fec && cnt_dec, JSDR_EINVAL, "null argument");
    JSDR_REQUIRE(b->fec, JSDR_ESTATE, "jsdr_bpsk_enable_fec has not been called");
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    const size_t nc = (size_t)b->nchan;
    JSDR_CUDA(cudaMemcpyAsync(cnt_fec, b->fec->d_cnt, sizeof(long long) * nc, cudaMemcpyDeviceToHost, ctx->aux));
    JSDR_CUDA(cudaMemcpyAsync(cnt_dec, b->fec->d_cnt + nc, sizeof(long long) * nc, cudaMemcpyDeviceToHost, ctx->aux));
    JSDR_CUDA(cudaStreamSynchronize(ctx->aux));
    return JSDR_OK;
}
