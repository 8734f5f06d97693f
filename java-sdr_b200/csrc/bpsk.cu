// bpsk.cu — FUNcubeBPSKDemod.java:366-595 for a bank of independent tuners.
//
// Every arithmetic step is binary64 with explicit round-to-nearest intrinsics
// (__dadd_rn / __dmul_rn never contract into FMA), in the reference's operation
// order, so results are bit-identical to the Java arithmetic:
//   RxMixTuner   :382-397   RxDownSample :467-492   RxDemodulate :505-595
//
// How the sequential reference becomes block-parallel:
//   * tuPhase, vcoPhase and dmBitPhase are data-independent accumulators
//     (:384-386, :511-513, :581-583).  A "scout" replays the adds exactly on a
//     side stream and leaves a phase checkpoint every kChunk input samples (per
//     channel) and the table index / roll-over flag per 9600 S/s sample (shared by
//     all channels).  The data kernels then start anywhere.
//   * the tuner + decimator and the matched filter are FIRs over staged,
//     already-mixed samples: one CTA per (channel, tile), shared memory padded so
//     the decimating reads do not bank-conflict.
//   * only the bit-timing tracker (:533-595) is data-dependent; it runs one
//     thread per channel over the 9600 S/s stream.
#include <math.h>
#include <string.h>

#include <vector>

#include "handles.h"

namespace jsdr {
namespace bpsk {

constexpr double kTwoPi = 2.0 * 3.141592653589793;       // Java 2.0*Math.PI
constexpr double kInvTwoPi = 1.0 / kTwoPi;
constexpr int kTileOut = 128;                            // decimator outputs per CTA
constexpr int kTileThreads = 256;                        // 128 outputs x {I,Q}
constexpr int kDmTile = 128;                             // matched-filter outputs per CTA

enum { FMT_F32 = 0, FMT_S16 = 1 };

// FUNcubeBPSKDemod.java:27-55 — F-suffixed literals inside a double[]: each is
// rounded to float first, then widened.
static const float kDsFilterF[27] = {
    -6.103515625000e-004F, -1.220703125000e-004F, +2.380371093750e-003F, +6.164550781250e-003F,
    +7.324218750000e-003F, +7.629394531250e-004F, -1.464843750000e-002F, -3.112792968750e-002F,
    -3.225708007813e-002F, -1.617431640625e-003F, +6.463623046875e-002F, +1.502380371094e-001F,
    +2.231445312500e-001F, +2.518310546875e-001F, +2.231445312500e-001F, +1.502380371094e-001F,
    +6.463623046875e-002F, -1.617431640625e-003F, -3.225708007813e-002F, -3.112792968750e-002F,
    -1.464843750000e-002F, +7.629394531250e-004F, +7.324218750000e-003F, +6.164550781250e-003F,
    +2.380371093750e-003F, -1.220703125000e-004F, -6.103515625000e-004F};
// :58-67 — the reference stores this table twice back to back; one copy here,
// indexed modulo 65.
static const float kDmFilterF[65] = {
    -0.0101130691F, -0.0086975143F, -0.0038246093F, +0.0033563764F, +0.0107237026F, +0.0157790936F,
    +0.0164594107F, +0.0119213911F, +0.0030315224F, -0.0076488191F, -0.0164594107F, -0.0197184277F,
    -0.0150109226F, -0.0023082460F, +0.0154712381F, +0.0327423589F, +0.0424493086F, +0.0379940454F,
    +0.0154712381F, -0.0243701991F, -0.0750320094F, -0.1244834076F, -0.1568500423F, -0.1553748911F,
    -0.1061032953F, -0.0015013786F, +0.1568500423F, +0.3572048240F, +0.5786381191F, +0.7940228249F,
    +0.9744923010F, +1.0945250059F, +1.1366117829F, +1.0945250059F, +0.9744923010F, +0.7940228249F,
    +0.5786381191F, +0.3572048240F, +0.1568500423F, -0.0015013786F, -0.1061032953F, -0.1553748911F,
    -0.1568500423F, -0.1244834076F, -0.0750320094F, -0.0243701991F, +0.0154712381F, +0.0379940454F,
    +0.0424493086F, +0.0327423589F, +0.0154712381F, -0.0023082460F, -0.0150109226F, -0.0197184277F,
    -0.0164594107F, -0.0076488191F, +0.0030315224F, +0.0119213911F, +0.0164594107F, +0.0157790936F,
    +0.0107237026F, +0.0033563764F, -0.0038246093F, -0.0086975143F, -0.0101130691F};

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ double phase_step(double p, double inc)
{   // :384-386 / :511-513  p += inc; if (p > 2pi) p -= 2pi
    p = __dadd_rn(p, inc);
    if (p > kTwoPi) p = __dadd_rn(p, -kTwoPi);
    return p;
}

// (int)(p*256.0/(2.0*Math.PI)) % 256 for p > 0 (:389).  The division is replaced
// by a multiply unless the quotient is within reach of an integer, where the
// IEEE division the reference performs decides.
__device__ __forceinline__ int table_index(double p)
{
    double t = __dmul_rn(p, 256.0);
    double a = t * kInvTwoPi;
    int k;
    if (fabs(a - rint(a)) < 1e-9) k = __double2int_rz(__ddiv_rn(t, kTwoPi));
    else k = __double2int_rz(a);
    return k & 255;
}

template <int FMT>
struct RawT;
template <>
struct RawT<FMT_F32> { typedef float2 type; };
template <>
struct RawT<FMT_S16> { typedef uint32_t type; };

template <int FMT>
__device__ __forceinline__ void raw_to_iq(typename RawT<FMT>::type w, int ic, int qc, double &i, double &q)
{
    if constexpr (FMT == FMT_F32) {
        i = (double)w.x;            // :372-373 (double)buf[n*2]
        q = (double)w.y;
    } else {
        // JavaAudio.java:281-288: s += (short)ic (16-bit wrap); (float)s/(float)Short.MAX_VALUE
        short si = (short)((int)(w & 0xffffu) + ic);
        short sq = (short)((int)(w >> 16) + qc);
        i = (double)__fdiv_rn((float)si, 32767.0f);
        q = (double)__fdiv_rn((float)sq, 32767.0f);
    }
}

// ------------------------------------------------------------------ scouts
// One thread per channel replays tuPhase over the block and leaves the phase
// before every kChunk-th sample.  Data independent: runs on the side stream.
__global__ void k_tuner_scout(const double *__restrict__ inc_, double *__restrict__ phase_,
                              double *__restrict__ chunk_phase, int nchan, int S)
{
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= nchan) return;
    double p = phase_[ch];
    const double inc = inc_[ch];
    int nchunks = (S + kChunk - 1) / kChunk;
    for (int c = 0; c < nchunks; c++) {
        chunk_phase[(size_t)c * nchan + ch] = p;
        int steps = min(kChunk, S - c * kChunk);
        for (int s = 0; s < steps; s++) p = phase_step(p, inc);
    }
    phase_[ch] = p;
}

// vcoPhase (:511-516) and dmBitPhase (:581-584) are the same for every channel of
// a bank: warp 0 replays one, warp 1 the other.
__global__ void k_vco_scout(double *__restrict__ state, uint8_t *__restrict__ vco_ix,
                            uint8_t *__restrict__ bit_roll, int NO, double vco_inc,
                            double bit_inc, double bit_time)
{
    if (threadIdx.x == 0) {
        double p = state[0];
        for (int m = 0; m < NO; m++) {
            p = phase_step(p, vco_inc);
            double q = __ddiv_rn(__dmul_rn(p, 256.0), kTwoPi);
            vco_ix[m] = (uint8_t)(__double2int_rz(q) & 255);
        }
        state[0] = p;
    } else if (threadIdx.x == 32) {
        double b = state[1];
        for (int m = 0; m < NO; m++) {
            b = __dadd_rn(b, bit_inc);
            uint8_t roll = 0;
            if (b >= bit_time) {
                b = __dadd_rn(b, -bit_time);
                roll = 1;
            }
            bit_roll[m] = roll;
        }
        state[1] = b;
    }
}

// ------------------------------------------------------------------ tuner + decimator
struct MixParams {
    const void *in;
    long long chan_stride;     // complex samples between channels (0: shared stream)
    int S;                     // samples in this block
    int ic, qc;
    const double *tu_inc;
    const double *chunk_phase;
    int nchan;
    const double2 *hist_in;    // [nchan][kMaxDsTaps], entry k is local sample k-H
    double2 *hist_out;
    const double *taps;
    int ntaps;
    const double *cossin;
    int D, n0, NO;             // first output's local sample index, outputs this block
    double2 *ds_out;
    int max_ds;
};

// Window index -> shared-memory slot.  For even D one pad slot is inserted per D
// samples so that the decimating reads (lanes D apart) have an odd stride.
// i / D is done as a multiply-high: exact for i, D < 65536 with magic = 2^32/D + 1.
__device__ __forceinline__ int mix_pad(int i, int D, unsigned magic)
{
    return (D & 1) ? i : i + (int)__umulhi((unsigned)i, magic);
}

// shared-memory carve-up for k_mixdecim (all offsets in bytes, 16-aligned)
struct MixSmem {
    int span_pad, raw_n;
    unsigned magic;
    size_t off_I, off_Q, off_tab, off_taps, off_raw, total;
};
static MixSmem mix_smem_layout(int D, int ntaps, int fmt)
{
    MixSmem L;
    int span = kTileOut * D + ntaps + 1;
    L.span_pad = ((D & 1) ? span : span + span / D) + 2;
    L.magic = (unsigned)(4294967296ULL / (unsigned)D) + 1u;
    int chunks = (kTileOut * D + ntaps) / kChunk + 3;
    L.raw_n = chunks * (kChunk + 1);
    size_t o = 0;
    L.off_I = o; o += sizeof(double) * L.span_pad;
    L.off_Q = o; o += sizeof(double) * L.span_pad;
    L.off_tab = o; o += sizeof(double) * 512;
    L.off_taps = o; o += sizeof(double) * kMaxDsTaps;
    L.off_raw = o; o += (fmt == FMT_S16 ? 4 : 8) * (size_t)L.raw_n;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

template <int FMT>
__global__ void __launch_bounds__(kTileThreads)
k_mixdecim(const MixParams p, const MixSmem L)
{
    typedef typename RawT<FMT>::type raw_t;
    extern __shared__ __align__(16) unsigned char smem[];
    double *sI = reinterpret_cast<double *>(smem + L.off_I);
    double *sQ = reinterpret_cast<double *>(smem + L.off_Q);
    double *sTab = reinterpret_cast<double *>(smem + L.off_tab);
    double *sTaps = reinterpret_cast<double *>(smem + L.off_taps);
    raw_t *sRaw = reinterpret_cast<raw_t *>(smem + L.off_raw);

    const int tid = threadIdx.x;
    const int ch = blockIdx.y;
    const int m0 = blockIdx.x * kTileOut;
    const int cnt = min(kTileOut, p.NO - m0);
    const int D = p.D, H = p.ntaps - 1;
    const int n_hi = p.n0 + (m0 + cnt - 1) * D;       // newest sample this tile needs
    const int n_lo = p.n0 + m0 * D - H;               // oldest (negative: history)
    const int c_lo = max(n_lo, 0) / kChunk;
    const int c_hi = n_hi / kChunk;
    const int r0 = c_lo * kChunk;

    for (int i = tid; i < 512; i += kTileThreads) sTab[i] = p.cossin[i];
    for (int i = tid; i < p.ntaps; i += kTileThreads) sTaps[i] = p.taps[i];

    // coalesced raw load, one pad word per chunk
    {
        const raw_t *src = reinterpret_cast<const raw_t *>(p.in) + (long long)ch * p.chan_stride;
        int r_end = min((c_hi + 1) * kChunk, p.S) - r0;
        for (int r = tid; r < r_end; r += kTileThreads) sRaw[r + r / kChunk] = src[r0 + r];
    }
    // history part of the window (samples before this block)
    if (n_lo < 0) {
        const double2 *h = p.hist_in + (size_t)ch * kMaxDsTaps;
        for (int ii = tid; ii < -n_lo; ii += kTileThreads) {
            double2 v = h[n_lo + ii + H];
            int f = mix_pad(ii, D, L.magic);
            sI[f] = v.x;
            sQ[f] = v.y;
        }
    }
    __syncthreads();

    // mix: one thread per checkpoint chunk, phase replayed exactly (:384-396)
    {
        const double inc = p.tu_inc[ch];
        for (int c = c_lo + tid; c <= c_hi; c += kTileThreads) {
            double ph = p.chunk_phase[(size_t)c * p.nchan + ch];
            const raw_t *rw = sRaw + (c - c_lo) * (kChunk + 1);
            int n = c * kChunk;
            const int s_end = min(kChunk, p.S - n);
            for (int s = 0; s < s_end; s++, n++) {
                ph = phase_step(ph, inc);
                const int ii = n - n_lo;               // window index of sample n
                if (ii < 0 || n > n_hi) continue;      // outside the window: only the phase advances
                double xi, xq;
                raw_to_iq<FMT>(rw[s], p.ic, p.qc, xi, xq);
                if (ph > 0.0) {                         // :388
                    int ix = table_index(ph);
                    xi = __dmul_rn(xi, sTab[ix]);        // i*cosTab[ix]
                    xq = __dmul_rn(xq, sTab[256 + ix]);  // q*sinTab[ix]
                }
                int f = mix_pad(ii, D, L.magic);
                sI[f] = xi;
                sQ[f] = xq;
            }
        }
    }
    __syncthreads();

    // decimating FIR (:477-486): newest sample first, one thread per output and component
    {
        const int lane_out = tid & (kTileOut - 1);
        const int comp = tid / kTileOut;               // 0: I, 1: Q (warp uniform)
        if (lane_out < cnt) {
            const double *sX = comp ? sQ : sI;
            const int top = (p.n0 + (m0 + lane_out) * D) - n_lo;
            double acc = 0.0;
            for (int k = 0; k < p.ntaps; k++)
                acc = __dadd_rn(acc, __dmul_rn(sX[mix_pad(top - k, D, L.magic)], sTaps[k]));
            // :469,486  fi * HOWARD_FUDGE_FACTOR, with 0.9*32768.0 folded by javac to one double
            double outv = __dmul_rn(acc, 0.9 * 32768.0);
            double *o = reinterpret_cast<double *>(p.ds_out + (size_t)ch * p.max_ds + m0 + lane_out);
            o[comp] = outv;
        }
    }
}

// New history = the last H mixed samples of the block (or old history shifted,
// for blocks shorter than H).  One CTA per channel.
template <int FMT>
__global__ void k_tuner_tail(const MixParams p)
{
    typedef typename RawT<FMT>::type raw_t;
    __shared__ double sTab[512];
    const int ch = blockIdx.x;
    const int H = p.ntaps - 1;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sTab[i] = p.cossin[i];
    __syncthreads();
    const int k = threadIdx.x;
    if (k >= H) return;
    const int n = p.S - H + k;
    double2 v;
    if (n < 0) {
        v = p.hist_in[(size_t)ch * kMaxDsTaps + k + p.S];
    } else {
        const raw_t *src = reinterpret_cast<const raw_t *>(p.in) + (long long)ch * p.chan_stride;
        int c = n / kChunk;
        double ph = p.chunk_phase[(size_t)c * p.nchan + ch];
        const double inc = p.tu_inc[ch];
        for (int s = c * kChunk; s <= n; s++) ph = phase_step(ph, inc);
        double xi, xq;
        raw_to_iq<FMT>(src[n], p.ic, p.qc, xi, xq);
        if (ph > 0.0) {
            int ix = table_index(ph);
            xi = __dmul_rn(xi, sTab[ix]);
            xq = __dmul_rn(xq, sTab[256 + ix]);
        }
        v = make_double2(xi, xq);
    }
    p.hist_out[(size_t)ch * kMaxDsTaps + k] = v;
}

// ------------------------------------------------------------------ matched filter
struct DmParams {
    const double2 *ds;         // [nchan][max_ds]
    int max_ds, NO;
    const uint8_t *vco_ix;
    const double2 *hist_in;    // [nchan][64], entry k is local sample k-64
    double2 *hist_out;
    const double *dmtaps;
    const double *cossin;
    int base65;                // cntDS before this block, mod 65
    double2 *dm_out;
};

__device__ __forceinline__ double2 vco_mix(const DmParams &p, const double *sTab, int ch, int m)
{   // :515-516
    double2 d = p.ds[(size_t)ch * p.max_ds + m];
    int ix = p.vco_ix[m];
    return make_double2(__dmul_rn(d.x, sTab[ix]), __dmul_rn(d.y, sTab[256 + ix]));
}

__global__ void __launch_bounds__(2 * kDmTile) k_matched(const DmParams p)
{
    __shared__ double sTab[512];
    __shared__ double sH[kDmTaps];
    __shared__ double vI[kDmTile + 64], vQ[kDmTile + 64];
    const int tid = threadIdx.x, ch = blockIdx.y;
    const int m0 = blockIdx.x * kDmTile;
    const int cnt = min(kDmTile, p.NO - m0);
    for (int i = tid; i < 512; i += blockDim.x) sTab[i] = p.cossin[i];
    for (int i = tid; i < kDmTaps; i += blockDim.x) sH[i] = p.dmtaps[i];
    __syncthreads();
    for (int ii = tid; ii < cnt + 64; ii += blockDim.x) {
        int m = m0 - 64 + ii;
        double2 v = (m < 0) ? p.hist_in[(size_t)ch * 64 + 64 + m] : vco_mix(p, sTab, ch, m);
        vI[ii] = v.x;
        vQ[ii] = v.y;
    }
    __syncthreads();
    const int lo = tid & (kDmTile - 1), comp = tid / kDmTile;
    if (lo < cnt) {
        // :519-523 sums by buffer slot: ages a0, a0+1, .., 64, 0, .., a0-1 with
        // a0 = (calls so far + 1) mod 65
        const double *vX = comp ? vQ : vI;
        int age = (p.base65 + m0 + lo + 1) % 65;
        const int top = lo + 64;
        double acc = 0.0;
        for (int n = 0; n < kDmTaps; n++) {
            acc = __dadd_rn(acc, __dmul_rn(vX[top - age], sH[age]));
            if (++age == 65) age = 0;
        }
        double *o = reinterpret_cast<double *>(p.dm_out + (size_t)ch * p.max_ds + m0 + lo);
        o[comp] = acc;
    }
}

__global__ void k_dm_tail(const DmParams p)
{
    __shared__ double sTab[512];
    const int ch = blockIdx.x, k = threadIdx.x;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sTab[i] = p.cossin[i];
    __syncthreads();
    if (k >= 64) return;
    int m = p.NO - 64 + k;
    double2 v = (m < 0) ? p.hist_in[(size_t)ch * 64 + k + p.NO] : vco_mix(p, sTab, ch, m);
    p.hist_out[(size_t)ch * 64 + k] = v;
}

// ------------------------------------------------------------------ bit timing + decision
struct TimingParams {
    const double2 *dm;
    int max_ds, NO, nchan;
    const uint8_t *bit_roll;
    TimingState *ts;
    int8_t *bits;
    long long *bit_at;
    int32_t *nbits;
    int max_bits;
    long long cnt_ds0;
};

__global__ void __launch_bounds__(128) k_timing(const TimingParams p)
{
    // :533-595, one thread per channel; dmEnergy[] lives in shared memory because it
    // is indexed by the running bit position
    __shared__ double sE[8][128];
    const int tid = threadIdx.x;
    const int ch = blockIdx.x * 128 + tid;
    if (ch >= p.nchan) return;
    TimingState st = p.ts[ch];
#pragma unroll
    for (int i = 0; i < 8; i++) sE[i][tid] = st.dmEnergy[i];
    const double S1 = 1.0 / 200.0, S2 = 1.0 / 800.0;
    const double C1 = 1.0 - S1, C2 = 1.0 - S2;
    double eOut = st.dmEnergyOut, lastI = st.lastI, lastQ = st.lastQ;
    int bitPos = st.bitPos, peakPos = st.peakPos, newPeak = st.newPeak;
    int nb = 0;
    const double2 *dm = p.dm + (size_t)ch * p.max_ds;
    int8_t *bits = p.bits + (size_t)ch * p.max_bits;
    long long *bit_at = p.bit_at + (size_t)ch * p.max_bits;
    for (int m = 0; m < p.NO; m++) {
        double2 f = dm[m];
        double energy1 = __dadd_rn(__dmul_rn(f.x, f.x), __dmul_rn(f.y, f.y));            // :534
        sE[bitPos][tid] = __dadd_rn(__dmul_rn(sE[bitPos][tid], C1), __dmul_rn(energy1, S1));   // :535
        if (bitPos == peakPos) {                                                       // :537
            eOut = __dadd_rn(__dmul_rn(eOut, C2), __dmul_rn(energy1, S2));
            double di = -__dadd_rn(__dmul_rn(lastI, f.x), __dmul_rn(lastQ, f.y));      // :539
            double dq = __dadd_rn(__dmul_rn(lastI, f.y), -__dmul_rn(lastQ, f.x));      // :540
            lastI = f.x;
            lastQ = f.y;
            double energy2 = __dsqrt_rn(__dadd_rn(__dmul_rn(di, di), __dmul_rn(dq, dq)));
            if (energy2 > 100.0) {                                                     // :544
                if (nb < p.max_bits) {
                    bits[nb] = (di < 0.0) ? 1 : -1;                                    // :545,554
                    bit_at[nb] = p.cnt_ds0 + m;
                }
                nb++;
            }
        }
        if (bitPos == ((peakPos + 4) & 7)) peakPos = newPeak;      // :577 dmHalfTable = {4,5,6,7,0,1,2,3}
        bitPos = (bitPos + 1) & 7;                                 // :579
        if (p.bit_roll[m]) {                                       // :582-592
            bitPos = 0;
            double eMax = -1.0e10;                                 // (double)-1.0e10F, exact
#pragma unroll
            for (int n = 0; n < 8; n++) {
                double e = sE[n][tid];
                if (e > eMax) { newPeak = n; eMax = e; }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) st.dmEnergy[i] = sE[i][tid];
    st.dmEnergyOut = eOut;
    st.lastI = lastI;
    st.lastQ = lastQ;
    st.bitPos = bitPos;
    st.peakPos = peakPos;
    st.newPeak = newPeak;
    st.cntBit += nb;
    p.ts[ch] = st;
    p.nbits[ch] = nb;
}

}  // namespace bpsk
}  // namespace jsdr

// =========================================================================== host
using namespace jsdr;
using namespace jsdr::bpsk;

namespace {

int upload(jsdr_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    JSDR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    JSDR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JSDR_OK;
}

// Buffers of the 9600 S/s stages, allocated when a receive first needs them (a
// stages==1 bank, e.g. the mix+FIR benchmark, never pays for them).
int ensure_stage_buffers(jsdr_bpsk *b)
{
    if (b->d_dm_out) return JSDR_OK;
    jsdr_ctx *ctx = b->ctx;
    const size_t nc = (size_t)b->nchan;
    struct { void **p; size_t bytes; } req[] = {
        {(void **)&b->d_vco_ix, (size_t)b->max_ds},
        {(void **)&b->d_bit_roll, (size_t)b->max_ds},
        {(void **)&b->d_dm_hist[0], sizeof(double2) * 64 * nc},
        {(void **)&b->d_dm_hist[1], sizeof(double2) * 64 * nc},
        {(void **)&b->d_bits, nc * b->max_bits},
        {(void **)&b->d_bit_at, sizeof(long long) * nc * b->max_bits},
        {(void **)&b->d_dm_out, sizeof(double2) * nc * b->max_ds},
    };
    for (auto &r : req) {
        cudaError_t e = cudaMalloc(r.p, r.bytes);
        if (e != cudaSuccess) {
            set_error("stage buffers: cudaMalloc(%zu): %s", r.bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return JSDR_ENOMEM;
        }
        JSDR_CUDA(cudaMemsetAsync(*r.p, 0, r.bytes, ctx->stream));
    }
    return JSDR_OK;
}

int bpsk_reset_ds(jsdr_bpsk *b)
{
    jsdr_ctx *ctx = b->ctx;
    for (int i = 0; i < 2; i++)
        JSDR_CUDA(cudaMemsetAsync(b->d_ds_hist[i], 0, sizeof(double2) * kMaxDsTaps * (size_t)b->nchan, ctx->stream));
    b->ds_cnt = 0;
    b->ds_hist_cur = 0;
    return JSDR_OK;
}

// `after_input` (optional) is called once the input is on the device and before
// the main stream waits for the scouts: the pump enqueues the FFT there, so the
// data-independent phase replay hides behind it.
typedef int (*after_input_fn)(void *user, const void *d_in);

template <int FMT>
int bpsk_receive(jsdr_bpsk *b, const void *in, int S, long long chan_stride, int ic, int qc, int mem,
                 after_input_fn after_input = nullptr, void *user = nullptr)
{
    JSDR_REQUIRE(b && in, JSDR_EINVAL, "null argument");
    JSDR_REQUIRE(S >= 0 && S <= b->max_block, JSDR_EINVAL, "nsamples exceeds max_block_samples");
    JSDR_REQUIRE(chan_stride == 0 || chan_stride >= S, JSDR_EINVAL, "chan_stride smaller than nsamples");
    JSDR_REQUIRE(mem == JSDR_MEM_HOST || mem == JSDR_MEM_DEVICE, JSDR_EINVAL, "bad mem");
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    b->last_nds = 0;
    if (S == 0) {
        JSDR_CUDA(cudaMemsetAsync(b->d_nbits, 0, sizeof(int32_t) * b->nchan, ctx->stream));
        return JSDR_OK;
    }
    const int D = b->D;
    const int NO = (b->ds_cnt + S) / D;
    const int n0 = D - 1 - b->ds_cnt;
    const int nchan = b->nchan;
    if (b->stages >= 2) JSDR_TRY(ensure_stage_buffers(b));

    // ---- fork: data-independent phase scouts on the side stream
    JSDR_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    JSDR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    k_tuner_scout<<<(nchan + 127) / 128, 128, 0, ctx->side>>>(b->d_tu_inc, b->d_tu_phase, b->d_chunk_phase, nchan, S);
    JSDR_TRY(launched(ctx, "k_tuner_scout"));
    if (b->stages >= 2 && NO > 0) {
        const double vco_inc = 2.0 * M_PI * 1200.0 / (double)9600;      // :88
        const double bit_inc = 1.0 / (double)9600, bit_time = 1.0 / (double)1200;   // :91-92
        k_vco_scout<<<1, 64, 0, ctx->side>>>(b->d_vco_state, b->d_vco_ix, b->d_bit_roll, NO, vco_inc, bit_inc, bit_time);
        JSDR_TRY(launched(ctx, "k_vco_scout"));
    }
    JSDR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));

    // ---- input
    const void *d_in = in;
    if (mem == JSDR_MEM_HOST) {
        const size_t esz = (FMT == FMT_S16) ? 4 : 8;
        const size_t total = (chan_stride == 0) ? (size_t)S : (size_t)chan_stride * (nchan - 1) + S;
        if (b->in_cap < total * esz) {
            cudaFree(b->d_in);
            b->d_in = nullptr;
            b->in_cap = 0;
            JSDR_CUDA(cudaMalloc(&b->d_in, total * esz));
            b->in_cap = total * esz;
        }
        JSDR_CUDA(cudaMemcpyAsync(b->d_in, in, total * esz, cudaMemcpyHostToDevice, ctx->stream));
        d_in = b->d_in;
    }
    if (after_input) JSDR_TRY(after_input(user, d_in));
    JSDR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));

    // ---- tuner + decimator
    MixParams mp;
    mp.in = d_in;
    mp.chan_stride = chan_stride;
    mp.S = S;
    mp.ic = ic;
    mp.qc = qc;
    mp.tu_inc = b->d_tu_inc;
    mp.chunk_phase = b->d_chunk_phase;
    mp.nchan = nchan;
    mp.hist_in = b->d_ds_hist[b->ds_hist_cur];
    mp.hist_out = b->d_ds_hist[b->ds_hist_cur ^ 1];
    mp.taps = b->d_taps;
    mp.ntaps = b->ntaps;
    mp.cossin = b->d_cossin;
    mp.D = D;
    mp.n0 = n0;
    mp.NO = NO;
    mp.ds_out = b->d_ds_out;
    mp.max_ds = b->max_ds;
    if (NO > 0) {
        MixSmem L = mix_smem_layout(D, b->ntaps, FMT);
        static size_t attr_set[2] = {0, 0};
        if (attr_set[FMT] < L.total) {
            JSDR_CUDA(cudaFuncSetAttribute(k_mixdecim<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
            attr_set[FMT] = L.total;
        }
        dim3 grid((NO + kTileOut - 1) / kTileOut, nchan);
        k_mixdecim<FMT><<<grid, kTileThreads, L.total, ctx->stream>>>(mp, L);
        JSDR_TRY(launched(ctx, "k_mixdecim"));
    }
    if (b->ntaps > 1) {
        k_tuner_tail<FMT><<<nchan, 128, 0, ctx->stream>>>(mp);
        JSDR_TRY(launched(ctx, "k_tuner_tail"));
    }
    b->ds_hist_cur ^= 1;
    b->ds_cnt = (b->ds_cnt + S) % D;
    b->cnt_raw += S;
    b->last_nds = NO;

    // ---- matched filter, bit timing
    if (b->stages >= 2 && NO > 0) {
        DmParams dp;
        dp.ds = b->d_ds_out;
        dp.max_ds = b->max_ds;
        dp.NO = NO;
        dp.vco_ix = b->d_vco_ix;
        dp.hist_in = b->d_dm_hist[b->dm_hist_cur];
        dp.hist_out = b->d_dm_hist[b->dm_hist_cur ^ 1];
        dp.dmtaps = b->d_dmtaps;
        dp.cossin = b->d_cossin;
        dp.base65 = (int)(b->cnt_ds % 65);
        dp.dm_out = b->d_dm_out;
        dim3 grid((NO + kDmTile - 1) / kDmTile, nchan);
        k_matched<<<grid, 2 * kDmTile, 0, ctx->stream>>>(dp);
        JSDR_TRY(launched(ctx, "k_matched"));
        k_dm_tail<<<nchan, 64, 0, ctx->stream>>>(dp);
        JSDR_TRY(launched(ctx, "k_dm_tail"));
        b->dm_hist_cur ^= 1;
        if (b->stages >= 3) {
            TimingParams tp;
            tp.dm = b->d_dm_out;
            tp.max_ds = b->max_ds;
            tp.NO = NO;
            tp.nchan = nchan;
            tp.bit_roll = b->d_bit_roll;
            tp.ts = b->d_ts;
            tp.bits = b->d_bits;
            tp.bit_at = b->d_bit_at;
            tp.nbits = b->d_nbits;
            tp.max_bits = b->max_bits;
            tp.cnt_ds0 = b->cnt_ds;
            k_timing<<<(nchan + 127) / 128, 128, 0, ctx->stream>>>(tp);
            JSDR_TRY(launched(ctx, "k_timing"));
        }
    }
    if (NO == 0)   // nothing reached the 9600 S/s stage in this call: no bits either
        JSDR_CUDA(cudaMemsetAsync(b->d_nbits, 0, sizeof(int32_t) * nchan, ctx->stream));
    b->cnt_ds += NO;
    if (mem == JSDR_MEM_HOST) JSDR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JSDR_OK;
}

}  // namespace

extern "C" int jsdr_bpsk_create(jsdr_ctx *ctx, int rate, int nchan, const double *tuning_hz,
                                int max_block_samples, jsdr_bpsk **out)
{
    JSDR_REQUIRE(ctx && out && tuning_hz, JSDR_EINVAL, "null argument");
    JSDR_REQUIRE(rate >= 9600 && nchan > 0 && max_block_samples > 0, JSDR_EINVAL,
                 "need rate >= 9600, nchan > 0, max_block_samples > 0");
    JSDR_TRY(ctx->bind());
    jsdr_bpsk *b = new jsdr_bpsk();
    b->ctx = ctx;
    b->rate = rate;
    b->D = rate / 9600;                 // :476 adsc.rate/DOWN_SAMPLE_RATE (integer division)
    b->nchan = nchan;
    b->max_block = max_block_samples;
    b->max_ds = max_block_samples / b->D + 2;
    b->max_chunks = (max_block_samples + kChunk - 1) / kChunk + 1;
    b->max_bits = b->max_ds;
    b->h_tuning.assign(tuning_hz, tuning_hz + nchan);
    const size_t nc = (size_t)nchan;
#define ALLOC(ptr, bytes)                                            \
    do {                                                             \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));       \
        if (e_ != cudaSuccess) {                                     \
            set_error("jsdr_bpsk_create: cudaMalloc(%zu): %s", (size_t)(bytes), cudaGetErrorString(e_)); \
            jsdr_bpsk_destroy(b);                                    \
            cudaGetLastError();                                      \
            return JSDR_ENOMEM;                                      \
        }                                                            \
        cudaMemsetAsync((ptr), 0, (bytes), ctx->stream);             \
    } while (0)
    ALLOC(b->d_taps, sizeof(double) * kMaxDsTaps);
    ALLOC(b->d_dmtaps, sizeof(double) * kDmTaps);
    ALLOC(b->d_cossin, sizeof(double) * 512);
    ALLOC(b->d_tu_inc, sizeof(double) * nc);
    ALLOC(b->d_tu_phase, sizeof(double) * nc);
    ALLOC(b->d_chunk_phase, sizeof(double) * nc * b->max_chunks);
    ALLOC(b->d_ds_hist[0], sizeof(double2) * kMaxDsTaps * nc);
    ALLOC(b->d_ds_hist[1], sizeof(double2) * kMaxDsTaps * nc);
    ALLOC(b->d_ds_out, sizeof(double2) * nc * b->max_ds);
    ALLOC(b->d_vco_state, sizeof(double) * 2);
    ALLOC(b->d_ts, sizeof(TimingState) * nc);
    ALLOC(b->d_nbits, sizeof(int32_t) * nc);
#undef ALLOC
    // tables and constants, computed on the host exactly as the reference's setup code does
    std::vector<double> cossin(512);
    for (int n = 0; n < 256; n++) {                       // :159-162
        cossin[n] = cos(n * 2.0 * M_PI / 256);
        cossin[256 + n] = sin(n * 2.0 * M_PI / 256);
    }
    std::vector<double> taps(kMaxDsTaps, 0.0), dmt(kDmTaps);
    for (int i = 0; i < 27; i++) taps[i] = (double)kDsFilterF[i];
    for (int i = 0; i < kDmTaps; i++) dmt[i] = (double)kDmFilterF[i];
    std::vector<double> inc(nchan);
    for (int c = 0; c < nchan; c++) inc[c] = 2.0 * M_PI * tuning_hz[c] / (double)rate;   // :196
    std::vector<TimingState> ts(nchan);
    memset(ts.data(), 0, sizeof(TimingState) * nc);
    for (int c = 0; c < nchan; c++) ts[c].dmEnergyOut = 1.0;                             // :499
    int rc = upload(ctx, b->d_cossin, cossin.data(), sizeof(double) * 512);
    if (rc == JSDR_OK) rc = upload(ctx, b->d_taps, taps.data(), sizeof(double) * kMaxDsTaps);
    if (rc == JSDR_OK) rc = upload(ctx, b->d_dmtaps, dmt.data(), sizeof(double) * kDmTaps);
    if (rc == JSDR_OK) rc = upload(ctx, b->d_tu_inc, inc.data(), sizeof(double) * nc);
    if (rc == JSDR_OK) rc = upload(ctx, b->d_ts, ts.data(), sizeof(TimingState) * nc);
    if (rc != JSDR_OK) {
        jsdr_bpsk_destroy(b);
        return rc;
    }
    *out = b;
    return JSDR_OK;
}

extern "C" int jsdr_bpsk_destroy(jsdr_bpsk *b)
{
    if (!b) return JSDR_OK;
    b->ctx->bind();
    cudaStreamSynchronize(b->ctx->side);
    cudaStreamSynchronize(b->ctx->stream);
    void *ptrs[] = {b->d_taps, b->d_dmtaps, b->d_cossin, b->d_tu_inc, b->d_tu_phase, b->d_chunk_phase,
                    b->d_ds_hist[0], b->d_ds_hist[1], b->d_ds_out, b->d_vco_state, b->d_vco_ix,
                    b->d_bit_roll, b->d_dm_hist[0], b->d_dm_hist[1], b->d_dm_out, b->d_ts, b->d_bits,
                    b->d_bit_at, b->d_nbits, b->d_in};
    for (void *p : ptrs) cudaFree(p);
    delete b;
    return JSDR_OK;
}

extern "C" int jsdr_bpsk_set_stages(jsdr_bpsk *b, int stages)
{
    JSDR_REQUIRE(b && stages >= 1 && stages <= 3, JSDR_EINVAL, "stages must be 1..3");
    b->stages = stages;
    return JSDR_OK;
}

extern "C" int jsdr_bpsk_set_tuning(jsdr_bpsk *b, int chan, double hz)
{
    JSDR_REQUIRE(b && chan >= 0 && chan < b->nchan, JSDR_EINVAL, "bad channel");
    JSDR_TRY(b->ctx->bind());
    b->h_tuning[chan] = hz;
    double inc = 2.0 * M_PI * hz / (double)b->rate;       // :188
    return upload(b->ctx, b->d_tu_inc + chan, &inc, sizeof(double));
}

extern "C" int jsdr_bpsk_set_ds_filter(jsdr_bpsk *b, const double *taps, int ntaps)
{
    JSDR_REQUIRE(b && taps && ntaps >= 1 && ntaps <= kMaxDsTaps, JSDR_EINVAL, "1..128 taps");
    JSDR_TRY(b->ctx->bind());
    std::vector<double> t(kMaxDsTaps, 0.0);
    for (int i = 0; i < ntaps; i++) t[i] = taps[i];
    JSDR_TRY(upload(b->ctx, b->d_taps, t.data(), sizeof(double) * kMaxDsTaps));
    b->ntaps = ntaps;
    return bpsk_reset_ds(b);
}

extern "C" int jsdr_bpsk_receive_f32(jsdr_bpsk *b, const float *iq, int nsamples, int64_t chan_stride, int mem)
{
    return bpsk_receive<FMT_F32>(b, iq, nsamples, chan_stride, 0, 0, mem);
}

extern "C" int jsdr_bpsk_receive_s16(jsdr_bpsk *b, const int16_t *raw, int nsamples, int64_t chan_stride,
                                     int ic, int qc, int mem)
{
    return bpsk_receive<FMT_S16>(b, raw, nsamples, chan_stride, ic, qc, mem);
}

extern "C" int jsdr_bpsk_last_counts(jsdr_bpsk *b, int32_t *n_ds)
{
    JSDR_REQUIRE(b && n_ds, JSDR_EINVAL, "null argument");
    *n_ds = b->last_nds;
    return JSDR_OK;
}

namespace {
int read_rows(jsdr_bpsk *b, const double2 *src, double *out, int mem)
{
    JSDR_REQUIRE(b && out, JSDR_EINVAL, "null argument");
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    if (b->last_nds == 0) return JSDR_OK;
    JSDR_CUDA(cudaMemcpy2DAsync(out, sizeof(double2) * b->last_nds, src, sizeof(double2) * b->max_ds,
                                sizeof(double2) * b->last_nds, b->nchan,
                                mem == JSDR_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                ctx->stream));
    if (mem == JSDR_MEM_HOST) JSDR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JSDR_OK;
}
}  // namespace

extern "C" int jsdr_bpsk_read_ds(jsdr_bpsk *b, double *out, int mem) { return read_rows(b, b ? b->d_ds_out : nullptr, out, mem); }

extern "C" int jsdr_bpsk_read_dm(jsdr_bpsk *b, double *out, int mem)
{
    JSDR_REQUIRE(b && b->stages >= 2 && b->d_dm_out, JSDR_ESTATE, "matched filter stage is disabled or has not run");
    return read_rows(b, b->d_dm_out, out, mem);
}

extern "C" int jsdr_bpsk_read_bits(jsdr_bpsk *b, int8_t *bits, int64_t *bit_at, int32_t *nbits,
                                   int max_bits, int mem)
{
    JSDR_REQUIRE(b && nbits, JSDR_EINVAL, "null argument");
    JSDR_REQUIRE(b->stages >= 3 && b->d_bits, JSDR_ESTATE, "bit decision stage is disabled or has not run");
    JSDR_REQUIRE(max_bits >= 0, JSDR_EINVAL, "negative max_bits");
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    const cudaMemcpyKind kind = mem == JSDR_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    JSDR_CUDA(cudaMemcpyAsync(nbits, b->d_nbits, sizeof(int32_t) * b->nchan, kind, ctx->stream));
    const int w = max_bits < b->max_bits ? max_bits : b->max_bits;
    if (bits && w > 0)
        JSDR_CUDA(cudaMemcpy2DAsync(bits, (size_t)max_bits, b->d_bits, (size_t)b->max_bits, (size_t)w,
                                    b->nchan, kind, ctx->stream));
    if (bit_at && w > 0)
        JSDR_CUDA(cudaMemcpy2DAsync(bit_at, sizeof(int64_t) * max_bits, b->d_bit_at,
                                    sizeof(long long) * b->max_bits, sizeof(int64_t) * w, b->nchan, kind,
                                    ctx->stream));
    if (mem == JSDR_MEM_HOST) JSDR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JSDR_OK;
}

extern "C" int jsdr_bpsk_read_counters(jsdr_bpsk *b, int64_t *counters)
{
    JSDR_REQUIRE(b && counters, JSDR_EINVAL, "null argument");
    jsdr_ctx *ctx = b->ctx;
    JSDR_TRY(ctx->bind());
    std::vector<TimingState> ts(b->nchan);
    JSDR_CUDA(cudaMemcpyAsync(ts.data(), b->d_ts, sizeof(TimingState) * (size_t)b->nchan,
                              cudaMemcpyDeviceToHost, ctx->stream));
    JSDR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < b->nchan; c++) {
        counters[4 * c + 0] = b->cnt_raw;
        counters[4 * c + 1] = b->cnt_ds;
        counters[4 * c + 2] = ts[c].cntBit;
        counters[4 * c + 3] = 0;
    }
    return JSDR_OK;
}

extern "C" int jsdr_bpsk_ds_device_ptr(jsdr_bpsk *b, double **dev_ptr)
{
    JSDR_REQUIRE(b && dev_ptr, JSDR_EINVAL, "null argument");
    *dev_ptr = reinterpret_cast<double *>(b->d_ds_out);
    return JSDR_OK;
}

// ------------------------------------------------------------------ the pump
namespace {
struct PumpJob {
    jsdr_fft *f;
    int batch, ic, qc;
    float *d_psd;
    int32_t *d_peak;
};
int pump_fft(void *user, const void *d_in)
{
    PumpJob *j = static_cast<PumpJob *>(user);
    return fft::launch(j->f, d_in, fft::IN_S16, j->batch, j->d_psd, j->d_peak, fft::OUT_PSD, j->ic, j->qc,
                       j->f->ctx->stream);
}
}  // namespace

// JavaAudio.run's fan-out (JavaAudio.java:262-304) for a batch: each channel's
// blocks go to the fft handler and to the tuner bank.
extern "C" int jsdr_pump_receive_s16(jsdr_fft *f, jsdr_bpsk *b, const int16_t *raw, int nblocks,
                                     float *psd, int32_t *peak_bin, int mem)
{
    JSDR_REQUIRE(f && b && raw && psd, JSDR_EINVAL, "null argument");
    JSDR_REQUIRE(f->ctx == b->ctx, JSDR_EINVAL, "handlers belong to different contexts");
    JSDR_REQUIRE(nblocks > 0, JSDR_EINVAL, "nblocks must be positive");
    const long long S = (long long)nblocks * f->n;
    const long long batch = (long long)nblocks * b->nchan;
    JSDR_REQUIRE(S <= b->max_block, JSDR_EINVAL, "nblocks*n exceeds the bank's max_block_samples");
    JSDR_REQUIRE(batch <= f->max_batch, JSDR_EINVAL, "nchan*nblocks exceeds the fft's max_batch");
    jsdr_ctx *ctx = f->ctx;
    JSDR_TRY(ctx->bind());
    PumpJob job;
    job.f = f;
    job.batch = (int)batch;
    job.ic = 0;
    job.qc = 0;
    const size_t psd_elems = (size_t)batch * (f->n + 2);
    if (mem == JSDR_MEM_DEVICE) {
        job.d_psd = psd;
        job.d_peak = peak_bin;
        return bpsk_receive<FMT_S16>(b, raw, (int)S, S, 0, 0, mem, pump_fft, &job);
    }
    if (f->out_cap < psd_elems * sizeof(float)) {
        cudaFree(f->d_out);
        f->d_out = nullptr;
        f->out_cap = 0;
        JSDR_CUDA(cudaMalloc(&f->d_out, psd_elems * sizeof(float)));
        f->out_cap = psd_elems * sizeof(float);
    }
    if (!f->d_peak) JSDR_CUDA(cudaMalloc(&f->d_peak, sizeof(int32_t) * (size_t)f->max_batch));
    job.d_psd = f->d_out;
    job.d_peak = f->d_peak;
    JSDR_TRY(bpsk_receive<FMT_S16>(b, raw, (int)S, S, 0, 0, mem, pump_fft, &job));
    JSDR_CUDA(cudaMemcpyAsync(psd, f->d_out, psd_elems * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (peak_bin)
        JSDR_CUDA(cudaMemcpyAsync(peak_bin, f->d_peak, sizeof(int32_t) * (size_t)batch,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    JSDR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JSDR_OK;
}
