// bpsk_stream2.cuh — the streaming tuner + decimator with a PERIOD ring staged by asynchronous
// copies (FUNcubeBPSKDemod.java:382-397 RxMixTuner, :467-492 RxDownSample), s16 input, D % 4 == 0.
// Included by bpsk.cu after bpsk_stream.cuh, whose mapping it keeps: one lane per channel, one
// warp per (32 channels, segment of R outputs), each lane walking its channel backwards in time one
// period of D samples at a time, every sample converted and mixed once and feeding the
// ceil(NTAPS/D) live outputs — so the arithmetic, and with it every output bit, is that kernel's.
//
// What changes is how the raw IQ reaches the lanes (profiles/r02_ncu_full_summary.txt: of the
// 39.5 thread-instructions per sample of the chunk-ring kernel, 3.3 were staging — HBM -> registers
// -> 48 STS.32 per 32-sample chunk plus their address arithmetic — and 1.0 the raw LDS.32):
//   * the ring holds four PERIODS per row, not two 32-sample chunks: period t of the segment lives
//     in slot t & 3, 80 bytes per row, row pitch 84 words (4 x odd: the 128-bit reads of 32 lanes,
//     each in its own row, fall in distinct bank quads);
//   * a period is staged three periods ahead by ONE bulk (TMA-engine) copy per lane — its own
//     row's 80 bytes, cp.async.bulk global -> shared, no registers, no STS — completing on an
//     mbarrier per slot and warp (expect_tx by lane 0, try_wait on the slot's phase parity by
//     all); the slot it lands in was read one period ago;
//   * a lane reads its period as five LDS.128 at fixed offsets from the slot base: no ring
//     masking, no mirror, a quarter of the load instructions.
// The copies need the period's first sample on a 16-byte boundary in every row: the input base and
// the channel stride 16-byte aligned and (first sample index) % 4 == 0, which holds whenever the
// block lengths are multiples of 4 (dsCnt then only takes multiples of 4).  The host checks and
// falls back to the chunk-ring kernel otherwise (and for I/Q correction, float input, D = 10).
// MEASURED SLOWER than the chunk ring (5.03 against 3.90 ms per 2^31 samples in the pump; 4.60 ms
// with five 16-byte cp.async copies per lane instead of the bulk copy): 107 M copies of 80 bytes per
// step are the wrong granularity for the copy engines.  It is therefore opt-in
// (jsdr_bpsk_set_kernel(JSDR_KERNEL_PRING) / JSDR_PRING=1) and kept as the record of that experiment
// (DESIGN.md section 9), bit-identical to the default kernel and tested as such.
// Periods that reach outside the block (history before sample 0, the zero tail after the last
// sample) are staged synchronously with bounds checks and run the sample-by-sample body.
#pragma once

namespace jsdr {
namespace bpsk {
namespace stream {

constexpr int kPSlots = 4;                                   // periods per row in the ring
template <int DD> __host__ __device__ constexpr int p_pitch() { return kPSlots * DD + 4; }   // words; (pitch/4) odd for DD = 20
constexpr int kPWarps = 16;

template <int DD>
constexpr size_t p_smem_bytes()
{
    return (size_t)kTabBytes + (size_t)kPWarps * (32 * p_pitch<DD>() * 4 + kPSlots * 8);
}

// ---- bulk (TMA-engine) copies tracked by an mbarrier: one 16-byte-aligned row segment per lane
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ uint4 lds_u4(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// (I, Q) of one raw s16 pair as the two doubles the reference mixes: (double)((float)s / 32767f)
// (JavaAudio.java:283, FUNcubeBPSKDemod.java:372-373), both halves at once in packed FP32
__device__ __forceinline__ void raw_to_doubles(uint32_t w, double &xi, double &xq)
{
    w ^= 0x80008000u;              // exact s16 -> float: splice into the mantissa of 2^23
    const float2 f2 = add2(make_float2(__uint_as_float(__byte_perm(w, 0x4b000000u, 0x7410)),
                                       __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7432))),
                           make_float2(-8421376.0f, -8421376.0f));
    // correctly rounded quotient: fma(s, r_hi, s*r_lo), see s16_over_32767 in bpsk.cu
    const float2 q2 = fma2(f2, make_float2(3.0518509447574615e-05f, 3.0518509447574615e-05f),
                           mul2(f2, make_float2(2.8422576792141996e-14f, 2.8422576792141996e-14f)));
    xi = (double)q2.x;
    xq = (double)q2.y;
}

template <int PREC, int NTAPS, int DD>
__global__ void __launch_bounds__(kPWarps * 32, 1) k_mixdecim_pring(const Params p)
{
    constexpr int NQ = (NTAPS + DD - 1) / DD;
    constexpr int PITCH = p_pitch<DD>();                 // words
    constexpr int H = NTAPS - 1;
    constexpr int QPP = DD / 4;                          // 128-bit reads per lane per period
    constexpr int AHEAD = kPSlots - 1;                   // periods staged ahead of the one in use
    static_assert(DD % 4 == 0 && DD <= 32, "period");
    static_assert(NTAPS <= kMaxTaps && NQ <= 4, "taps");

    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if constexpr (PREC == PREC_F64) {
        double2 *tab = reinterpret_cast<double2 *>(smem);
        for (int i = tid; i < 257 * 8; i += kPWarps * 32) tab[i] = p.cossin[i >> 3];
    } else {
        float2 *tab = reinterpret_cast<float2 *>(smem);
        for (int i = tid; i < 257 * 16; i += kPWarps * 32) {
            const double2 cs = p.cossin[i >> 4];
            tab[i] = make_float2((float)(cs.x / 32767.0), (float)(cs.y / 32767.0));
        }
    }
    __syncthreads();
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned lane_tab = smem_base + ((PREC == PREC_F64) ? (lane & 7) * 16 : (lane & 15) * 8);
    constexpr int WARP_BYTES = 32 * PITCH * 4 + kPSlots * 8;
    const unsigned ring = smem_base + kTabBytes + warp * WARP_BYTES;             // this warp's 32 rows
    const unsigned myrow = ring + lane * (PITCH * 4);
    const unsigned mbar0 = ring + 32 * PITCH * 4;                                // one mbarrier per slot
    uint32_t *ring_g = reinterpret_cast<uint32_t *>(smem + kTabBytes + warp * WARP_BYTES);
    if (lane < kPSlots) mbar_init(mbar0 + lane * 8, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    unsigned use_count = 0;                              // periods consumed by this warp so far (mbarrier phase)

    for (;;) {
        int wg = 0;
        if (lane == 0) wg = (int)atomicAdd(p.work_counter, 1u);
        wg = __shfl_sync(0xffffffffu, wg, 0);
        if (wg >= p.ncw * p.nseg) break;
        const int cw = wg % p.ncw, seg = wg / p.ncw;
        const int ch0 = cw * 32;
        const int rows = min(32, p.nchan - ch0);
        const int ch = min(ch0 + lane, p.nchan - 1);
        const bool lane_live = ch0 + lane < p.nchan;
        const int m_top = (seg + 1) * p.R - 1;
        const int n_top = p.n0 + m_top * DD;
        const int nper = p.R + NQ - 1;
        const unsigned long long dx = p.tu_dx56[ch];
        const bool exact_lane = (dx == 0ull);
        const uint32_t *in = reinterpret_cast<const uint32_t *>(p.in);

        // ---- staging of period t (samples s_hi-DD+1 .. s_hi) into slot t & 3
        const unsigned use_base = use_count;             // slots and mbarrier phases run on across items
        // row pointer of this lane for the bulk copies (row = lane)
        const uint32_t *my_src = in + (long long)(ch0 + min(lane, rows - 1)) * p.chan_stride;
        auto stage = [&](int t) {
            if (t >= nper) return;
            const int s_lo = n_top - t * DD - (DD - 1);
            const unsigned slot = (use_base + (unsigned)t) & (kPSlots - 1);
            const unsigned mbar = mbar0 + slot * 8;
            if (s_lo >= 0 && s_lo + DD <= p.S && rows == 32) {
                if (lane == 0) mbar_arrive_expect_tx(mbar, 32u * DD * 4u);
                bulk_g2s(myrow + slot * (DD * 4), my_src + s_lo, DD * 4, mbar);
            } else {                                     // block edges, partial channel groups: with bounds checks
                for (int q = lane; q < 32 * DD; q += 32) {
                    const int row = q / DD, e = q - row * DD, s = s_lo + e;
                    uint32_t v = 0u;
                    if (row < rows && s >= 0 && s < p.S) v = in[(long long)(ch0 + row) * p.chan_stride + s];
                    ring_g[row * PITCH + slot * DD + e] = v;
                }
                __syncwarp();                            // the stores above are visible to the warp before the arrival
                if (lane == 0) mbar_arrive(mbar);
            }
        };

        Acc<PREC> acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) acc[q].zero();
        constexpr int kReanchor = 12;
        const unsigned long long dxD = dx * (unsigned long long)DD;
        unsigned long long x_top = 0;
        bool bad_anchor = false;

        __syncwarp();                                    // the previous item's reads are done
#pragma unroll
        for (int t = 0; t < AHEAD; t++) stage(t);

#pragma unroll 1
        for (int tt = 0; tt < nper; tt++) {
            const int s_hi = n_top - tt * DD;
            const int s_lo = s_hi - DD + 1;
            __syncwarp();                                // the slot of period tt + AHEAD is the one read last period
            stage(tt + AHEAD);
            // this period's samples have landed (the mbarrier's phase flips once per use of a slot);
            // the five 128-bit reads are issued before the table addresses so that they overlap them
            const unsigned uidx = use_base + (unsigned)tt;
            const unsigned slot_addr = myrow + (uidx & (kPSlots - 1)) * (DD * 4);
            mbar_wait(mbar0 + (uidx & (kPSlots - 1)) * 8, (uidx / kPSlots) & 1u);
            uint32_t rw[DD];
            if (s_lo >= 0) {
#pragma unroll
                for (int k = 0; k < QPP; k++) {
                    const uint4 v = lds_u4(slot_addr + k * 16);
                    rw[4 * k + 0] = v.x; rw[4 * k + 1] = v.y; rw[4 * k + 2] = v.z; rw[4 * k + 3] = v.w;
                }
            }
            // ---- table addresses of the period's samples (as in k_mixdecim_stream)
            unsigned taddr[DD];
            unsigned xt, ndx;
            bool near;
            uint16_t ix[DD];
            {
                if (tt % kReanchor == 0) {
                    const int c = min(max(s_hi >> 5, 0), p.nchunks - 1);
                    const double ck = p.ckpt[(size_t)c * p.nchan + ch];
                    x_top = phase_to_x56(ck) + (unsigned long long)((long long)(s_hi - (c << 5) + 1)) * dx;
                    const int c_lo = min(max((s_hi - kReanchor * DD + 1) >> 5, 0), p.nchunks - 1);
                    const double ck_lo = p.ckpt[(size_t)c_lo * p.nchan + ch];
                    bad_anchor = !(ck >= 0.0) || !(ck_lo >= 0.0);
                }
                xt = (unsigned)(x_top >> 32);
                ndx = 0u - (unsigned)(dx >> 32);
                const unsigned nt = (xt << 8) + 256u, ndx8 = ndx << 8;
                x_top -= dxD;
                near = exact_lane || bad_anchor;
#pragma unroll
                for (int j = 0; j < DD; j++) {
                    const unsigned xj = xt + (unsigned)j * ndx;
                    taddr[j] = lane_tab + ((xj >> 17) & 0x7f80u);
                    near |= (nt + (unsigned)j * ndx8) <= 256u * (unsigned)(DD + 2);
                }
                if (__any_sync(0xffffffffu, near)) {
                    if (near) {
                        exact_period_indices(p.ckpt, p.tu_inc[ch], p.nchan, p.S, ch, s_hi, DD, ix);
#pragma unroll
                        for (int j = 0; j < DD; j++) taddr[j] = lane_tab + ((unsigned)ix[j] << 7);
                    }
                }
            }
            if (s_lo >= 0) {
                // ---- every sample of the period is a sample of this call
#pragma unroll
                for (int j = 0; j < DD; j++) {             // newest first: sample s_hi - j is element DD-1-j
                    const uint32_t w = rw[DD - 1 - j];
                    if constexpr (PREC == PREC_F64) {
                        double xi, xq;
                        raw_to_doubles(w, xi, xq);
                        const double2 cs = lds_d2(taddr[j]);
                        const double mi = __dmul_rn(xi, cs.x);   // :389-390 i*cosTab[ix], q*sinTab[ix]
                        const double mq = __dmul_rn(xq, cs.y);
#pragma unroll
                        for (int q = 0; q < NQ; q++) {
                            const int k = j + q * DD;
                            if (k < NTAPS) {
                                acc[q].i = __dadd_rn(acc[q].i, __dmul_rn(mi, p.taps[k]));   // :479-483, age order
                                acc[q].q = __dadd_rn(acc[q].q, __dmul_rn(mq, p.taps[k]));
                            }
                        }
                    } else {
                        const uint32_t u = w ^ 0x80008000u;
                        const float2 f2 = add2(make_float2(__uint_as_float(__byte_perm(u, 0x4b000000u, 0x7410)),
                                                           __uint_as_float(__byte_perm(u, 0x4b000000u, 0x7432))),
                                               make_float2(-8421376.0f, -8421376.0f));
                        const float2 cs = lds_f2(taddr[j]);
                        const float mi = f2.x * cs.x, mq = f2.y * cs.y;
#pragma unroll
                        for (int q = 0; q < NQ; q++) {
                            const int k = j + q * DD;
                            if (k < NTAPS) {
                                acc[q].i = fmaf(mi, p.tapsf[k], acc[q].i);
                                acc[q].q = fmaf(mq, p.tapsf[k], acc[q].q);
                            }
                        }
                    }
                }
            } else {
                // ---- the period reaches before sample 0: those samples are the carried history
                const uint32_t *srow = ring_g + lane * PITCH + (uidx & (kPSlots - 1)) * DD;
#pragma unroll 1
                for (int j = 0; j < DD; j++) {
                    const int s = s_hi - j;
                    // (dynamic j: the address is recomputed rather than read from taddr[], which would
                    // otherwise have to live in local memory for every period)
                    const unsigned ta = near ? lane_tab + ((unsigned)ix[j] << 7)
                                             : lane_tab + (((xt + (unsigned)j * ndx) >> 17) & 0x7f80u);
                    double2 hv = make_double2(0.0, 0.0);
                    if (s < 0 && s + H >= 0) hv = p.hist_in[(size_t)ch * kMaxDsTaps + s + H];
                    const uint32_t w = srow[DD - 1 - j];
                    if constexpr (PREC == PREC_F64) {
                        double mi = hv.x, mq = hv.y;
                        if (s >= 0) {
                            double xi, xq;
                            raw_to_doubles(w, xi, xq);
                            const double2 cs = lds_d2(ta);
                            mi = __dmul_rn(xi, cs.x);
                            mq = __dmul_rn(xq, cs.y);
                        }
#pragma unroll
                        for (int q = 0; q < NQ; q++) {
                            const int k = j + q * DD;
                            if (k < NTAPS) {
                                acc[q].i = __dadd_rn(acc[q].i, __dmul_rn(mi, p.taps[k]));
                                acc[q].q = __dadd_rn(acc[q].q, __dmul_rn(mq, p.taps[k]));
                            }
                        }
                    } else {
                        float mi = (float)hv.x, mq = (float)hv.y;
                        if (s >= 0) {
                            const uint32_t u = w ^ 0x80008000u;
                            const float2 f2 = add2(make_float2(__uint_as_float(__byte_perm(u, 0x4b000000u, 0x7410)),
                                                               __uint_as_float(__byte_perm(u, 0x4b000000u, 0x7432))),
                                                   make_float2(-8421376.0f, -8421376.0f));
                            const float2 cs = lds_f2(ta);
                            mi = f2.x * cs.x;
                            mq = f2.y * cs.y;
                        }
#pragma unroll
                        for (int q = 0; q < NQ; q++) {
                            const int k = j + q * DD;
                            if (k < NTAPS) {
                                acc[q].i = fmaf(mi, p.tapsf[k], acc[q].i);
                                acc[q].q = fmaf(mq, p.tapsf[k], acc[q].q);
                            }
                        }
                    }
                }
            }

            // ---- the oldest role is complete: scale (:469,486), store, rotate
            const int r_out = tt - (NQ - 1);
            const int m = m_top - r_out;
            if (r_out >= 0 && m < p.NO && lane_live) {
                double2 o;
                if constexpr (PREC == PREC_F64)
                    o = make_double2(__dmul_rn(acc[NQ - 1].i, 0.9 * 32768.0), __dmul_rn(acc[NQ - 1].q, 0.9 * 32768.0));
                else
                    o = make_double2((double)acc[NQ - 1].i * (0.9 * 32768.0), (double)acc[NQ - 1].q * (0.9 * 32768.0));
                p.ds_out[(size_t)ch * p.max_ds + m] = o;
            }
#pragma unroll
            for (int q = NQ - 1; q > 0; q--) acc[q] = acc[q - 1];
            acc[0].zero();
        }
        use_count += (unsigned)nper;
    }
}

}  // namespace stream
}  // namespace bpsk
}  // namespace jsdr
