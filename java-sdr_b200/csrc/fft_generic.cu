// fft_generic.cu — see fft_generic.cuh.
#include "fft_generic.cuh"

namespace jsdr {
namespace fftg {

StagePlan make_plan(int n)
{
    StagePlan p;
    p.nstages = 0;
    if (n < 1) return p;
    int r = n;
    const int cand[5] = {4, 2, 3, 5, 7};
    for (int c : cand)
        while (r % c == 0 && p.nstages < 24) {
            p.radix[p.nstages++] = c;
            r /= c;
        }
    if (r != 1) p.nstages = 0;
    if (n == 1) {   // a single identity "stage" keeps the callers uniform
        p.nstages = 1;
        p.radix[0] = 1;
    }
    return p;
}

namespace {

__device__ __forceinline__ void sincos_turns(double x, double *s, double *c) { sincospi(2.0 * x, s, c); }
__device__ __forceinline__ void sincos_turns(float x, float *s, float *c) { sincospif(2.0f * x, s, c); }

// One Stockham stage of radix R over sub-transforms of length Ns*R.  Thread = one
// butterfly j of one block: inputs in[j + t*N/R], twiddles exp(sign*2*pi*i*t*k/(Ns*R)) with
// k = j mod Ns, outputs out[(j-k)*R + k + t*Ns].
template <typename T, int R>
__global__ void __launch_bounds__(256) k_stage(const typename Cx<T>::type *__restrict__ in,
                                               typename Cx<T>::type *__restrict__ out, int N, int Ns, long long nbfly,
                                               int sign)
{
    typedef typename Cx<T>::type C;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nbfly) return;
    const int M = N / R;
    const long long blk = gid / M;
    const int j = (int)(gid - blk * M);
    const int k = j % Ns;
    const C *src = in + blk * N + j;
    C v[R];
#pragma unroll
    for (int t = 0; t < R; t++) v[t] = src[(long long)t * M];
    if (Ns > 1) {
        const int L = Ns * R;
#pragma unroll
        for (int t = 1; t < R; t++) {
            // t*k < L*R: reduce on the integers so the angle is exact before it is rounded
            const T turns = (T)((t * k) % L) / (T)L;
            T s, c;
            sincos_turns(turns, &s, &c);
            s = (sign < 0) ? -s : s;
            const C a = v[t];
            v[t].x = a.x * c - a.y * s;
            v[t].y = a.x * s + a.y * c;
        }
    }
    C o[R];
    if (R == 1) {
        o[0] = v[0];
    } else if (R == 2) {
        o[0].x = v[0].x + v[1].x; o[0].y = v[0].y + v[1].y;
        o[1].x = v[0].x - v[1].x; o[1].y = v[0].y - v[1].y;
    } else if (R == 4) {
        const C a0 = {v[0].x + v[2].x, v[0].y + v[2].y}, a1 = {v[0].x - v[2].x, v[0].y - v[2].y};
        const C a2 = {v[1].x + v[3].x, v[1].y + v[3].y}, d = {v[1].x - v[3].x, v[1].y - v[3].y};
        // forward: -i*d = (d.y, -d.x); inverse: +i*d = (-d.y, d.x)
        const C a3 = (sign < 0) ? C{d.y, -d.x} : C{-d.y, d.x};
        o[0].x = a0.x + a2.x; o[0].y = a0.y + a2.y;
        o[1].x = a1.x + a3.x; o[1].y = a1.y + a3.y;
        o[2].x = a0.x - a2.x; o[2].y = a0.y - a2.y;
        o[3].x = a1.x - a3.x; o[3].y = a1.y - a3.y;
    } else {
        // small odd radix: direct sum with the R-th roots of unity
        T wr[R], wi[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            T s, c;
            sincos_turns((T)q / (T)R, &s, &c);
            wr[q] = c;
            wi[q] = (sign < 0) ? -s : s;
        }
#pragma unroll
        for (int q = 0; q < R; q++) {
            T ar = v[0].x, ai = v[0].y;
#pragma unroll
            for (int t = 1; t < R; t++) {
                const int e = (t * q) % R;
                ar += v[t].x * wr[e] - v[t].y * wi[e];
                ai += v[t].x * wi[e] + v[t].y * wr[e];
            }
            o[q].x = ar;
            o[q].y = ai;
        }
    }
    C *dst = out + blk * N + (long long)(j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; t++) dst[(long long)t * Ns] = o[t];
}

template <typename T, int R>
int launch_stage(jsdr_ctx *ctx, cudaStream_t st, const typename Cx<T>::type *in, typename Cx<T>::type *out, int n,
                 int Ns, long long batch, int sign)
{
    const long long nbfly = batch * (n / R);
    const long long grid = (nbfly + 255) / 256;
    if (grid <= 0) return JSDR_OK;
    if (grid > 0x7fffffffLL) {
        set_error("fft_generic: batch too large for one launch");
        return JSDR_EINVAL;
    }
    ProfScope prof(ctx, JSDR_K_FFT, st);
    k_stage<T, R><<<(unsigned)grid, 256, 0, st>>>(in, out, n, Ns, nbfly, sign);
    return launched(ctx, "fftg::k_stage");
}

}  // namespace

template <typename T>
typename Cx<T>::type *run(jsdr_ctx *ctx, cudaStream_t st, typename Cx<T>::type *a, typename Cx<T>::type *b, int n,
                         long long batch, int sign, int *rc)
{
    *rc = JSDR_OK;
    const StagePlan p = make_plan(n);
    if (p.nstages == 0) {
        set_error("fft_generic: n=%d has a prime factor other than 2, 3, 5, 7", n);
        *rc = JSDR_EUNSUPPORTED;
        return nullptr;
    }
    typename Cx<T>::type *src = a, *dst = b;
    int Ns = 1;
    for (int i = 0; i < p.nstages && *rc == JSDR_OK; i++) {
        const int R = p.radix[i];
        switch (R) {
        case 1: *rc = launch_stage<T, 1>(ctx, st, src, dst, n, Ns, batch, sign); break;
        case 2: *rc = launch_stage<T, 2>(ctx, st, src, dst, n, Ns, batch, sign); break;
        case 3: *rc = launch_stage<T, 3>(ctx, st, src, dst, n, Ns, batch, sign); break;
        case 4: *rc = launch_stage<T, 4>(ctx, st, src, dst, n, Ns, batch, sign); break;
        case 5: *rc = launch_stage<T, 5>(ctx, st, src, dst, n, Ns, batch, sign); break;
        default: *rc = launch_stage<T, 7>(ctx, st, src, dst, n, Ns, batch, sign); break;
        }
        Ns *= R;
        typename Cx<T>::type *t = src;
        src = dst;
        dst = t;
    }
    return (*rc == JSDR_OK) ? src : nullptr;
}

template float2 *run<float>(jsdr_ctx *, cudaStream_t, float2 *, float2 *, int, long long, int, int *);
template double2 *run<double>(jsdr_ctx *, cudaStream_t, double2 *, double2 *, int, long long, int, int *);

}  // namespace fftg
}  // namespace jsdr
