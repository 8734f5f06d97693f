// common.cuh — shared plumbing for libjsdrcuda.so (context, error reporting,
// streaming load/store helpers).  sm_100a only; there is no CPU fallback.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <exception>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/jsdrcuda.h"

namespace jsdr {

void set_error(const char *fmt, ...);

#define JSDR_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            jsdr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,             \
                            cudaGetErrorString(_e));                                 \
            cudaGetLastError(); /* reported: the next launch check must not see it */ \
            return _e == cudaErrorMemoryAllocation ? JSDR_ENOMEM : JSDR_ECUDA;       \
        }                                                                            \
    } while (0)

#define JSDR_REQUIRE(cond, code, msg)                                                \
    do {                                                                             \
        if (!(cond)) {                                                               \
            jsdr::set_error("%s: %s", __func__, msg);                                \
            return code;                                                             \
        }                                                                            \
    } while (0)

#define JSDR_TRY(expr)                                                               \
    do {                                                                             \
        int _rc = (expr);                                                            \
        if (_rc != JSDR_OK) return _rc;                                              \
    } while (0)

// Every entry point is a function-try-block closed by this: the header promises that nothing
// throws across the C boundary (a C++ exception unwinding into a JVM frame ends the process), so a
// failed host allocation (new, std::vector) comes back as a status like every other failure.
#define JSDR_CATCH_ALL                                                               \
    catch (const std::bad_alloc &)                                                   \
    {                                                                                \
        jsdr::set_error("%s: out of host memory", __func__);                         \
        return JSDR_ENOMEM;                                                          \
    }                                                                                \
    catch (const std::exception &e_)                                                 \
    {                                                                                \
        jsdr::set_error("%s: %s", __func__, e_.what());                              \
        return JSDR_EINTERNAL;                                                       \
    }                                                                                \
    catch (...)                                                                      \
    {                                                                                \
        jsdr::set_error("%s: unknown C++ exception", __func__);                      \
        return JSDR_EINTERNAL;                                                       \
    }

}  // namespace jsdr

struct jsdr_prof_span {
    int kind;
    cudaEvent_t a, b;
};

struct jsdr_ctx {
    int device = 0;
    int sm_count = 0;
    int l2_prefetch = 1;             // FFT: bulk-prefetch the input of the CTA n x (resident CTAs) ahead into L2 (JSDR_L2_PREFETCH=n, 0: off)
    cudaStream_t stream = nullptr;   // main stream: data kernels
    cudaStream_t side = nullptr;     // side stream: data-independent phase scouts
    cudaStream_t side2 = nullptr;    // the VCO / bit-phase replay: one short serial chain, beside the tuner scout
    cudaStream_t aux = nullptr;      // low priority: work that runs beside the main stream's kernel
    cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host<->device copies of the chunked host path
    cudaEvent_t ev_chunk_in[16] = {nullptr}, ev_chunk_done[16] = {nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int64_t launches = 0;
    // per-kernel timing: event pairs recorded around the launches while `profiling` is on
    bool profiling = false;
    std::vector<jsdr_prof_span> spans;       // recorded since the last read
    std::vector<jsdr_prof_span> free_spans;  // event pairs ready for reuse

    int bind() const { return cudaSetDevice(device) == cudaSuccess ? JSDR_OK : JSDR_ECUDA; }
};

namespace jsdr {

// Function attributes (dynamic shared memory limits, occupancy) are per device: state kept by
// a launch site is indexed by the context's device, so that contexts on different GPUs of one
// process each set their own.
constexpr int kMaxDevices = 64;
// Contexts on different GPUs may be driven from different threads at the same time (the header's
// threading rule), so first-use initialisation is serialised: whoever comes second waits until the
// first has FINISHED (a flag set before the work would let the second thread launch a kernel whose
// shared-memory limit or constant tables are not there yet).
struct PerDeviceFlag {
    std::mutex mu;
    std::atomic<bool> done[kMaxDevices];
    PerDeviceFlag()
    {
        for (auto &d : done) d.store(false, std::memory_order_relaxed);
    }
    template <class F>
    int once(int dev, F &&init)                            // init() -> JSDR status; repeated until it succeeds
    {
        const bool known = dev >= 0 && dev < kMaxDevices;  // unknown device: redo every time
        if (known && done[dev].load(std::memory_order_acquire)) return JSDR_OK;
        std::lock_guard<std::mutex> lock(mu);
        if (known && done[dev].load(std::memory_order_relaxed)) return JSDR_OK;
        const int rc = init();
        if (rc == JSDR_OK && known) done[dev].store(true, std::memory_order_release);
        return rc;
    }
};

// Check the launch that was just made and count it.
static inline int launched(jsdr_ctx *ctx, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return JSDR_ECUDA;
    }
    ctx->launches++;
    return JSDR_OK;
}

// Bracket one kernel launch with CUDA events on its stream when profiling is on.  At most
// kMaxProfSpans launches are kept between two reads (a caller that switches profiling on and never
// reads must not grow event pairs without bound); later launches are simply not timed.
constexpr size_t kMaxProfSpans = 1 << 16;
struct ProfScope {
    jsdr_ctx *ctx;
    cudaStream_t st;
    jsdr_prof_span sp;
    bool on;
    ProfScope(jsdr_ctx *c, int kind, cudaStream_t s)
        : ctx(c), st(s), on(c->profiling && c->spans.size() < kMaxProfSpans && c->spans.size() < c->spans.capacity())
    {
        if (!on) return;
        if (!c->free_spans.empty()) {
            sp = c->free_spans.back();
            c->free_spans.pop_back();
        } else if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) {
            cudaGetLastError();
            on = false;
            return;
        }
        sp.kind = kind;
        cudaEventRecord(sp.a, st);
    }
    ~ProfScope()
    {
        if (!on) return;
        cudaEventRecord(sp.b, st);
        ctx->spans.push_back(sp);
    }
};

// ---- device helpers -------------------------------------------------------
#ifdef __CUDACC__
// streaming (read-once) global loads: bypass L1 allocation
__device__ __forceinline__ float2 ldg_stream_f2(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f32(float *p, float v)
{
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ void stg_stream_f2(float2 *p, float2 v)
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y));
}
#endif

}  // namespace jsdr
