// Shared device utilities: warp shuffles of multi-double elements, coalesced tile staging with a
// conflict-free transposed shared-memory layout, block reductions, and the chained
// ("decoupled look-back") tile-prefix protocol used by the forward and adjoint scans.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "models.cuh"

namespace ssde {

constexpr unsigned FULL = 0xffffffffu;

// row flags (uint8 per row)
enum : uint8_t {
    ROW_START = 1,     // first row of a track (ID(i) != ID(i-1)); never set for a continued shard
    ROW_LAST = 2,      // last row of a track: its prediction / outgoing transition is discarded
    ROW_OBS = 4,       // CTCRW: obs(i,0) is not NA (nllk_ctcrw.hpp:214 tests column 0 only)
    ROW_NA0 = 8,       // BM/OU: dimension d of this row is NA  -> bit (3 + d)
};

// ---------------------------------------------------------------------------------------------
// elements as arrays of doubles
// ---------------------------------------------------------------------------------------------
template <class E>
__device__ __forceinline__ E shfl_down_elem(const E& e, int delta) {
    E r;
    const double* s = reinterpret_cast<const double*>(&e);
    double* d = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < E::NDBL; ++i) d[i] = __shfl_down_sync(FULL, s[i], delta);
    return r;
}
template <class E>
__device__ __forceinline__ E shfl_up_elem(const E& e, int delta) {
    E r;
    const double* s = reinterpret_cast<const double*>(&e);
    double* d = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < E::NDBL; ++i) d[i] = __shfl_up_sync(FULL, s[i], delta);
    return r;
}
template <class E>
__device__ __forceinline__ E shfl_idx_elem(const E& e, int src) {
    E r;
    const double* s = reinterpret_cast<const double*>(&e);
    double* d = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < E::NDBL; ++i) d[i] = __shfl_sync(FULL, s[i], src);
    return r;
}
template <class E>
__device__ __forceinline__ void store_elem(double* dst, const E& e) {
    const double* s = reinterpret_cast<const double*>(&e);
#pragma unroll
    for (int i = 0; i < E::NDBL; ++i) dst[i] = s[i];
}
template <class E>
__device__ __forceinline__ E load_elem(const double* src) {
    E r;
    double* d = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < E::NDBL; ++i) d[i] = src[i];
    return r;
}
// L2-coherent (L1-bypassing) load of an element published by another CTA
template <class E>
__device__ __forceinline__ E load_elem_cg(const double* src) {
    E r;
    double* d = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < E::NDBL; ++i) d[i] = __ldcg(src + i);
    return r;
}

// Time-ordered composition traits.  join(far, near): `near` covers tiles closer (in processing
// order) to the current one.  Forward scan processes tiles in time order, so far = earlier in
// time; the adjoint scan processes tiles in reverse time, so far = later in time.
// An element whose linear part has decayed below CONST_MAP_TOL (models.cuh) acts as a constant
// map: what comes before it cannot change the result by more than that factor (40+ orders of
// magnitude below fp64 rounding), so a look-back may stop there as if it had found an inclusive
// prefix.  With observations on most rows the filter forgets its initial condition within a few
// hundred rows (A underflows to exactly 0), which makes almost every tile aggregate a constant map.
template <class M>
struct FwdOps {
    using Elem = typename M::FwdElem;
    static __device__ __forceinline__ Elem identity() { return M::fwd_identity(); }
    static __device__ __forceinline__ Elem join(const Elem& far, const Elem& near) { return M::fwd_combine(far, near); }
    static __device__ __forceinline__ bool is_const(const Elem& e) { return M::fwd_is_const(e); }
};
template <class M>
struct BwdOps {
    using Elem = typename M::BwdElem;
    static __device__ __forceinline__ Elem identity() { return M::bwd_identity(); }
    static __device__ __forceinline__ Elem join(const Elem& far, const Elem& near) {
        return M::bwd_combine(near, far);      // near = earlier rows (E1), far = later rows (E2)
    }
    static __device__ __forceinline__ bool is_const(const Elem& e) { return M::bwd_is_const(e); }
};

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ Dual warp_sum(Dual v) {
    return Dual(warp_sum(v.v), warp_sum(v.d));
}

// Sum over a block of NT threads; result valid in thread 0.  `red` has NT/32 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) t += red[i];
    }
    return t;
}

// ---------------------------------------------------------------------------------------------
// tile staging: rows [r0, r0 + NT*LC) x NC doubles (row-major in global memory) are loaded with
// coalesced 16-byte loads and stored transposed so that thread t reads component c of its k-th
// row at  s[(k*NC + c)*(NT + 1) + t]  without bank conflicts.
// ---------------------------------------------------------------------------------------------
template <int NC, int NT, int LC>
struct Staged {
    static constexpr int STRIDE = NT + 1;
    static constexpr int SIZE = LC * NC * STRIDE;       // doubles
    static __device__ __forceinline__ int at(int k, int c, int t) { return (k * NC + c) * STRIDE + t; }
};

template <int NC, int NT, int LC>
__device__ __forceinline__ void stage_rows(double* s, const double* __restrict__ g, int64_t r0,
                                           int64_t n) {
    using L = Staged<NC, NT, LC>;
    constexpr int TOTAL = NT * LC * NC;                 // doubles in a full tile
    const double* base = g + r0 * NC;
    int64_t valid = (n - r0) * NC;
    if (valid > TOTAL) valid = TOTAL;
    if ((TOTAL % 2 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
        const double2* b2 = reinterpret_cast<const double2*>(base);
        for (int i = threadIdx.x; i < TOTAL / 2; i += NT) {
            double2 v = make_double2(0.0, 0.0);
            const int e0 = 2 * i;
            if (e0 + 1 < valid) v = __ldg(b2 + i);
            else if (e0 < valid) v.x = __ldg(base + e0);
            int r = e0 / NC, c = e0 % NC;
            s[L::at(r % LC, c, r / LC)] = v.x;
            r = (e0 + 1) / NC; c = (e0 + 1) % NC;
            s[L::at(r % LC, c, r / LC)] = v.y;
        }
    } else {
        for (int i = threadIdx.x; i < TOTAL; i += NT) {
            const double v = (i < valid) ? __ldg(base + i) : 0.0;
            const int r = i / NC, c = i % NC;
            s[L::at(r % LC, c, r / LC)] = v;
        }
    }
}

// uint8 flags of a tile: s[k*(NT+4) + t]   (row = t*LC + k)
template <int NT, int LC>
__device__ __forceinline__ void stage_flags(uint8_t* s, const uint8_t* __restrict__ g, int64_t r0,
                                            int64_t n) {
    constexpr int TOTAL = NT * LC;
    for (int i = threadIdx.x; i < TOTAL; i += NT) {
        const uint8_t v = (r0 + i < n) ? g[r0 + i] : (uint8_t)0xff;   // 0xff = row beyond the end
        s[(i % LC) * (NT + 4) + (i / LC)] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// chained tile prefix (decoupled look-back, Merrill & Garland) with multi-double payloads.
//   status word = (epoch << 2) | code,  code: 1 = aggregate available, 2 = inclusive prefix
//   available, 3 = aggregate available and it is a constant map (serves as an inclusive prefix).  A word whose epoch differs from the current launch's counts as "not ready", so
//   the descriptor arrays never need clearing between evaluations.
// ---------------------------------------------------------------------------------------------
struct ScanDesc {
    unsigned* status;      // [ntiles]
    double* agg;           // [ntiles * NDBL]
    double* incl;          // [ntiles * NDBL]
    unsigned* ticket;      // dynamic tile counter (zeroed by the host before every launch)
    unsigned* error;       // bit 0: a spin-wait timed out; bit 1: innovation variance F <= 0 (forward re-run)
    unsigned epoch;
#ifdef SSDE_STATS
    unsigned long long* stats;   // diagnostics build only: [0] look-backs, [1] windows, [2] spins, [3] cycles
#endif
};

__device__ __forceinline__ unsigned ld_status(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Publish the aggregate of tile `t` (one thread).
template <class Ops>
__device__ __forceinline__ void publish_agg(const ScanDesc& d, int t, const typename Ops::Elem& e) {
    store_elem(d.agg + (size_t)t * Ops::Elem::NDBL, e);
    st_status(d.status + t, (d.epoch << 2) | (Ops::is_const(e) ? 3u : 1u));   // release: orders the element stores before it
}
template <class Ops>
__device__ __forceinline__ void publish_incl(const ScanDesc& d, int t, const typename Ops::Elem& e) {
    store_elem(d.incl + (size_t)t * Ops::Elem::NDBL, e);
    st_status(d.status + t, (d.epoch << 2) | 2u);
}

// Executed by one full warp.  Returns (in every lane) the composite of all tiles processed
// before ticket `t` (identity for t == 0).  Lane l inspects tile base - l; the window is consumed
// as soon as every tile up to the nearest one that already knows its inclusive prefix has at
// least published its aggregate -- tiles farther away are never waited for.
// `lo`: first ticket of this launch (tickets below it count as the identity prefix; a tail-only
// summary pass starts in the middle of the tile sequence).
template <class Ops>
__device__ __forceinline__ typename Ops::Elem lookback(const ScanDesc& d, int t, int lo = 0) {
    using Elem = typename Ops::Elem;
    const int lane = threadIdx.x & 31;
    Elem acc = Ops::identity();
    bool have = false;                                     // acc still is the identity: no need to join
    int base = t - 1;
#ifdef SSDE_STATS
    const long long c0 = clock64();
    unsigned long long n_win = 0, n_spin = 0;
    struct Fin { const ScanDesc& d; const long long c0; unsigned long long &w, &s; int lane;
                 __device__ ~Fin() { if (lane == 0) { atomicAdd(d.stats + 0, 1ull); atomicAdd(d.stats + 1, w); atomicAdd(d.stats + 2, s);
                                                       atomicAdd(d.stats + 3, (unsigned long long)(clock64() - c0)); } } } fin{d, c0, n_win, n_spin, lane};
#endif
    while (base >= lo) {
        const int idx = base - lane;
        unsigned spins = 0;
#ifdef SSDE_STATS
        ++n_win;
#endif
        int first;                                         // nearest lane holding an inclusive prefix
        unsigned code;
        while (true) {
            code = 2u;                                     // virtual tile lo - 1: identity prefix
            if (idx >= lo) {
                const unsigned st = ld_status(d.status + idx);
                code = ((st >> 2) == d.epoch) ? (st & 3u) : 0u;
            }
            const unsigned incl = __ballot_sync(FULL, code >= 2u);
            const unsigned none = __ballot_sync(FULL, code == 0u);
            first = incl ? (__ffs(incl) - 1) : 32;
            const unsigned need = (first >= 31) ? FULL : ((2u << first) - 1u);
            if ((none & need) == 0u) break;
#ifdef SSDE_STATS
            ++n_spin;
#endif
            if (++spins > (1u << 22)) {                    // ~seconds: give up instead of hanging
                if (lane == 0) atomicOr(d.error, 1u);
                return acc;
            }
            __nanosleep(20);
        }
        Elem e = Ops::identity();
        if (idx >= lo) {
            if (lane < first || (lane == first && code == 3u)) e = load_elem_cg<Elem>(d.agg + (size_t)idx * Elem::NDBL);
            else if (lane == first) e = load_elem_cg<Elem>(d.incl + (size_t)idx * Elem::NDBL);
        }
        // ordered tree reduction: higher lanes are farther away; lanes beyond `first` hold the
        // identity, so only ceil(log2(first + 1)) levels are needed
#pragma unroll 1
        for (int o = 1; o < 32 && o <= first; o <<= 1) {
            Elem f = shfl_down_elem(e, o);
            if (lane + o < 32) e = Ops::join(f, e);
        }
        e = shfl_idx_elem(e, 0);
        acc = have ? Ops::join(e, acc) : e;
        have = true;
        if (first < 32) break;
        base -= 32;
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier): one lane of a warp
// fetches the next contiguous block of its warp-tile while the warp computes on the current one.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arrive (count 1) and announce `bytes` of asynchronous traffic for the current phase
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// hint: bring [p, p + bytes) into L2 (bytes a multiple of 16)
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// order this thread's earlier generic-proxy accesses to shared memory before later async-proxy
// (TMA) writes to the same locations
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace ssde
