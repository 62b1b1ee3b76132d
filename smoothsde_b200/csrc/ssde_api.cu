// C ABI of the engine (include/smoothsde_b200.h): handle life-cycle, conversion of the
// reference's data list to the device layout, and the per-evaluation kernel pipeline.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <memory>
#include <thread>
#include <chrono>
#include <exception>
#include <sched.h>
#include <sys/mman.h>
#include <string>
#include <vector>

#include "../../include/smoothsde_b200.h"
#include "kernels_ctcrw.cuh"
#include "kernels_sde.cuh"
#include "kernels_sim.cuh"

using namespace ssde;

namespace {

thread_local std::string g_create_error;

constexpr int FWD_NT = KNT_DEFAULT, BWD_NT = KNT_DEFAULT;   // decoupled models; the coupled ones use M::KNT = 64
constexpr int PAD_ROWS = 1024;          // n_pad is a multiple of the largest tile (FWD_NT * LC)
static_assert(PAD_ROWS % (FWD_NT * LC) == 0 && PAD_ROWS % (BWD_NT * LC) == 0 && PAD_ROWS % (SDE_NT * LC) == 0, "tiles must divide the padding unit");
constexpr int RED_BLOCKS = 64;          // partial sums of the per-tile outputs

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            err = std::string(#expr) + ": " + cudaGetErrorString(e__);                       \
            return SSDE_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    bool owned = false;
    ~DevBuf() { if (owned && p) cudaFree(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

namespace {
inline bool is_kalman(int model) { return model == SSDE_CTCRW || model == SSDE_OU_SSM || model == SSDE_BM_SSM; }
inline int sde_par_count(int model, int n_dim) { return (model == SSDE_BM || model == SSDE_BM_SSM) ? n_dim + 1 : n_dim + 2; }
inline int state_means(int model, int n_dim) { return model == SSDE_CTCRW ? 2 * n_dim : n_dim; }   // columns of a0 / aest_all
}  // namespace

struct ssde_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int model = 0, n_dim = 0, n_par = 0;
    int64_t n = 0, n_pad = 0, nnz = 0;
    int n_tracks = 0;
    int p_fe = 0, p_re = 0, n_s = 0, npar = 0;
    int o_sig = -1, o_fe = 0, o_ll = 0, o_re = 0;
    int o_dec = -1, n_dec = 0;       // BM/OU decay models: log_decay[n_dec] sits between log_lambda and coeff_re (nllk_sde.hpp:42-45)
    DevBuf t_decay, dec_of_col, grad_decay;
    bool has_smooth = false, add_penalty = true, penalty_consts = false;
    int include_penalty = 1;
    int shard_flags = 0;
    int num_sms = 148;
    PriorCov P0{{1.0, 0.0, 10.0}, {0.0}};
    bool dense = false;              // coupled filter (user H_array or a P0 that is not of the default shape)
    DevBuf Hrow;                     // permuted planes of the packed H_array rows (dense && user H only)
    // design + data
    DevBuf desc, col, val, obs, dt, flags, track_starts, a0;
    DevBuf mu_cols, mu_zero;         // theta entries that feed the mu_d predictors; device flag "all of them are 0"
    int n_mu_cols = -1;              // -1: unknown (flag stays 0)
    // penalty (CSR of S, per-smooth offsets, constants)
    DevBuf S_rowptr, S_col, S_val, sm_off;
    double pen_const = 0.0;
    // work buffers
    DevBuf par, theta, grad_theta, wg, ckpt, tile_llk, tile_gh, block_llk, part, out, sb;
    DevBuf f_status, f_agg, f_incl, b_status, b_agg, b_incl, counters;   // counters: ticket_f, ticket_b, error
    DevBuf aest;
    DevBuf stats;                    // diagnostics build (SSDE_STATS) only
    DevBuf s_in, g_in;               // incoming state / adjoint of a time shard (2 n_dim + 3 doubles each)
    bool have_s_in = false, have_g_in = false;
    // tangent (Hessian-vector) pass: Dual-sized copies of the work buffers, allocated on first use
    bool tan_ready = false;
    DevBuf t_dir, t_theta_dot, t_grad_theta, t_wg, t_ckpt, t_tile_gh, t_f_agg, t_f_incl, t_b_agg, t_b_incl, t_out, t_hess;
    int grid_f2 = 0, grid_b2 = 0, grid_lp2 = 0;
    bool sde_stream = false;         // BM / OU: every warp-tile is uniform with <= SDE_SMAX slots -> sde_stream_kernel
    int grid_stream = 0;
    HotRanges hot{};                 // columns whose gradient entries get per-CTA shared-memory accumulators
    HessHot hess_hot{};              // ... and whose Hessian entries do (sde_hess_kernel)
    int grid_hess = 0;
    double* h_pinned = nullptr;      // pinned host staging: par in, out back
    unsigned epoch = 0;
    int ntiles_f = 0, ntiles_b = 0, grid_lp = 0, grid_f = 0, grid_b = 0;
    int64_t ntiles_lp = 0;
    int64_t nchunks = 0;
    int fs = 0, fwd_elem = 0, bwd_elem = 0;      // Kalman models: scalars of a state, doubles of the scan elements
    int last_launches = 0;
    bool timed = false;
    // optional per-kernel timing (ssde_set_profile): event k is recorded before kernel k
    bool profile = false;
    std::vector<cudaEvent_t> pev;
    std::vector<const char*> pnames;
    int pcount = 0;
    std::string err;

    ~ssde_handle() {
        if (h_pinned) cudaFreeHost(h_pinned);
        for (cudaEvent_t e : pev) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <class T>
int dev_alloc(DevBuf& b, size_t count, std::string& err) {
    b.owned = true;
    CUDA_TRY(cudaMalloc(&b.p, std::max<size_t>(count, 1) * sizeof(T)));
    return SSDE_OK;
}
// Uninitialised host storage for the big packing buffers: std::vector would zero-fill them on one
// thread (page faults included) before the worker threads overwrite every entry that is ever read.
template <class T>
struct RawVec {
    struct Free { void operator()(T* q) const { std::free(q); } };
    std::unique_ptr<T, Free> p;
    size_t n = 0;
    // 2 MB-aligned and advised as transparent huge pages: the first touch of a multi-GB buffer by 16
    // threads is otherwise dominated by 4 KB page faults
    void alloc(size_t count) {
        static_assert(std::is_trivial<T>::value, "RawVec holds plain data");
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        void* q = nullptr;
        if (bytes >= (size_t(8) << 20)) {
            const size_t al = size_t(2) << 20;
            if (posix_memalign(&q, al, (bytes + al - 1) / al * al) != 0) q = nullptr;
            if (q) madvise(q, (bytes + al - 1) / al * al, MADV_HUGEPAGE);
        } else {
            q = std::malloc(bytes);
        }
        if (!q) throw std::bad_alloc();
        p.reset(static_cast<T*>(q));
        n = count;
    }
    T* data() { return p.get(); }
    const T* data() const { return p.get(); }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return p.get()[i]; }
    const T& operator[](size_t i) const { return p.get()[i]; }
};
template <class T>
int dev_upload(DevBuf& b, const RawVec<T>& v, std::string& err) {
    int rc = dev_alloc<T>(b, v.size(), err);
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SSDE_OK;
}
template <class T>
int dev_upload(DevBuf& b, const std::vector<T>& v, std::string& err) {
    int rc = dev_alloc<T>(b, v.size(), err);
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SSDE_OK;
}

// ---------------------------------------------------------------------------------------------
// small kernels: theta gather, finalisation (reductions over tiles, penalty, packing)
// ---------------------------------------------------------------------------------------------
// `dir` / `theta_dot` (both or neither): direction of the tangent pass, gathered like theta.
__global__ void gather_theta_kernel(const double* __restrict__ par, double* __restrict__ theta,
                                    const double* __restrict__ dir, double* __restrict__ theta_dot,
                                    int p_fe, int p_re, int o_fe, int o_re,
                                    const int32_t* __restrict__ mu_cols, int n_mu_cols, int* __restrict__ mu_zero) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p_fe) { theta[i] = par[o_fe + i]; if (dir) theta_dot[i] = dir[o_fe + i]; }
    else if (i < p_fe + p_re) { theta[i] = par[o_re + (i - p_fe)]; if (dir) theta_dot[i] = dir[o_re + (i - p_fe)]; }
    if (blockIdx.x == 0 && mu_zero) {
        // mu_d == 0 on every row iff every theta entry its design columns touch is exactly 0
        // (and, in a tangent pass, so is the direction)
        int nz = (n_mu_cols < 0) ? 1 : 0;
        for (int k = threadIdx.x; k < n_mu_cols; k += blockDim.x) {
            const int c = mu_cols[k];
            const int o = (c < p_fe) ? o_fe + c : o_re + (c - p_fe);
            if (par[o] != 0.0 || (dir && dir[o] != 0.0)) nz = 1;
        }
        nz = __syncthreads_or(nz);
        if (threadIdx.x == 0) *mu_zero = nz ? 0 : 1;
    }
}

struct FinArgs {
    const double* par;
    const double* par_dot;    // tangent pass: direction in the parameter vector (else nullptr)
    const double* grad_theta; // [p] (tangent pass: [2 p], tangents second)
    const double* part_llk;   // per-tile / per-block log-likelihood partial sums
    int n_part;
    const double* tile_gh;    // CTCRW: per-tile d nllk / d h partial sums (or nullptr)
    const double* tile_gh_dot;
    int n_gh;
    const uint32_t* S_rowptr;
    const uint32_t* S_col;
    const double* S_val;
    const int32_t* sm_off;    // [n_s + 1] offsets of the smooth blocks in coeff_re
    double* sb;               // [2 p_re] scratch: S b (as R)
    int p_fe, p_re, n_s, npar, o_sig, o_fe, o_ll, o_re;
    int o_dec, n_dec;         // decay models: gradient entries of log_decay from grad_decay ([2 n_dec] in a tangent pass)
    const double* grad_decay;
    int penalty;              // 0 none, 1 Kalman form (nllk_ctcrw.hpp:254-280), 2 nllk_sde form
    double pen_const;         // sum_i [Sn_i/2 log(2 pi) - 1/2 log det S_i]  (nllk_sde.hpp:114-116)
    int want_grad;
    const unsigned* error;
    double* out;              // [1 + npar + 1]: nllk, gradient, status
    double* hv;               // tangent pass: [npar] Hessian-vector product
};

// deterministic sum of a strided array by one block
__device__ double block_reduce_array(const double* x, int n, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    // tree over the block (blockDim = 256)
    __syncthreads();
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    return red[0];
}

// Fixed-shape partial sums of the per-tile outputs (deterministic: block b always sums the same
// slice in the same order).  part[b] = sum of x[slice b], part[RED_BLOCKS + b] likewise for y,
// part[2 RED_BLOCKS + b] for z.
__global__ void __launch_bounds__(256) reduce_tiles_kernel(const double* __restrict__ x, int nx,
                                                           const double* __restrict__ y,
                                                           const double* __restrict__ z, int ny,
                                                           double* __restrict__ part) {
    __shared__ double red[256];
    const int b = blockIdx.x, nb = gridDim.x;
    {
        const int per = (nx + nb - 1) / nb;
        const int lo = min(b * per, nx), hi = min(lo + per, nx);
        const double s = block_reduce_array(x + lo, hi - lo, red);
        if (threadIdx.x == 0) part[b] = s;
    }
    const double* more[2] = {y, z};
    for (int m = 0; m < 2; ++m) {
        if (!more[m]) continue;
        __syncthreads();
        const int per = (ny + nb - 1) / nb;
        const int lo = min(b * per, ny), hi = min(lo + per, ny);
        const double s = block_reduce_array(more[m] + lo, hi - lo, red);
        if (threadIdx.x == 0) part[(m + 1) * nb + b] = s;
    }
}

// Reductions over tiles, chain rule for log_sigma_obs, smoothing penalty (nllk_ctcrw.hpp:254-280 /
// nllk_sde.hpp:89-124) and packing.  R = Dual: the same arithmetic on (value, tangent) pairs
// also yields the Hessian-vector product of the penalised objective in a.hv.
template <class R>
__global__ void __launch_bounds__(256) finalize_kernel(FinArgs a) {
    constexpr bool TAN = !std::is_same<R, double>::value;
    __shared__ double red[256];
    __shared__ R s_quad[64];
    const int tid = threadIdx.x;
    const int p = a.p_fe + a.p_re;
    auto P = [&](int i) -> R { return ScalarOf<R>::make(a.par[i], TAN ? a.par_dot[i] : 0.0); };
    auto put = [&](int i, const R& g) {            // gradient entry i (and its tangent)
        a.out[1 + i] = value(g);
        if (TAN) a.hv[i] = tangent(g);
    };
    R nllk = -block_reduce_array(a.part_llk, a.n_part, red);
    __syncthreads();
    if (a.want_grad) {
        for (int i = tid; i < a.npar; i += blockDim.x) put(i, R(0.0));
        __syncthreads();
        if (a.tile_gh) {
            const double ghv = block_reduce_array(a.tile_gh, a.n_gh, red);
            __syncthreads();
            double ghd = 0.0;
            if (TAN) { ghd = block_reduce_array(a.tile_gh_dot, a.n_gh, red); __syncthreads(); }
            if (tid == 0) { const R h = exp(2.0 * P(a.o_sig)); put(a.o_sig, 2.0 * h * ScalarOf<R>::make(ghv, ghd)); }
        }
        for (int i = tid; i < p; i += blockDim.x) {
            const R g = ScalarOf<R>::make(a.grad_theta[i], TAN ? a.grad_theta[p + i] : 0.0);
            put(i < a.p_fe ? a.o_fe + i : a.o_re + (i - a.p_fe), g);
        }
        for (int i = tid; i < a.n_dec; i += blockDim.x)
            put(a.o_dec + i, ScalarOf<R>::make(a.grad_decay[i], TAN ? a.grad_decay[a.n_dec + i] : 0.0));
    }
    __syncthreads();
    if (a.penalty) {
        R* sb = reinterpret_cast<R*>(a.sb);
        for (int r = tid; r < a.p_re; r += blockDim.x) {
            R acc = 0.0;
            for (uint32_t k = a.S_rowptr[r]; k < a.S_rowptr[r + 1]; ++k) acc = fmad(a.S_val[k], P(a.o_re + (int)a.S_col[k]), acc);
            sb[r] = acc;
        }
        __syncthreads();
        for (int i0 = 0; i0 < a.n_s; i0 += 64) {
            // quadratic forms of up to 64 smooths at a time: one block-strided pass each
            for (int i = i0; i < a.n_s && i < i0 + 64; ++i) {
                R q = 0.0;
                for (int r = a.sm_off[i] + tid; r < a.sm_off[i + 1]; r += blockDim.x) q += P(a.o_re + r) * sb[r];
                double part[2] = {value(q), tangent(q)};
                for (int c = 0; c < (TAN ? 2 : 1); ++c) {
                    __syncthreads();
                    red[tid] = part[c];
                    __syncthreads();
                    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
                        if (tid < o) red[tid] += red[tid + o];
                        __syncthreads();
                    }
                    part[c] = red[0];
                }
                if (tid == 0) s_quad[i - i0] = ScalarOf<R>::make(part[0], part[1]);
                __syncthreads();
            }
            if (tid == 0) {
                for (int i = i0; i < a.n_s && i < i0 + 64; ++i) {
                    const double Sn = (double)(a.sm_off[i + 1] - a.sm_off[i]);
                    const R ll = P(a.o_ll + i), lam = exp(ll);
                    nllk += -0.5 * Sn * ll + 0.5 * lam * s_quad[i - i0];
                    if (a.want_grad) put(a.o_ll + i, -0.5 * Sn + 0.5 * lam * s_quad[i - i0]);
                }
            }
            __syncthreads();
            if (a.want_grad) {
                for (int i = i0; i < a.n_s && i < i0 + 64; ++i) {
                    const R lam = exp(P(a.o_ll + i));
                    for (int r = a.sm_off[i] + tid; r < a.sm_off[i + 1]; r += blockDim.x) {
                        const R g = lam * sb[r];
                        a.out[1 + a.o_re + r] += value(g);
                        if (TAN) a.hv[a.o_re + r] += tangent(g);
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0 && a.penalty == 2) nllk += a.pen_const;
    }
    if (tid == 0) {
        a.out[0] = value(nllk);
        a.out[1 + a.npar] = (double)(*a.error);
    }
}

// hess[(p_fe + r), (p_fe + c)] += exp(log_lambda_i) S[r, c] for the rows r of smooth i (one thread per row of S)
__global__ void hess_penalty_kernel(const double* __restrict__ par, int o_ll, int n_s, const int32_t* __restrict__ sm_off,
                                    const uint32_t* __restrict__ S_rowptr, const uint32_t* __restrict__ S_col,
                                    const double* __restrict__ S_val, int p_fe, int p_re, double* __restrict__ hess) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p_re) return;
    int i = 0;
    while (i + 1 < n_s && r >= sm_off[i + 1]) ++i;
    const double lam = exp(par[o_ll + i]);
    const size_t p = (size_t)p_fe + p_re;
    for (uint32_t k = S_rowptr[r]; k < S_rowptr[r + 1]; ++k)
        atomicAdd(hess + (size_t)(p_fe + (int)S_col[k]) * p + (p_fe + r), lam * S_val[k]);
}

// dir = e_j
__global__ void unit_vector_kernel(double* __restrict__ dir, int n, int j) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dir[i] = (i == j) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// host-side packing of the triplet design (ssde_create): O(nnz) bucket passes + small per-row
// sorts, spread over the host cores (a comparison sort of 1.5e8 triplets alone took ~10 s)
// ---------------------------------------------------------------------------------------------
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(std::getenv("SSDE_PACK_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ssde pack] %-28s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};
int host_threads(int64_t work) {
    unsigned hc = std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = (unsigned)CPU_COUNT(&set);
    int t = (int)std::min<unsigned>(std::max(hc, 1u), 32u);
    if (work < (1 << 16)) t = 1;
    return t;
}
// fn(thread, lo, hi) over [0, n) in contiguous pieces; exceptions of the workers are rethrown here
template <class Fn>
void parallel_ranges(int64_t n, int nthreads, Fn&& fn) {
    if (nthreads <= 1 || n < 2) { fn(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> ex((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        th.emplace_back([&, t, lo, hi] {
            try { fn(t, lo, hi); } catch (...) { ex[(size_t)t] = std::current_exception(); }
        });
    }
    for (auto& x : th) x.join();
    for (auto& e : ex) if (e) std::rethrow_exception(e);
}

// Row-major packed design: the nonzeros of row i (all parameters, parameter-major, columns
// ascending, duplicates summed in triplet order) start at rowptr[i]; byte p of cnt[i] = nonzeros
// of parameter p.  Rows are NOT contiguous in col / val (the space of summed duplicates and of
// explicit zeros stays unused), `nnz` is the number of packed nonzeros.
struct Packed {
    std::vector<uint32_t> rowptr, cnt;
    RawVec<uint32_t> col;
    RawVec<double> val;
    int64_t nnz = 0;
};

int pack_design(const ssde_desc& d, int n_par, Packed& out, std::string& err) {
    const int64_t n = d.n;
    const ssde_triplet* mats[2] = {&d.X_fe, &d.X_re};
    const int64_t nnz_fe = d.X_fe.nnz, total = d.X_fe.nnz + d.X_re.nnz;
    const int nt = host_threads(total);
    auto triplet = [&](int64_t k, int64_t& r, int64_t& c, double& x, uint32_t& coff) {
        const ssde_triplet& T = *mats[k < nnz_fe ? 0 : 1];
        const int64_t kk = k < nnz_fe ? k : k - nnz_fe;
        r = T.i[kk]; c = T.j[kk]; x = T.x[kk];
        coff = k < nnz_fe ? 0u : (uint32_t)d.X_fe.ncol;
        return r >= 0 && r < T.nrow && c >= 0 && c < T.ncol;
    };
    PhaseTimer pt;
    // (1) nonzeros per row
    std::vector<uint32_t> fill((size_t)n, 0u);
    std::vector<int> bad((size_t)nt, 0);
    parallel_ranges(total, nt, [&](int t, int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; ++k) {
            int64_t r, c; double x; uint32_t coff;
            if (!triplet(k, r, c, x, coff)) { bad[(size_t)t] = 1; return; }
            if (x == 0.0) continue;
            __atomic_fetch_add(&fill[(size_t)(r % n)], 1u, __ATOMIC_RELAXED);
        }
    });
    for (int b : bad) if (b) { err = "design triplet index out of range"; return SSDE_ERR_BAD_ARG; }
    pt.lap("pack: count");
    // (2) row starts
    out.rowptr.assign((size_t)n + 1, 0u);
    uint64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) { out.rowptr[(size_t)i] = (uint32_t)acc; acc += fill[(size_t)i]; fill[(size_t)i] = 0u; }
    if (acc > 0xfffffff0ull) { err = "more than 2^32 nonzeros in one shard"; return SSDE_ERR_UNSUPPORTED; }
    out.rowptr[(size_t)n] = (uint32_t)acc;
    // (3) scatter into the rows' buckets; the triplet index travels along so that the per-row sort
    //     (and with it the order in which duplicates are summed) does not depend on the thread count
    struct Ent { uint32_t col, k; double x; };
    RawVec<Ent> ents;
    RawVec<uint8_t> par;
    ents.alloc((size_t)acc);
    par.alloc((size_t)acc);
    if (total > 0xfffffff0ll) { err = "more than 2^32 triplets in one shard"; return SSDE_ERR_UNSUPPORTED; }
    parallel_ranges(total, nt, [&](int, int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; ++k) {
            int64_t r, c; double x; uint32_t coff;
            triplet(k, r, c, x, coff);
            if (x == 0.0) continue;
            const int64_t i = r % n;
            const uint32_t slot = out.rowptr[(size_t)i] + __atomic_fetch_add(&fill[(size_t)i], 1u, __ATOMIC_RELAXED);
            ents[slot] = Ent{coff + (uint32_t)c, (uint32_t)k, x};
            par[slot] = (uint8_t)(r / n);
        }
    });
    pt.lap("pack: alloc + scatter");
    // (4) per row: sort by (parameter, column, triplet index), sum duplicates, compact in place
    out.cnt.assign((size_t)n, 0u);
    out.col.alloc((size_t)acc);
    out.val.alloc((size_t)acc);
    std::vector<int64_t> nnz_t((size_t)nt, 0);
    std::vector<int> over((size_t)nt, 0);
    parallel_ranges(n, nt, [&](int t, int64_t lo, int64_t hi) {
        std::vector<uint32_t> idx;
        for (int64_t i = lo; i < hi; ++i) {
            const uint32_t b = out.rowptr[(size_t)i], m = fill[(size_t)i];
            idx.resize(m);
            for (uint32_t u = 0; u < m; ++u) idx[u] = b + u;
            std::sort(idx.begin(), idx.end(), [&](uint32_t a_, uint32_t b_) {
                if (par[a_] != par[b_]) return par[a_] < par[b_];
                if (ents[a_].col != ents[b_].col) return ents[a_].col < ents[b_].col;
                return ents[a_].k < ents[b_].k;
            });
            uint32_t w = 0, o = b;
            uint32_t pc[MAX_NP] = {0, 0, 0, 0};
            for (uint32_t u = 0; u < m;) {
                const uint32_t e0 = idx[u];
                double x = ents[e0].x;
                uint32_t v = u + 1;
                while (v < m && par[idx[v]] == par[e0] && ents[idx[v]].col == ents[e0].col) { x += ents[idx[v]].x; ++v; }
                out.col[o] = ents[e0].col; out.val[o] = x; ++o;
                if (++pc[par[e0]] > 255u) over[(size_t)t] = 1;
                u = v;
            }
            for (int p = 0; p < n_par; ++p) w |= (pc[p] & 255u) << (8 * p);
            out.cnt[(size_t)i] = w;
            nnz_t[(size_t)t] += o - b;
        }
    });
    for (int b : over) if (b) { err = "more than 255 nonzeros for one (row, parameter)"; return SSDE_ERR_UNSUPPORTED; }
    out.nnz = 0;
    for (int64_t v : nnz_t) out.nnz += v;
    pt.lap("pack: row sort + dedupe");
    return SSDE_OK;
}

// Warp-tile transposed layout (design.cuh) from the row-major packed form.
struct V2Host {
    std::vector<WtDesc> desc;
    RawVec<double> val;
    RawVec<uint32_t> col;
};

// allow_alias: parameters whose value slots are identical in every row of a uniform warp-tile
// (tau ~ s(time), nu ~ s(time)) store them once (design.cuh, alias_of); every kernel but
// sde_decay_kernel reads the alias flags, so decay models keep the plain layout.
int build_v2(const Packed& pk, int64_t n, int64_t n_pad, int n_par, bool allow_alias, V2Host& out, std::string& err) {
    const int64_t nwt = n_pad / WT;
    out.desc.assign((size_t)nwt, WtDesc{0, 0, 0u, WT_UNIFORM});
    const int nt = host_threads(n * 8);
    PhaseTimer pt;
    if (std::getenv("SSDE_NO_ALIAS")) allow_alias = false;      // A/B runs of the two layouts
    // (1) per warp-tile: union of columns / max count per parameter -> slots, uniform or not
    std::vector<std::vector<uint32_t>> wcols((size_t)nwt);      // uniform warp-tiles: their column list
    std::vector<uint64_t> wS((size_t)nwt, 0);                   // column slots
    std::vector<uint64_t> wSV((size_t)nwt, 0);                  // value slots (< column slots with aliases)
    parallel_ranges(nwt, nt, [&](int, int64_t qlo, int64_t qhi) {
        std::vector<std::vector<uint32_t>> U((size_t)n_par);
        for (int64_t q = qlo; q < qhi; ++q) {
            const int64_t r0 = q * WT, r1 = std::min<int64_t>(n, r0 + WT);
            WtDesc& d = out.desc[(size_t)q];
            if (r0 >= r1) continue;              // padding warp-tile: no slots
            uint32_t kmax_row[MAX_NP] = {0, 0, 0, 0};
            for (int p = 0; p < n_par; ++p) U[(size_t)p].clear();
            for (int64_t r = r0; r < r1; ++r) {
                uint32_t rp = pk.rowptr[(size_t)r];
                for (int p = 0; p < n_par; ++p) {
                    const uint32_t k = (pk.cnt[(size_t)r] >> (8 * p)) & 255u;
                    kmax_row[p] = std::max(kmax_row[p], k);
                    for (uint32_t j = 0; j < k; ++j) U[(size_t)p].push_back(pk.col[rp + j]);
                    rp += k;
                }
            }
            size_t s_union = 0, s_max = 0;
            bool fits = true;
            for (int p = 0; p < n_par; ++p) {
                auto& u = U[(size_t)p];
                std::sort(u.begin(), u.end());
                u.erase(std::unique(u.begin(), u.end()), u.end());
                s_union += u.size(); s_max += kmax_row[p];
                if (u.size() > 255) fits = false;
            }
            const bool uniform = fits && s_union <= std::max(s_max + 4, 2 * s_max);
            uint32_t kw = 0;
            size_t S = 0;
            for (int p = 0; p < n_par; ++p) {
                const uint32_t k = uniform ? (uint32_t)U[(size_t)p].size() : kmax_row[p];
                kw |= k << (8 * p);
                S += k;
            }
            d.kmax = kw;
            d.flags = uniform ? WT_UNIFORM : 0u;
            wS[(size_t)q] = S;
            if (uniform && allow_alias) {
                // parameter p is an alias of t < p when, in every row, both have the same number of
                // nonzeros, with bit-identical values, at the same slot positions of their column lists
                for (int p = 1; p < n_par; ++p) {
                    const size_t kp = U[(size_t)p].size();
                    if (kp == 0) continue;
                    for (int t = 0; t < p; ++t) {
                        if (alias_of(d.flags, t) >= 0 || U[(size_t)t].size() != kp) continue;
                        bool same = true;
                        for (int64_t r = r0; r < r1 && same; ++r) {
                            const uint32_t cw = pk.cnt[(size_t)r];
                            const uint32_t cp = (cw >> (8 * p)) & 255u, ct = (cw >> (8 * t)) & 255u;
                            if (cp != ct) { same = false; break; }
                            uint32_t op = pk.rowptr[(size_t)r], ot = op;
                            for (int pp = 0; pp < p; ++pp) op += (cw >> (8 * pp)) & 255u;
                            for (int pp = 0; pp < t; ++pp) ot += (cw >> (8 * pp)) & 255u;
                            for (uint32_t j = 0; j < cp; ++j) {
                                const size_t sp = (size_t)(std::lower_bound(U[(size_t)p].begin(), U[(size_t)p].end(), pk.col[op + j]) - U[(size_t)p].begin());
                                const size_t st = (size_t)(std::lower_bound(U[(size_t)t].begin(), U[(size_t)t].end(), pk.col[ot + j]) - U[(size_t)t].begin());
                                if (sp != st || std::memcmp(&pk.val[op + j], &pk.val[ot + j], sizeof(double)) != 0) { same = false; break; }
                            }
                        }
                        if (same) { d.flags |= (uint32_t)(t + 1) << (WT_ALIAS_SHIFT + 2 * p); break; }
                    }
                }
            }
            wSV[(size_t)q] = (uint64_t)value_slots_of(d.kmax, d.flags);
            if (uniform) {
                auto& c = wcols[(size_t)q];
                c.reserve(S);
                for (int p = 0; p < n_par; ++p) c.insert(c.end(), U[(size_t)p].begin(), U[(size_t)p].end());
            }
        }
    });
    pt.lap("v2: warp-tile analysis");
    // (2) offsets; identical consecutive column lists of uniform warp-tiles are stored once
    uint64_t nval = 0, ncol = 0;
    std::vector<int64_t> owner((size_t)nwt, -1); // uniform warp-tiles: the warp-tile whose column list this one uses
    int64_t prev = -1;                           // previous uniform warp-tile that owns a column list
    for (int64_t q = 0; q < nwt; ++q) {
        WtDesc& d = out.desc[(size_t)q];
        d.val_off = (int64_t)nval;
        d.col_off = 0;
        const uint64_t S = wS[(size_t)q];
        if (S == 0 && !(q * WT < n)) continue;
        nval += wSV[(size_t)q] * WT;
        if (d.flags & WT_UNIFORM) {
            if (prev >= 0 && wcols[(size_t)q] == wcols[(size_t)prev]) {
                d.col_off = out.desc[(size_t)prev].col_off;
                std::vector<uint32_t>().swap(wcols[(size_t)q]);
                owner[(size_t)q] = prev;
            } else {
                d.col_off = (int64_t)ncol;
                ncol += S;
                prev = q;
                owner[(size_t)q] = q;
            }
        } else {
            d.col_off = (int64_t)ncol;
            ncol += S * WT;
        }
    }
    out.val.alloc((size_t)nval);
    out.col.alloc((size_t)std::max<uint64_t>(ncol, 1));
    out.col[0] = 0u;
    pt.lap("v2: offsets + alloc");
    // (3) fill values (and columns) of every warp-tile
    parallel_ranges(nwt, nt, [&](int, int64_t qlo, int64_t qhi) {
        for (int64_t q = qlo; q < qhi; ++q) {
            const int64_t r0 = q * WT, r1 = std::min<int64_t>(n, r0 + WT);
            if (r0 >= r1) continue;
            const WtDesc& d = out.desc[(size_t)q];
            const bool uniform = (d.flags & WT_UNIFORM) != 0;
            const size_t S = (size_t)wS[(size_t)q];
            const size_t SV = (size_t)wSV[(size_t)q];               // value slots per row (aliased parameters store none)
            uint64_t vofs = 0;
            value_slots_of(d.kmax, d.flags, &vofs);
            double* v = out.val.data() + d.val_off;
            std::memset(v, 0, sizeof(double) * SV * WT);             // explicit zeros for rows that lack a column of the union
            uint32_t* c = nullptr;
            const uint32_t* ulist = nullptr;     // this warp-tile's column list (parameter-major, ascending per parameter)
            if (uniform) {
                // a shared list is written to out.col by its owner -- possibly another thread -- so
                // slots are looked up in the owner's private copy, not in out.col
                const int64_t o = owner[(size_t)q];
                if (o == q) std::copy(wcols[(size_t)q].begin(), wcols[(size_t)q].end(), out.col.data() + d.col_off);
                ulist = wcols[(size_t)o].data();
            } else {
                c = out.col.data() + d.col_off;
                std::memset(c, 0, sizeof(uint32_t) * S * WT);
            }
            for (int64_t r = r0; r < r1; ++r) {
                const int rr = (int)(r - r0), k = rr % LC, lane = rr / LC;
                uint32_t rp = pk.rowptr[(size_t)r];
                size_t slot0 = 0;
                for (int p = 0; p < n_par; ++p) {
                    const uint32_t cntp = (pk.cnt[(size_t)r] >> (8 * p)) & 255u;
                    const size_t kp = (d.kmax >> (8 * p)) & 255u;
                    const size_t vslot0 = (size_t)((vofs >> (16 * p)) & 0xffffull);
                    for (uint32_t j = 0; j < cntp && alias_of(d.flags, p) < 0; ++j) {
                        size_t slot;
                        if (uniform) slot = (size_t)(std::lower_bound(ulist + slot0, ulist + slot0 + kp, pk.col[rp + j]) - (ulist + slot0));
                        else slot = j;
                        v[((size_t)k * SV + vslot0 + slot) * 32 + lane] = pk.val[rp + j];
                        if (c) c[((size_t)k * S + slot0 + slot) * 32 + lane] = pk.col[rp + j];
                    }
                    rp += cntp;
                    slot0 += kp;
                }
            }
        }
    });
    pt.lap("v2: fill");
    (void)err;
    return SSDE_OK;
}

// log det of a dense SPD block by Cholesky; NaN if not positive definite (the reference's
// atomic::matinvpd gives garbage there too, nllk_sde.hpp:109-111)
double logdet_spd(std::vector<double>& A, int m) {
    double ld = 0.0;
    for (int j = 0; j < m; ++j) {
        double s = A[(size_t)j * m + j];
        for (int k = 0; k < j; ++k) s -= A[(size_t)j * m + k] * A[(size_t)j * m + k];
        if (!(s > 0.0)) return NAN;
        const double l = std::sqrt(s);
        A[(size_t)j * m + j] = l;
        ld += 2.0 * std::log(l);
        for (int i = j + 1; i < m; ++i) {
            double t = A[(size_t)i * m + j];
            for (int k = 0; k < j; ++k) t -= A[(size_t)i * m + k] * A[(size_t)j * m + k];
            A[(size_t)i * m + j] = t / l;
        }
    }
    return ld;
}

int setup_penalty(ssde_handle* h, const ssde_triplet& S, int n_smooth, const int32_t* ncol_re,
                  std::string& err) {
    h->n_s = n_smooth;
    h->has_smooth = n_smooth > 0 && ncol_re[0] > 0;       // `if(ncol_re(0) > 0)`, nllk_ctcrw.hpp:256
    std::vector<int32_t> off(n_smooth + 1, 0);
    if (h->has_smooth) {
        for (int i = 0; i < n_smooth; ++i) off[i + 1] = off[i] + ncol_re[i];
        if (off[n_smooth] != h->p_re) { err = "sum(ncol_re) != ncol(X_re)"; return SSDE_ERR_BAD_ARG; }
        if (S.nrow != h->p_re || S.ncol != h->p_re) { err = "S must be p_re x p_re"; return SSDE_ERR_BAD_ARG; }
    }
    // CSR of S with duplicates summed
    std::vector<uint32_t> rp(h->p_re + 1, 0), cols;
    std::vector<double> vals;
    if (h->has_smooth) {
        struct E { int32_t r, c; double x; };
        std::vector<E> es;
        for (int64_t k = 0; k < S.nnz; ++k) {
            if (S.i[k] < 0 || S.i[k] >= S.nrow || S.j[k] < 0 || S.j[k] >= S.ncol) { err = "S triplet index out of range"; return SSDE_ERR_BAD_ARG; }
            es.push_back({S.i[k], S.j[k], S.x[k]});
        }
        std::sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
        size_t k = 0;
        for (int r = 0; r < h->p_re; ++r) {
            rp[r] = (uint32_t)cols.size();
            while (k < es.size() && es[k].r == r) {
                int c = es[k].c; double x = es[k].x; ++k;
                while (k < es.size() && es[k].r == r && es[k].c == c) { x += es[k].x; ++k; }
                cols.push_back((uint32_t)c); vals.push_back(x);
            }
        }
        rp[h->p_re] = (uint32_t)cols.size();
        // additive constants of nllk_sde's penalty (nllk_sde.hpp:109-116)
        if (!is_kalman(h->model)) {
            double cst = 0.0;
            for (int i = 0; i < n_smooth; ++i) {
                const int m = ncol_re[i], o = off[i];
                bool diag = true;
                double ld = 0.0;
                for (int r = o; r < o + m && diag; ++r)
                    for (uint32_t q = rp[r]; q < rp[r + 1]; ++q)
                        if ((int)cols[q] != r && vals[q] != 0.0) { diag = false; break; }
                if (diag) {
                    for (int r = o; r < o + m; ++r) {
                        double dd = 0.0;
                        for (uint32_t q = rp[r]; q < rp[r + 1]; ++q) if ((int)cols[q] == r) dd += vals[q];
                        ld += std::log(dd);
                    }
                } else {
                    std::vector<double> A((size_t)m * m, 0.0);
                    for (int r = o; r < o + m; ++r)
                        for (uint32_t q = rp[r]; q < rp[r + 1]; ++q) {
                            const int c = (int)cols[q] - o;
                            if (c >= 0 && c < m) A[(size_t)(r - o) * m + c] = vals[q];
                        }
                    ld = logdet_spd(A, m);
                }
                cst += 0.5 * m * std::log(2.0 * M_PI) - 0.5 * ld;
            }
            h->pen_const = cst;
        }
    }
    // hot columns of the transposed design product: the fixed effects and every small smooth block are
    // touched by every warp-tile, large blocks (one random intercept per track) by a few
    h->hot.n = 0;
    {
        const int p_theta = h->p_fe + h->p_re;
        int used = 0;
        auto add = [&](int lo, int hi) {
            if (hi <= lo || h->hot.n >= SDE_MAX_HOT_RANGES || used + (hi - lo) > SDE_HOT) return;
            h->hot.lo[h->hot.n] = lo; h->hot.hi[h->hot.n] = hi; h->hot.off[h->hot.n] = used;
            used += hi - lo; ++h->hot.n;
        };
        if (p_theta <= SDE_HOT) add(0, p_theta);
        else {
            add(0, std::min(h->p_fe, 128));
            if (h->has_smooth)
                for (int i = 0; i < n_smooth; ++i)
                    if (ncol_re[i] <= 64) add(h->p_fe + off[i], h->p_fe + off[i + 1]);
        }
    }
    h->hess_hot.n = 0;
    for (int r = 0; r < h->hot.n; ++r)
        for (int c = h->hot.lo[r]; c < h->hot.hi[r] && h->hess_hot.n < SDE_HESS_HOT; ++c) h->hess_hot.cols[h->hess_hot.n++] = c;
    int rc;
    if ((rc = dev_upload(h->S_rowptr, rp, err))) return rc;
    if ((rc = dev_upload(h->S_col, cols, err))) return rc;
    if ((rc = dev_upload(h->S_val, vals, err))) return rc;
    if ((rc = dev_upload(h->sm_off, off, err))) return rc;
    if ((rc = dev_alloc<double>(h->sb, 2 * (size_t)h->p_re, err))) return rc;
    return SSDE_OK;
}

template <class K>
int max_grid(K kernel, int nt, size_t smem, int num_sms, std::string& err, int& grid) {
    int occ = 0;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nt, smem));
    if (occ < 1) { err = "kernel does not fit on an SM"; return SSDE_ERR_CUDA; }
    grid = occ * num_sms;
    return SSDE_OK;
}

// Calls fn(M{}) with the traits class (models.cuh) of the handle's Kalman model and dimension.
template <class R, class Fn>
int with_kalman_model(const ssde_handle* h, Fn&& fn) {
    const int nd = h->n_dim;
#ifdef SSDE_MINIMAL
    // kernel-tuning build (scripts/tune_build.sh): only the benchmark's instantiation, seconds to compile
    if constexpr (std::is_same<R, double>::value) { if (h->model == SSDE_CTCRW && nd == 2 && !h->dense) return fn(CtcrwModel<2, R>{}); }
    return SSDE_ERR_UNSUPPORTED;
#else
    switch (h->model) {
        case SSDE_CTCRW:
            if (h->dense) return nd == 1 ? fn(DenseModel<CtcrwModel<1, R>>{}) : fn(DenseModel<CtcrwModel<2, R>>{});
            return nd == 1 ? fn(CtcrwModel<1, R>{}) : fn(CtcrwModel<2, R>{});
        case SSDE_OU_SSM:
            if (h->dense) return nd == 1 ? fn(DenseModel<OuSsmModel<1, R>>{}) : fn(DenseModel<OuSsmModel<2, R>>{});
            return nd == 1 ? fn(OuSsmModel<1, R>{}) : fn(OuSsmModel<2, R>{});
        case SSDE_BM_SSM:
            if (h->dense) return nd == 1 ? fn(DenseModel<BmSsmModel<1, R>>{}) : (nd == 2 ? fn(DenseModel<BmSsmModel<2, R>>{}) : fn(DenseModel<BmSsmModel<3, R>>{}));
            return nd == 1 ? fn(BmSsmModel<1, R>{}) : (nd == 2 ? fn(BmSsmModel<2, R>{}) : fn(BmSsmModel<3, R>{}));
    }
    return SSDE_ERR_UNSUPPORTED;
#endif
}

template <class M>
int ctcrw_grids(ssde_handle* h) {
    std::string& err = h->err;
    int rc;
    if ((rc = max_grid(ctcrw_fwd_kernel<M, M::KNT, M::FMINB>, M::KNT, sizeof(FwdSmem<M, M::KNT>), h->num_sms, err, h->grid_f))) return rc;
    if ((rc = max_grid(ctcrw_bwd_kernel<M, M::KNT, M::MINB>, M::KNT, sizeof(BwdSmem<M, M::KNT>), h->num_sms, err, h->grid_b))) return rc;
    return SSDE_OK;
}

template <int MODEL, int ND>
int sde_grid(ssde_handle* h) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    int rc = max_grid(sde_fused_kernel<MODEL, ND>, SDE_NT, sizeof(SdeSmem<NP>), h->num_sms, h->err, h->grid_lp);
    if (rc) return rc;
    if ((rc = max_grid(sde_stream_kernel<MODEL, ND>, SDE_NT, sizeof(SdeStreamSmem), h->num_sms, h->err, h->grid_stream))) return rc;
    return max_grid(sde_hess_kernel<MODEL, ND>, SDE_NT, sizeof(SdeHessSmem<NP>), h->num_sms, h->err, h->grid_hess);
}

// allocate the per-evaluation work buffers once the data are on the device
int finish_setup(ssde_handle* h) {
    std::string& err = h->err;
    int rc;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    const int p = h->p_fe + h->p_re;
    h->o_sig = is_kalman(h->model) ? 0 : -1;
    h->o_fe = is_kalman(h->model) ? 1 : 0;
    h->o_ll = h->o_fe + h->p_fe;
    h->o_dec = h->n_dec > 0 ? h->o_ll + h->n_s : -1;
    h->o_re = h->o_ll + h->n_s + h->n_dec;
    h->npar = h->o_re + h->p_re;
    if ((rc = dev_alloc<double>(h->par, h->npar, err))) return rc;
    if ((rc = dev_alloc<double>(h->theta, p, err))) return rc;
    if ((rc = dev_alloc<double>(h->grad_theta, p, err))) return rc;
    if ((rc = dev_alloc<double>(h->grad_decay, 2 * (size_t)std::max(h->n_dec, 1), err))) return rc;
    if ((rc = dev_alloc<double>(h->out, h->npar + 2, err))) return rc;
    if ((rc = dev_alloc<double>(h->part, 3 * RED_BLOCKS, err))) return rc;
    if ((rc = dev_alloc<int>(h->mu_zero, 1, err))) return rc;
    CUDA_TRY(cudaMemset(h->mu_zero.p, 0, sizeof(int)));
    if (!h->mu_cols.p && (rc = dev_alloc<int32_t>(h->mu_cols, 1, err))) return rc;
    if ((rc = dev_alloc<unsigned long long>(h->stats, 32, err))) return rc;
    CUDA_TRY(cudaMemset(h->stats.p, 0, 32 * sizeof(unsigned long long)));
    if ((rc = dev_alloc<unsigned>(h->counters, 4, err))) return rc;
    CUDA_TRY(cudaMemset(h->counters.p, 0, 4 * sizeof(unsigned)));
    CUDA_TRY(cudaMallocHost(&h->h_pinned, sizeof(double) * (2 * (size_t)h->npar + 4)));
    if (is_kalman(h->model)) {
        h->nchunks = h->n_pad / LC;
        rc = with_kalman_model<double>(h, [&](auto m) -> int {
            using M = decltype(m);
            int rc;
            h->ntiles_f = h->ntiles_b = (int)(h->n_pad / (M::KNT * LC));
            h->fs = M::FS; h->fwd_elem = M::FwdElem::NDBL; h->bwd_elem = M::BwdElem::NDBL;
            if ((rc = dev_alloc<double>(h->ckpt, (size_t)h->nchunks * M::FS, err))) return rc;
            if ((rc = dev_alloc<double>(h->wg, (size_t)h->n_pad * M::NW, err))) return rc;
            if ((rc = dev_alloc<double>(h->s_in, M::FS, err))) return rc;
            if ((rc = dev_alloc<double>(h->g_in, M::FS, err))) return rc;
            if ((rc = dev_alloc<double>(h->f_agg, (size_t)h->ntiles_f * M::FwdElem::NDBL, err))) return rc;
            if ((rc = dev_alloc<double>(h->f_incl, (size_t)h->ntiles_f * M::FwdElem::NDBL, err))) return rc;
            if ((rc = dev_alloc<double>(h->b_agg, (size_t)h->ntiles_b * M::BwdElem::NDBL, err))) return rc;
            if ((rc = dev_alloc<double>(h->b_incl, (size_t)h->ntiles_b * M::BwdElem::NDBL, err))) return rc;
            return ctcrw_grids<M>(h);
        });
        if (rc) return rc;
        if ((rc = dev_alloc<double>(h->tile_llk, h->n_pad / WT, err))) return rc;
        if ((rc = dev_alloc<double>(h->tile_gh, h->n_pad / WT, err))) return rc;
        if ((rc = dev_alloc<unsigned>(h->f_status, h->ntiles_f, err))) return rc;
        if ((rc = dev_alloc<unsigned>(h->b_status, h->ntiles_b, err))) return rc;
        CUDA_TRY(cudaMemset(h->f_status.p, 0, sizeof(unsigned) * std::max(h->ntiles_f, 1)));
        CUDA_TRY(cudaMemset(h->b_status.p, 0, sizeof(unsigned) * std::max(h->ntiles_b, 1)));
        h->grid_f = std::min(h->grid_f, std::max(h->ntiles_f, 1));
        h->grid_b = std::min(h->grid_b, std::max(h->ntiles_b, 1));
    } else {
        h->ntiles_lp = h->n_pad / (SDE_NT * LC);
        if (h->model == SSDE_BM) {
            if (h->n_dim == 1) rc = sde_grid<MODEL_BM, 1>(h);
            else if (h->n_dim == 2) rc = sde_grid<MODEL_BM, 2>(h);
            else rc = sde_grid<MODEL_BM, 3>(h);
        } else {
            rc = (h->n_dim == 1) ? sde_grid<MODEL_OU, 1>(h) : sde_grid<MODEL_OU, 2>(h);
        }
        if (rc) return rc;
        h->grid_lp = (int)std::max<int64_t>(1, std::min<int64_t>(h->grid_lp, h->ntiles_lp));
        h->grid_stream = (int)std::max<int64_t>(1, std::min<int64_t>(h->grid_stream, h->ntiles_lp));
        h->grid_hess = (int)std::max<int64_t>(1, std::min<int64_t>(h->grid_hess, h->ntiles_lp));
        if ((rc = dev_alloc<double>(h->block_llk, std::max(h->grid_lp, h->grid_stream), err))) return rc;
        // streaming kernel: only if every warp-tile of the design is uniform with <= SDE_SMAX slots
        DevBuf shape;
        if ((rc = dev_alloc<int>(shape, 4, err))) return rc;
        CUDA_TRY(cudaMemset(shape.p, 0, 4 * sizeof(int)));
        design_shape_kernel<<<std::max(1, h->num_sms), 256>>>(h->desc.as<WtDesc>(), h->n_pad / WT, shape.as<int>());
        CUDA_TRY(cudaGetLastError());
        int sh[4] = {0, 0, 0, 0};
        CUDA_TRY(cudaMemcpy(sh, shape.p, 4 * sizeof(int), cudaMemcpyDeviceToHost));
        if (sh[3] && h->n_dec > 0) { err = "descriptor flags alias parameters: not supported together with decay terms"; return SSDE_ERR_BAD_ARG; }
        h->sde_stream = h->n_dec == 0 && sh[0] <= SDE_SMAX && sh[1] == 0 && sh[2] <= SDE_KPM;
    }
    return SSDE_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluation pipeline
// ---------------------------------------------------------------------------------------------
// counts a kernel launch and, in profiling mode, records an event in front of it
void mark(ssde_handle* h, cudaStream_t st, const char* name) {
    if (name) ++h->last_launches;
    if (!h->profile) return;
    if ((int)h->pev.size() <= h->pcount) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->pev.push_back(e);
        h->pnames.push_back(name);
    }
    h->pnames[h->pcount] = name;
    cudaEventRecord(h->pev[h->pcount], st);
    ++h->pcount;
}

DesignV2 design_of(const ssde_handle* h) {
    return DesignV2{h->n, h->n_pad, h->desc.as<WtDesc>(), h->val.as<double>(), h->col.as<uint32_t>()};
}

constexpr int TAN_MINB = 1;             // resident CTAs per SM the tangent kernels are compiled for

template <class R>
KalmanArgs<R> ctcrw_args(ssde_handle* h, const double* d_par, const double* d_dir, double* aest) {
    constexpr bool TAN = !std::is_same<R, double>::value;
    KalmanArgs<R> a;
    a.X = design_of(h);
    a.theta = Theta{h->theta.as<double>(), TAN ? h->t_theta_dot.as<double>() : nullptr};
    a.obs = h->obs.as<double>(); a.dt = h->dt.as<double>(); a.flags = h->flags.as<uint8_t>();
    a.track_starts = h->track_starts.as<int64_t>(); a.a0 = h->a0.as<double>(); a.n_tracks = h->n_tracks;
    a.P0 = h->P0; a.Hrow = h->Hrow.as<double>(); a.par = d_par; a.par_dot = TAN ? d_dir : nullptr;
    a.s_in = h->have_s_in ? h->s_in.as<R>() : nullptr;
    a.g_in = h->have_g_in ? h->g_in.as<R>() : nullptr;
    a.mu_zero = h->mu_zero.as<int>();
    a.nchunks = h->nchunks;
    a.ckpt = TAN ? h->t_ckpt.as<R>() : h->ckpt.as<R>();
    a.wg = TAN ? h->t_wg.as<R>() : h->wg.as<R>();
    a.tile_llk = h->tile_llk.as<double>();
    a.tile_gh = TAN ? h->t_tile_gh.as<double>() : h->tile_gh.as<double>();
    a.grad_theta = TAN ? h->t_grad_theta.as<double>() : h->grad_theta.as<double>();
    a.p_theta = h->p_fe + h->p_re;
    a.aest = aest;
    unsigned* cnt = h->counters.as<unsigned>();
    a.fdesc = {h->f_status.as<unsigned>(), (TAN ? h->t_f_agg : h->f_agg).as<double>(), (TAN ? h->t_f_incl : h->f_incl).as<double>(), cnt + 0, cnt + 2, h->epoch};
    a.bdesc = {h->b_status.as<unsigned>(), (TAN ? h->t_b_agg : h->b_agg).as<double>(), (TAN ? h->t_b_incl : h->b_incl).as<double>(), cnt + 1, cnt + 2, h->epoch};
#ifdef SSDE_STATS
    a.fdesc.stats = h->stats.as<unsigned long long>();
    a.bdesc.stats = h->stats.as<unsigned long long>() + 16;
#endif
    a.ntiles = h->ntiles_f;
    a.summary = 0;
    a.tile_lo = 0;
    a.rerun = 1;
    a.llk_bwd = 0;
    return a;
}

// 1 (default): an evaluation that runs the adjoint kernel takes the likelihood terms from that
// kernel's forward recomputation and skips the forward kernel's re-run (phase 4)
#ifndef SSDE_LLK_IN_BWD
#define SSDE_LLK_IN_BWD 1
#endif

// the scan descriptors are keyed by an epoch so that they never need clearing; every kernel that
// walks the tiles (full pass or summary pass) gets a fresh epoch and fresh tickets
int new_scan_epoch(ssde_handle* h, cudaStream_t st) {
    std::string& err = h->err;
    ++h->epoch;
    if (h->epoch >= (1u << 30)) h->epoch = 1;     // status arrays were zeroed at creation; 0 is never a live epoch
    CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 2 * sizeof(unsigned), st));
    return SSDE_OK;
}

constexpr int TAIL_TILES = 4;           // tiles of a tail-only summary pass (4096 rows)

template <class M>
int launch_ctcrw_fwd(ssde_handle* h, const double* d_par, const double* d_dir, cudaStream_t st, double* aest, bool summary,
                     bool tail = false, bool rerun = true) {
    using R = typename M::R;
    constexpr bool TAN = !std::is_same<R, double>::value;
    std::string& err = h->err;
    int rc = new_scan_epoch(h, st);
    if (rc) return rc;
    KalmanArgs<R> a = ctcrw_args<R>(h, d_par, d_dir, aest);
    a.summary = summary ? 1 : 0;
    a.rerun = (rerun || aest) ? 1 : 0;
    a.tile_lo = (summary && tail) ? std::max(h->ntiles_f - TAIL_TILES, 0) : 0;
    mark(h, st, TAN ? "ctcrw_fwd_tangent" : (summary ? (tail ? "ctcrw_fwd_tail" : "ctcrw_fwd_summary") : "ctcrw_fwd"));
    if constexpr (TAN) ctcrw_fwd_kernel<M, M::KNT, TAN_MINB><<<h->grid_f2, M::KNT, sizeof(FwdSmem<M, M::KNT>), st>>>(a);
    else ctcrw_fwd_kernel<M, M::KNT, M::FMINB><<<h->grid_f, M::KNT, sizeof(FwdSmem<M, M::KNT>), st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

template <class M>
int launch_ctcrw_bwd(ssde_handle* h, const double* d_par, const double* d_dir, cudaStream_t st, bool summary, bool tail = false,
                     bool llk = false) {
    using R = typename M::R;
    constexpr bool TAN = !std::is_same<R, double>::value;
    std::string& err = h->err;
    int rc = new_scan_epoch(h, st);
    if (rc) return rc;
    KalmanArgs<R> a = ctcrw_args<R>(h, d_par, d_dir, nullptr);
    a.ntiles = h->ntiles_b;
    a.summary = summary ? 1 : 0;
    a.llk_bwd = llk ? 1 : 0;
    a.tile_lo = (summary && tail) ? std::max(h->ntiles_b - TAIL_TILES, 0) : 0;
    mark(h, st, TAN ? "ctcrw_bwd_tangent" : (summary ? (tail ? "ctcrw_bwd_tail" : "ctcrw_bwd_summary") : "ctcrw_bwd"));
    if constexpr (TAN) ctcrw_bwd_kernel<M, M::KNT, TAN_MINB><<<h->grid_b2, M::KNT, sizeof(BwdSmem<M, M::KNT>), st>>>(a);
    else ctcrw_bwd_kernel<M, M::KNT, M::MINB><<<h->grid_b, M::KNT, sizeof(BwdSmem<M, M::KNT>), st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

int launch_reduce(ssde_handle* h, int order, bool tan, cudaStream_t st) {
    std::string& err = h->err;
    const int nwt = (int)(h->n_pad / WT);
    const double* gh = tan ? h->t_tile_gh.as<double>() : h->tile_gh.as<double>();
    mark(h, st, "reduce_tiles");
    reduce_tiles_kernel<<<RED_BLOCKS, 256, 0, st>>>(h->tile_llk.as<double>(), nwt, order >= 1 ? gh : nullptr,
                                                     tan ? gh + nwt : nullptr, nwt, h->part.as<double>());
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

template <class M>
int launch_ctcrw(ssde_handle* h, const double* d_par, const double* d_dir, int order, cudaStream_t st, double* aest) {
    int rc;
    const bool llk_bwd = SSDE_LLK_IN_BWD && order >= 1 && !aest;
    if ((rc = launch_ctcrw_fwd<M>(h, d_par, d_dir, st, aest, false, false, !llk_bwd))) return rc;
    if (order >= 1 && (rc = launch_ctcrw_bwd<M>(h, d_par, d_dir, st, false, false, llk_bwd))) return rc;
    return launch_reduce(h, order, !std::is_same<typename M::R, double>::value, st);
}

template <int MODEL, int ND, class R>
int launch_sde(ssde_handle* h, const double* d_par, const double* d_dir, int order, cudaStream_t st) {
    constexpr bool TAN = !std::is_same<R, double>::value;
    std::string& err = h->err;
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    SdeArgs a;
    a.X = design_of(h);
    a.theta = Theta{h->theta.as<double>(), TAN ? h->t_theta_dot.as<double>() : nullptr};
    a.obs = h->obs.as<double>(); a.dt = h->dt.as<double>(); a.flags = h->flags.as<uint8_t>();
    a.want_grad = order >= 1;
    a.grad_theta = TAN ? h->t_grad_theta.as<double>() : h->grad_theta.as<double>(); a.p_theta = h->p_fe + h->p_re;
    a.block_llk = h->block_llk.as<double>();
    a.ntiles = h->ntiles_lp;
    if (h->n_dec > 0) {
        DecayArgs dc{h->t_decay.as<double>(), h->dec_of_col.as<int32_t>(), d_par, TAN ? d_dir : nullptr, h->o_dec, h->n_dec,
                     h->grad_decay.as<double>()};
        mark(h, st, TAN ? "sde_decay_tangent" : "sde_decay");
        sde_decay_kernel<MODEL, ND, R><<<TAN ? h->grid_lp2 : h->grid_lp, SDE_NT, 0, st>>>(a, dc);
        CUDA_TRY(cudaGetLastError());
        return SSDE_OK;
    }
    if constexpr (!TAN) {
        if (h->sde_stream) {
            mark(h, st, "sde_stream");
            sde_stream_kernel<MODEL, ND><<<h->grid_stream, SDE_NT, sizeof(SdeStreamSmem), st>>>(a, h->hot);
            CUDA_TRY(cudaGetLastError());
            return SSDE_OK;
        }
    }
    mark(h, st, TAN ? "sde_fused_tangent" : "sde_fused");
    sde_fused_kernel<MODEL, ND, R><<<TAN ? h->grid_lp2 : h->grid_lp, SDE_NT, sizeof(SdeSmem<NP, R>), st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

// d_dir != nullptr: tangent pass along that direction of the parameter vector
int eval_prologue(ssde_handle* h, const double* d_par, const double* d_dir, int order, cudaStream_t st) {
    std::string& err = h->err;
    const int p = h->p_fe + h->p_re;
    CUDA_TRY(cudaMemsetAsync(h->counters.as<unsigned>() + 2, 0, sizeof(unsigned), st));
    if (order >= 1) {
        if (d_dir) CUDA_TRY(cudaMemsetAsync(h->t_grad_theta.p, 0, sizeof(double) * 2 * std::max(p, 1), st));
        else CUDA_TRY(cudaMemsetAsync(h->grad_theta.p, 0, sizeof(double) * std::max(p, 1), st));
        if (h->n_dec > 0) CUDA_TRY(cudaMemsetAsync(h->grad_decay.p, 0, sizeof(double) * 2 * h->n_dec, st));
    }
    mark(h, st, "gather_theta");
    gather_theta_kernel<<<std::max((p + 255) / 256, 1), 256, 0, st>>>(d_par, h->theta.as<double>(), d_dir,
                                                                     d_dir ? h->t_theta_dot.as<double>() : nullptr,
                                                                     h->p_fe, h->p_re, h->o_fe, h->o_re,
                                                                     h->mu_cols.as<int32_t>(), h->n_mu_cols, h->mu_zero.as<int>());
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

int eval_epilogue(ssde_handle* h, const double* d_par, const double* d_dir, int order, double* d_out, double* d_hv, cudaStream_t st) {
    std::string& err = h->err;
    FinArgs f{};
    if (is_kalman(h->model)) {
        f.part_llk = h->part.as<double>(); f.n_part = RED_BLOCKS;
        f.tile_gh = (order >= 1) ? h->part.as<double>() + RED_BLOCKS : nullptr; f.n_gh = RED_BLOCKS;
        f.tile_gh_dot = h->part.as<double>() + 2 * RED_BLOCKS;
    } else {
        f.part_llk = h->block_llk.as<double>(); f.n_part = d_dir ? h->grid_lp2 : ((h->sde_stream && h->n_dec == 0) ? h->grid_stream : h->grid_lp);
        f.tile_gh = nullptr; f.tile_gh_dot = nullptr; f.n_gh = 0;
    }
    f.par = d_par; f.par_dot = d_dir;
    f.grad_theta = d_dir ? h->t_grad_theta.as<double>() : h->grad_theta.as<double>();
    f.S_rowptr = h->S_rowptr.as<uint32_t>(); f.S_col = h->S_col.as<uint32_t>(); f.S_val = h->S_val.as<double>();
    f.sm_off = h->sm_off.as<int32_t>(); f.sb = h->sb.as<double>();
    f.p_fe = h->p_fe; f.p_re = h->p_re; f.n_s = h->n_s; f.npar = h->npar;
    f.o_sig = h->o_sig; f.o_fe = h->o_fe; f.o_ll = h->o_ll; f.o_re = h->o_re;
    f.o_dec = h->o_dec; f.n_dec = h->n_dec; f.grad_decay = h->grad_decay.as<double>();
    f.penalty = 0;
    if (h->has_smooth && h->add_penalty) {
        if (is_kalman(h->model)) f.penalty = 1;                          // ignores include_penalty
        else if (h->include_penalty) f.penalty = 2;
    }
    f.pen_const = h->pen_const;
    f.want_grad = order >= 1;
    f.error = h->counters.as<unsigned>() + 2;
    f.out = d_out;
    f.hv = d_hv;
    mark(h, st, "finalize");
    if (d_dir) finalize_kernel<Dual><<<1, 256, 0, st>>>(f);
    else finalize_kernel<double><<<1, 256, 0, st>>>(f);
    mark(h, st, nullptr);
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

template <class R>
int launch_model(ssde_handle* h, const double* d_par, const double* d_dir, int order, cudaStream_t st, double* aest) {
    if (is_kalman(h->model))
        return with_kalman_model<R>(h, [&](auto m) -> int { return launch_ctcrw<decltype(m)>(h, d_par, d_dir, order, st, aest); });
    if (h->model == SSDE_BM) {
        if (h->n_dim == 1) return launch_sde<MODEL_BM, 1, R>(h, d_par, d_dir, order, st);
        if (h->n_dim == 2) return launch_sde<MODEL_BM, 2, R>(h, d_par, d_dir, order, st);
        return launch_sde<MODEL_BM, 3, R>(h, d_par, d_dir, order, st);
    }
    return (h->n_dim == 1) ? launch_sde<MODEL_OU, 1, R>(h, d_par, d_dir, order, st) : launch_sde<MODEL_OU, 2, R>(h, d_par, d_dir, order, st);
}

int run_eval(ssde_handle* h, const double* d_par, int order, double* d_out, cudaStream_t st, double* aest) {
    std::string& err = h->err;
    if (order < 0 || order > 1) { err = "order must be 0 or 1 here"; return SSDE_ERR_BAD_ARG; }
    if ((h->shard_flags & (SSDE_SHARD_CONT_PREV | SSDE_SHARD_CONT_NEXT)) &&
        !((h->shard_flags & SSDE_SHARD_CONT_PREV) ? h->have_s_in : true)) {
        err = "time shard: use ssde_eval_stage (the incoming state is unknown)";
        return SSDE_ERR_BAD_ARG;
    }
    h->last_launches = 0;
    h->pcount = 0;
    int rc = eval_prologue(h, d_par, nullptr, order, st);
    if (rc) return rc;
    if ((rc = launch_model<double>(h, d_par, nullptr, order, st, aest))) return rc;
    return eval_epilogue(h, d_par, nullptr, order, d_out, nullptr, st);
}

// ---- tangent pass -------------------------------------------------------------------------------
template <class M>
int ctcrw_tan_grids(ssde_handle* h) {
    std::string& err = h->err;
    int rc;
    if ((rc = max_grid(ctcrw_fwd_kernel<M, M::KNT, TAN_MINB>, M::KNT, sizeof(FwdSmem<M, M::KNT>), h->num_sms, err, h->grid_f2))) return rc;
    if ((rc = max_grid(ctcrw_bwd_kernel<M, M::KNT, TAN_MINB>, M::KNT, sizeof(BwdSmem<M, M::KNT>), h->num_sms, err, h->grid_b2))) return rc;
    return SSDE_OK;
}
template <int MODEL, int ND>
int sde_tan_grid(ssde_handle* h) {
    constexpr int NP = (MODEL == MODEL_BM) ? ND + 1 : ND + 2;
    return max_grid(sde_fused_kernel<MODEL, ND, Dual>, SDE_NT, sizeof(SdeSmem<NP, Dual>), h->num_sms, h->err, h->grid_lp2);
}

// Dual-sized work buffers and launch geometry of the tangent kernels (first use only)
int tangent_setup(ssde_handle* h) {
    if (h->tan_ready) return SSDE_OK;
    std::string& err = h->err;
    if (h->shard_flags & (SSDE_SHARD_CONT_PREV | SSDE_SHARD_CONT_NEXT)) {
        err = "Hessian-vector products of a time-sharded track are not built yet";
        return SSDE_ERR_UNSUPPORTED;
    }
    int rc;
    const int p = h->p_fe + h->p_re;
    if ((rc = dev_alloc<double>(h->t_dir, h->npar, err))) return rc;
    if ((rc = dev_alloc<double>(h->t_theta_dot, p, err))) return rc;
    if ((rc = dev_alloc<double>(h->t_grad_theta, 2 * (size_t)p, err))) return rc;
    if ((rc = dev_alloc<double>(h->t_out, 2 * (size_t)h->npar + 2, err))) return rc;
    if (is_kalman(h->model)) {
        rc = with_kalman_model<Dual>(h, [&](auto m) -> int {
            using M = decltype(m);
            int rc;
            if ((rc = dev_alloc<double>(h->t_ckpt, 2 * (size_t)h->nchunks * M::FS, err))) return rc;
            if ((rc = dev_alloc<double>(h->t_wg, 2 * (size_t)h->n_pad * M::NW, err))) return rc;
            if ((rc = dev_alloc<double>(h->t_f_agg, (size_t)h->ntiles_f * M::FwdElem::NDBL, err))) return rc;
            if ((rc = dev_alloc<double>(h->t_f_incl, (size_t)h->ntiles_f * M::FwdElem::NDBL, err))) return rc;
            if ((rc = dev_alloc<double>(h->t_b_agg, (size_t)h->ntiles_b * M::BwdElem::NDBL, err))) return rc;
            if ((rc = dev_alloc<double>(h->t_b_incl, (size_t)h->ntiles_b * M::BwdElem::NDBL, err))) return rc;
            return ctcrw_tan_grids<M>(h);
        });
        if (rc) return rc;
        if ((rc = dev_alloc<double>(h->t_tile_gh, 2 * (size_t)(h->n_pad / WT), err))) return rc;
        h->grid_f2 = std::min(h->grid_f2, std::max(h->ntiles_f, 1));
        h->grid_b2 = std::min(h->grid_b2, std::max(h->ntiles_b, 1));
    } else {
        if (h->model == SSDE_BM) {
            if (h->n_dim == 1) rc = sde_tan_grid<MODEL_BM, 1>(h);
            else if (h->n_dim == 2) rc = sde_tan_grid<MODEL_BM, 2>(h);
            else rc = sde_tan_grid<MODEL_BM, 3>(h);
        } else {
            rc = (h->n_dim == 1) ? sde_tan_grid<MODEL_OU, 1>(h) : sde_tan_grid<MODEL_OU, 2>(h);
        }
        if (rc) return rc;
        // block_llk has grid_lp entries: never launch more tangent CTAs than that
        h->grid_lp2 = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(h->grid_lp2, h->grid_lp), h->ntiles_lp));
    }
    h->tan_ready = true;
    return SSDE_OK;
}

// One tangent pass: d_out[1 + npar + 1] = nllk, gradient, status;  d_hv[npar] = H * dir.
int run_hvp(ssde_handle* h, const double* d_par, const double* d_dir, double* d_out, double* d_hv, cudaStream_t st) {
    int rc = tangent_setup(h);
    if (rc) return rc;
    h->last_launches = 0;
    h->pcount = 0;
    if ((rc = eval_prologue(h, d_par, d_dir, 1, st))) return rc;
    if ((rc = launch_model<Dual>(h, d_par, d_dir, 1, st, nullptr))) return rc;
    return eval_epilogue(h, d_par, d_dir, 1, d_out, d_hv, st);
}

// No C++ exception may cross the C ABI (include/smoothsde_b200.h): std::vector / std::sort of the
// host-side packing can throw std::bad_alloc on a 1e8-row design.
template <class Fn>
int guarded(std::string& err, Fn&& fn) {
    try {
        return fn();
    } catch (const std::bad_alloc&) {
        err = "out of host memory";
    } catch (const std::exception& e) {
        err = std::string("C++ exception: ") + e.what();
    } catch (...) {
        err = "unknown C++ exception";
    }
    return SSDE_ERR_BAD_ARG;
}
}  // namespace
// message for a non-zero device status word (bit 0: look-back time-out; bit 1: F <= 0).  With
// several shards the words are summed, so any non-zero value is a failure and the bits are a hint.
const char* device_failure(double status) {
    const unsigned e = (unsigned)status;
    if (e & 2u) return "device-side failure: innovation variance F <= 0 in the filter (the reference's detF <= 0 branch, "
                       "nllk_ctcrw.hpp:226-228, is not built)";
    return "device-side failure (scan look-back timed out)";
}
namespace {
int check_common(int model, int n_dim, std::string& err) {
    if (model < 0 || model > 4) { err = "Unknown SDE type"; return SSDE_ERR_UNKNOWN_TYPE; }
    const int n_par = sde_par_count(model, n_dim);
    if (n_dim < 1 || n_par > 4) {
        err = "n_dim not supported for this model (n_par <= 4: BM, BM_SSM n_dim <= 3; OU, OU_SSM, CTCRW n_dim <= 2)";
        return SSDE_ERR_UNSUPPORTED;
    }
    return SSDE_OK;
}

}  // namespace

// =============================================================================================
// extern "C"
// =============================================================================================
extern "C" {

const char* ssde_version(void) { return "smoothsde_b200 0.2 (sm_100a)"; }
const char* ssde_create_error(void) { return g_create_error.c_str(); }
const char* ssde_last_error(const ssde_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void ssde_destroy(ssde_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int ssde_n_par(const ssde_handle* h) { return h ? h->npar : -1; }
int ssde_debug_stats(ssde_handle* h, uint64_t out[32], int reset) {
    if (!h || !out) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out, h->stats.p, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(h->stats.p, 0, 32 * sizeof(uint64_t)));
#ifdef SSDE_STATS
    return SSDE_OK;
#else
    err = "library was built without SSDE_STATS";
    return SSDE_ERR_UNSUPPORTED;
#endif
}
// Diagnostics: threshold below which a scan element's linear part counts as zero (models.cuh).
// tol < 0 disables the constant-map shortcut of the look-back altogether; the default restores it.
int ssde_debug_const_map_tol(int device, double tol) {
    if (cudaSetDevice(device) != cudaSuccess) return SSDE_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return SSDE_ERR_CUDA;
    return cudaMemcpyToSymbol(c_const_map_tol, &tol, sizeof(double)) == cudaSuccess ? SSDE_OK : SSDE_ERR_CUDA;
}
int ssde_device(const ssde_handle* h) { return h ? h->device : -1; }
void* ssde_stream(const ssde_handle* h) { return h ? (void*)h->stream : nullptr; }

int ssde_par_layout(const ssde_handle* h, int32_t offsets[4], int32_t sizes[4]) {
    if (!h) return SSDE_ERR_BAD_ARG;
    offsets[0] = h->o_sig; offsets[1] = h->o_fe; offsets[2] = h->o_ll; offsets[3] = h->o_re;
    sizes[0] = (h->o_sig >= 0) ? 1 : 0; sizes[1] = h->p_fe; sizes[2] = h->n_s; sizes[3] = h->p_re;
    return SSDE_OK;
}

int ssde_decay_layout(const ssde_handle* h, int32_t* offset, int32_t* size) {
    if (!h || !offset || !size) return SSDE_ERR_BAD_ARG;
    *offset = h->o_dec; *size = h->n_dec;
    return SSDE_OK;
}

int64_t ssde_padded_rows(int64_t n) { return (n + PAD_ROWS - 1) / PAD_ROWS * PAD_ROWS; }
int ssde_layout_info(int32_t info[4]) {
    if (!info) return SSDE_ERR_BAD_ARG;
    info[0] = LC; info[1] = WT; info[2] = PAD_ROWS; info[3] = (int32_t)sizeof(WtDesc);
    return SSDE_OK;
}

// Host-only: the design of `d` in the device layout (no GPU needed).  Used by hosts that want to
// build an ssde_packed_desc themselves and by the CPU tests of the layout.
static int pack_host_impl(const ssde_desc* d, ssde_host_pack* out) {
    std::string& err = g_create_error;
    err.clear();
    if (!d || !out) { err = "null argument"; return SSDE_ERR_BAD_ARG; }
    std::memset(out, 0, sizeof(*out));
    int rc = check_common(d->model, d->n_dim, err);
    if (rc) return rc;
    const int n_par = sde_par_count(d->model, d->n_dim);
    if (d->n < 1 || d->X_fe.nrow != n_par * d->n || d->X_re.nrow != n_par * d->n) { err = "X_fe / X_re must have n_par * n rows"; return SSDE_ERR_BAD_ARG; }
    Packed pk;
    if ((rc = pack_design(*d, n_par, pk, err))) return rc;
    V2Host v2;
    const int64_t n_pad = ssde_padded_rows(d->n);
    if ((rc = build_v2(pk, d->n, n_pad, n_par, !(d->t_decay && d->t_decay_len > 1), v2, err))) return rc;
    out->n_pad = n_pad;
    out->n_desc = (int64_t)v2.desc.size(); out->n_val = (int64_t)v2.val.size(); out->n_col = (int64_t)v2.col.size();
    out->desc = (ssde_wt_desc*)std::malloc(sizeof(WtDesc) * std::max<size_t>(v2.desc.size(), 1));
    out->val = (double*)std::malloc(sizeof(double) * std::max<size_t>(v2.val.size(), 1));
    out->col = (uint32_t*)std::malloc(sizeof(uint32_t) * std::max<size_t>(v2.col.size(), 1));
    if (!out->desc || !out->val || !out->col) { ssde_pack_free(out); err = "out of memory"; return SSDE_ERR_BAD_ARG; }
    std::memcpy(out->desc, v2.desc.data(), sizeof(WtDesc) * v2.desc.size());
    std::memcpy(out->val, v2.val.data(), sizeof(double) * v2.val.size());
    std::memcpy(out->col, v2.col.data(), sizeof(uint32_t) * v2.col.size());
    return SSDE_OK;
}

int ssde_pack_host(const ssde_desc* d, ssde_host_pack* out) {
    return guarded(g_create_error, [&] { return pack_host_impl(d, out); });
}

void ssde_pack_free(ssde_host_pack* p) {
    if (!p) return;
    std::free(p->desc); std::free(p->val); std::free(p->col);
    std::memset(p, 0, sizeof(*p));
}

static int create_impl(const ssde_desc* d, ssde_handle** out) {
    std::string& err = g_create_error;
    err.clear();
    if (!d || !out) { err = "null argument"; return SSDE_ERR_BAD_ARG; }
    *out = nullptr;
    int rc = check_common(d->model, d->n_dim, err);
    if (rc) return rc;
    const int nd = d->n_dim;
    const int n_par = sde_par_count(d->model, nd);
    const int64_t n = d->n;
    if (n < 1 || !d->ID || !d->times || !d->obs) { err = "ID, times and obs are required"; return SSDE_ERR_BAD_ARG; }
    if (d->X_fe.nrow != n_par * n || d->X_re.nrow != n_par * n) { err = "X_fe / X_re must have n_par * n rows"; return SSDE_ERR_BAD_ARG; }
    const bool user_H = d->H_array && d->H_len > 1;            // H_array.size() > 1, nllk_ctcrw.hpp:203
    if (user_H && !is_kalman(d->model)) { err = "H_array exists for the Kalman models only"; return SSDE_ERR_BAD_ARG; }
    if (user_H && d->H_len != (int64_t)d->n_dim * d->n_dim * d->n) { err = "H_array must be n_dim x n_dim x n"; return SSDE_ERR_BAD_ARG; }
    if (!is_kalman(d->model) && (d->shard_flags & (SSDE_SHARD_CONT_PREV | SSDE_SHARD_CONT_NEXT))) {
        err = "BM / OU shards must hold whole tracks (time-sharding exists for the Kalman models only)";
        return SSDE_ERR_UNSUPPORTED;
    }
    if (cudaSetDevice(d->device) != cudaSuccess) { err = "cudaSetDevice failed: no usable CUDA device (there is no CPU fallback)"; return SSDE_ERR_CUDA; }

    ssde_handle* h = new (std::nothrow) ssde_handle();
    if (!h) { err = "out of memory"; return SSDE_ERR_BAD_ARG; }
    std::unique_ptr<ssde_handle> guard(h);         // released on success; frees the handle on any error or exception
    h->device = d->device; h->model = d->model; h->n_dim = nd; h->n_par = n_par; h->n = n;
    h->n_pad = ssde_padded_rows(n);
    h->p_fe = (int)d->X_fe.ncol; h->p_re = (int)d->X_re.ncol;
    h->include_penalty = d->include_penalty; h->shard_flags = d->shard_flags;
    h->add_penalty = !(d->shard_flags & SSDE_SHARD_NO_PENALTY);
    auto fail = [&](int code) { err = h->err.empty() ? err : h->err; return code; };

    // rows in the permuted order of design.cuh: flags, dt, obs planes; track starts
    const int64_t n_pad = h->n_pad;
    std::vector<uint8_t> flags(n_pad, 0xff);
    std::vector<double> dt(n_pad, 1.0), obs((size_t)n_pad * nd, 0.0);
    std::vector<int64_t> starts;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t pos = row_pos(i);
        const bool start = (i == 0) ? !(d->shard_flags & SSDE_SHARD_CONT_PREV) : (d->ID[i] != d->ID[i - 1]);
        const bool last = (i == n - 1) ? !(d->shard_flags & SSDE_SHARD_CONT_NEXT) : (d->ID[i + 1] != d->ID[i]);
        uint8_t f = (start ? ROW_START : 0) | (last ? ROW_LAST : 0);
        for (int k = 0; k < nd; ++k) {
            const double y = d->obs[(size_t)k * n + i];
            if (std::isnan(y)) f |= (uint8_t)(ROW_NA0 << k);
            else obs[(size_t)k * n_pad + pos] = y;
        }
        if (is_kalman(d->model)) {
            // Only column 0 is tested for NA (nllk_ctcrw.hpp:214, nllk_ou_ssm.hpp:182): a row whose
            // column 0 is observed while another column is NA makes the reference's objective NaN
            // (u and uFu inherit the NA, :221-234).  Reject such data instead of filtering y_k = 0.
            if (!std::isnan(d->obs[i]) && (f & ~(uint8_t)0x07)) {
                err = "row " + std::to_string(i) + ": obs[, 1] is observed but another column is NA; the reference's "
                      "objective is NaN for such a row (only column 1 is tested, nllk_ctcrw.hpp:214)";
                return fail(SSDE_ERR_BAD_ARG);
            }
            f &= (uint8_t)0x07;                                    // NA bits unused
            if (!std::isnan(d->obs[i])) f |= ROW_OBS;              // column 0 only, nllk_ctcrw.hpp:214, nllk_ou_ssm.hpp:182
        }
        flags[pos] = f;
        if (!last) dt[pos] = ((i == n - 1) ? d->t_next : d->times[i + 1]) - d->times[i];
        // the Kalman models never use the dt of a track-start row: it carries the track index instead
        if (start && is_kalman(d->model)) dt[pos] = (double)starts.size();
        if (start) starts.push_back(i);
    }
    if (is_kalman(d->model)) {
        if (!d->a0 || !d->P0) { err = "the Kalman models need a0 and P0"; return fail(SSDE_ERR_BAD_ARG); }
        if (d->n_ID != (int)starts.size()) { err = "nrow(a0) != number of tracks starting on this shard"; return fail(SSDE_ERR_BAD_ARG); }
        const int m = state_means(d->model, nd);
        // default-shaped P0 (CTCRW: block-diagonal with identical 2x2 blocks, R/sde.R:584; SSM: c I,
        // R/sde.R:553) and H = sigma_obs^2 I: the dimensions decouple.  Anything else runs the
        // coupled filter (dense_math.cuh).
        bool shaped = true;
        if (d->model == SSDE_CTCRW) {
            const double p11 = d->P0[0], p12 = d->P0[(size_t)1 * m + 0], p22 = d->P0[(size_t)1 * m + 1];
            for (int r = 0; r < m; ++r)
                for (int c = 0; c < m; ++c) {
                    double want = 0.0;
                    if (r / 2 == c / 2) want = (r % 2 == 0 && c % 2 == 0) ? p11 : ((r % 2 == 1 && c % 2 == 1) ? p22 : p12);
                    if (d->P0[(size_t)c * m + r] != want) shaped = false;
                }
            h->P0.blk = {p11, p12, p22};
        } else {
            const double c0 = d->P0[0];
            for (int r = 0; r < m; ++r)
                for (int c = 0; c < m; ++c)
                    if (d->P0[(size_t)c * m + r] != (r == c ? c0 : 0.0)) shaped = false;
            h->P0.blk = {c0, 0.0, c0};
        }
        for (int r = 0; r < m; ++r)
            for (int c = r; c < m; ++c) {
                if (d->P0[(size_t)c * m + r] != d->P0[(size_t)r * m + c]) { err = "P0 must be symmetric"; return fail(SSDE_ERR_BAD_ARG); }
                h->P0.dense[r * m - r * (r - 1) / 2 + (c - r)] = d->P0[(size_t)c * m + r];
            }
        h->dense = user_H || !shaped;
        if (user_H) {
            // H_array[, , i] is column-major n_dim x n_dim; keep the packed upper triangle as permuted planes
            const int nh = nd * (nd + 1) / 2;
            std::vector<double> Hp((size_t)n_pad * nh, 0.0);
            for (int64_t i = 0; i < n; ++i) {
                const double* Hi = d->H_array + (size_t)i * nd * nd;
                int k = 0;
                for (int r = 0; r < nd; ++r)
                    for (int c = r; c < nd; ++c, ++k) Hp[(size_t)k * n_pad + row_pos(i)] = 0.5 * (Hi[(size_t)c * nd + r] + Hi[(size_t)r * nd + c]);
            }
            if ((rc = dev_upload(h->Hrow, Hp, h->err))) return fail(rc);
        }
        std::vector<double> a0((size_t)starts.size() * m);
        for (size_t k = 0; k < starts.size(); ++k)
            for (int c = 0; c < m; ++c) a0[k * m + c] = d->a0[(size_t)c * starts.size() + k];
        if ((rc = dev_upload(h->a0, a0, h->err))) return fail(rc);
    }
    h->n_tracks = (int)starts.size();
    if (d->t_decay && d->t_decay_len > 1) {
        // decay terms, nllk_sde.hpp:47-59 / R/sde.R:162-177: t_decay has one entry per row of X_re
        if (is_kalman(d->model)) { err = "decay terms exist for nllk_sde models (BM, OU) only"; return fail(SSDE_ERR_BAD_ARG); }
        if (d->t_decay_len != (int64_t)n_par * n) { err = "t_decay must have n_par * n entries (R/sde.R:170)"; return fail(SSDE_ERR_BAD_ARG); }
        if (d->n_col_decay < 1 || !d->col_decay || !d->ind_decay) { err = "col_decay / ind_decay missing"; return fail(SSDE_ERR_BAD_ARG); }
        std::vector<int32_t> doc((size_t)h->p_fe + h->p_re, -1);
        int kmax = 0;
        for (int i = 0; i < d->n_col_decay; ++i) {
            const int c = d->col_decay[i], k = d->ind_decay[i];            // 1-based, nllk_sde.hpp:51-52
            if (c < 1 || c > h->p_re) { err = "'col_decay' should be between 1 and ncol(X_re) (R/sde.R:637-639)"; return fail(SSDE_ERR_BAD_ARG); }
            if (k < 1 || k > 16) { err = "ind_decay must be in 1..16"; return fail(SSDE_ERR_UNSUPPORTED); }
            doc[(size_t)h->p_fe + c - 1] = k - 1;
            kmax = std::max(kmax, k);
        }
        h->n_dec = kmax;
        std::vector<double> td((size_t)n_pad * n_par, 0.0);
        for (int p = 0; p < n_par; ++p)
            for (int64_t i = 0; i < n; ++i) td[(size_t)p * n_pad + row_pos(i)] = d->t_decay[(size_t)p * n + i];
        if ((rc = dev_upload(h->t_decay, td, h->err))) return fail(rc);
        if ((rc = dev_upload(h->dec_of_col, doc, h->err))) return fail(rc);
    }
    Packed pk;
    if ((rc = pack_design(*d, n_par, pk, h->err))) return fail(rc);
    h->nnz = pk.nnz;
    if (is_kalman(d->model)) {
        std::vector<int32_t> mc;
        for (int64_t r = 0; r < n; ++r) {
            uint32_t rp = pk.rowptr[r];
            for (int p = 0; p < nd; ++p) {
                const uint32_t k = (pk.cnt[r] >> (8 * p)) & 255u;
                for (uint32_t j = 0; j < k; ++j) mc.push_back((int32_t)pk.col[rp + j]);
                rp += k;
            }
            if (mc.size() > (1u << 20)) { std::sort(mc.begin(), mc.end()); mc.erase(std::unique(mc.begin(), mc.end()), mc.end()); }
        }
        std::sort(mc.begin(), mc.end());
        mc.erase(std::unique(mc.begin(), mc.end()), mc.end());
        h->n_mu_cols = (int)mc.size();
        if ((rc = dev_upload(h->mu_cols, mc, h->err))) return fail(rc);
    }
    V2Host v2;
    if ((rc = build_v2(pk, n, n_pad, n_par, h->n_dec == 0, v2, h->err))) return fail(rc);
    if ((rc = dev_upload(h->desc, v2.desc, h->err))) return fail(rc);
    if ((rc = dev_upload(h->col, v2.col, h->err))) return fail(rc);
    if ((rc = dev_upload(h->val, v2.val, h->err))) return fail(rc);
    if ((rc = dev_upload(h->obs, obs, h->err))) return fail(rc);
    if ((rc = dev_upload(h->dt, dt, h->err))) return fail(rc);
    if ((rc = dev_upload(h->flags, flags, h->err))) return fail(rc);
    if ((rc = dev_upload(h->track_starts, starts, h->err))) return fail(rc);
    if ((rc = setup_penalty(h, d->S, d->n_smooth, d->ncol_re, h->err))) return fail(rc);
    if ((rc = finish_setup(h))) return fail(rc);
    *out = guard.release();
    return SSDE_OK;
}

int ssde_create(const ssde_desc* d, ssde_handle** out) {
    const int rc = guarded(g_create_error, [&] { return create_impl(d, out); });
    if (rc && out) *out = nullptr;
    return rc;
}

static int create_packed_impl(const ssde_packed_desc* d, ssde_handle** out) {
    std::string& err = g_create_error;
    err.clear();
    if (!d || !out) { err = "null argument"; return SSDE_ERR_BAD_ARG; }
    *out = nullptr;
    int rc = check_common(d->model, d->n_dim, err);
    if (rc) return rc;
    const int n_par = sde_par_count(d->model, d->n_dim);
    if (d->n_par != n_par) { err = "n_par does not match model / n_dim"; return SSDE_ERR_BAD_ARG; }
    if (d->n < 1 || d->n_pad != ssde_padded_rows(d->n)) { err = "n_pad must be ssde_padded_rows(n)"; return SSDE_ERR_BAD_ARG; }
    if (!d->d_desc || !d->d_val || !d->d_col || !d->d_obs || !d->d_dt || !d->d_flags) { err = "null device array"; return SSDE_ERR_BAD_ARG; }
    if (cudaSetDevice(d->device) != cudaSuccess) { err = "cudaSetDevice failed: no usable CUDA device (there is no CPU fallback)"; return SSDE_ERR_CUDA; }
    ssde_handle* h = new (std::nothrow) ssde_handle();
    if (!h) { err = "out of memory"; return SSDE_ERR_BAD_ARG; }
    std::unique_ptr<ssde_handle> guard(h);
    auto fail = [&](int code) { err = h->err.empty() ? err : h->err; return code; };
    h->device = d->device; h->model = d->model; h->n_dim = d->n_dim; h->n_par = n_par; h->n = d->n; h->n_pad = d->n_pad; h->nnz = d->nnz;
    h->p_fe = d->p_fe; h->p_re = d->p_re; h->include_penalty = d->include_penalty; h->shard_flags = d->shard_flags;
    h->add_penalty = !(d->shard_flags & SSDE_SHARD_NO_PENALTY);
    h->desc.p = (void*)d->d_desc; h->col.p = (void*)d->d_col; h->val.p = (void*)d->d_val;
    h->obs.p = (void*)d->d_obs; h->dt.p = (void*)d->d_dt; h->flags.p = (void*)d->d_flags;
    h->n_tracks = d->n_ID;
    h->P0.blk = {d->P0[0], d->P0[1], d->P0[2]};
    if (d->mu_cols && d->n_mu_cols >= 0) {
        std::vector<int32_t> mc(d->mu_cols, d->mu_cols + d->n_mu_cols);
        h->n_mu_cols = d->n_mu_cols;
        if ((rc = dev_upload(h->mu_cols, mc, h->err))) return fail(rc);
    }
    std::vector<int64_t> starts(d->track_starts, d->track_starts + d->n_ID);
    if ((rc = dev_upload(h->track_starts, starts, h->err))) return fail(rc);
    if (is_kalman(d->model)) {
        if (!d->a0) { err = "the Kalman models need a0"; return fail(SSDE_ERR_BAD_ARG); }
        std::vector<double> a0(d->a0, d->a0 + (size_t)d->n_ID * state_means(d->model, d->n_dim));
        if ((rc = dev_upload(h->a0, a0, h->err))) return fail(rc);
    }
    if ((rc = setup_penalty(h, d->S, d->n_smooth, d->ncol_re, h->err))) return fail(rc);
    if ((rc = finish_setup(h))) return fail(rc);
    *out = guard.release();
    return SSDE_OK;
}

int ssde_create_packed(const ssde_packed_desc* d, ssde_handle** out) {
    const int rc = guarded(g_create_error, [&] { return create_packed_impl(d, out); });
    if (rc && out) *out = nullptr;
    return rc;
}

int ssde_eval_device(ssde_handle* h, const double* d_par, int order, double* d_out, void* stream) {
    if (!h || !d_par || !d_out) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    h->timed = (st == h->stream);
    if (h->timed) CUDA_TRY(cudaEventRecord(h->ev0, st));
    int rc = run_eval(h, d_par, order, d_out, st, nullptr);
    if (rc) return rc;
    if (h->timed) CUDA_TRY(cudaEventRecord(h->ev1, st));
    return SSDE_OK;
}

int ssde_shard_elem_doubles(const ssde_handle* h, int which) {
    if (!h || !is_kalman(h->model)) return -1;
    return (which == 0 ? h->fwd_elem : h->bwd_elem) + 1;      // + the constant-map flag
}

int ssde_eval_stage(ssde_handle* h, const double* d_par, int stage, const double* d_elems, int n_shards,
                    int my_shard, double* d_out, void* stream) {
    if (!h || !d_par) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    if (!is_kalman(h->model)) { err = "time-sharded evaluation exists for the Kalman models only"; return SSDE_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const size_t fe = (size_t)h->fwd_elem, be = (size_t)h->bwd_elem;
    int rc = SSDE_OK;
    if (stage == 0 || stage == 3) {
        // parameters -> theta; forward summary (stage 0: tail of the shard only; stage 3: the whole
        // shard); composite element of the shard + constant-map flag -> d_out[fe + 1]
        if (!d_out) return SSDE_ERR_BAD_ARG;
        h->last_launches = 0; h->pcount = 0;
        h->have_s_in = h->have_g_in = false;
        if ((rc = eval_prologue(h, d_par, nullptr, 1, st))) return rc;
        rc = with_kalman_model<double>(h, [&](auto m) -> int {
            using M = decltype(m);
            int rc = launch_ctcrw_fwd<M>(h, d_par, nullptr, st, nullptr, true, stage == 0);
            if (rc) return rc;
            shard_elem_kernel<typename M::FwdElem, FwdOps<M>><<<1, 32, 0, st>>>(h->f_incl.as<double>() + (size_t)(h->ntiles_f - 1) * fe, d_out);
            ++h->last_launches;
            return SSDE_OK;
        });
        if (rc) return rc;
    } else if (stage == 1 || stage == 4) {
        // stage 1: gathered forward elements -> incoming state; forward pass; adjoint summary over the
        // head of the shard -> d_out[be + 1].  stage 4: the adjoint summary again, over the whole shard.
        if (!d_out || (stage == 1 && (!d_elems || my_shard < 0 || my_shard >= n_shards))) return SSDE_ERR_BAD_ARG;
        rc = with_kalman_model<double>(h, [&](auto m) -> int {
            using M = decltype(m);
            int rc;
            if (stage == 1) {
                shard_state_kernel<M><<<1, 32, 0, st>>>(d_elems, n_shards, my_shard, h->P0, h->s_in.as<double>());
                ++h->last_launches;
                h->have_s_in = true;
                // the adjoint pass of stage 2 always follows: it sums the likelihood terms
                if ((rc = launch_ctcrw_fwd<M>(h, d_par, nullptr, st, nullptr, false, false, !SSDE_LLK_IN_BWD))) return rc;
            }
            if ((rc = launch_ctcrw_bwd<M>(h, d_par, nullptr, st, true, stage == 1))) return rc;
            shard_elem_kernel<typename M::BwdElem, BwdOps<M>><<<1, 32, 0, st>>>(h->b_incl.as<double>() + (size_t)(h->ntiles_b - 1) * be, d_out);
            ++h->last_launches;
            return SSDE_OK;
        });
        if (rc) return rc;
    } else if (stage == 2) {
        // gathered adjoint elements -> incoming adjoint; adjoint pass; finalize -> d_out[1 + n_par + 1]
        if (!d_elems || !d_out || my_shard < 0 || my_shard >= n_shards) return SSDE_ERR_BAD_ARG;
        rc = with_kalman_model<double>(h, [&](auto m) -> int {
            using M = decltype(m);
            shard_adjoint_kernel<M><<<1, 32, 0, st>>>(d_elems, n_shards, my_shard, h->g_in.as<double>());
            ++h->last_launches;
            h->have_g_in = true;
            return launch_ctcrw_bwd<M>(h, d_par, nullptr, st, false, false, SSDE_LLK_IN_BWD != 0);
        });
        if (rc) return rc;
        if ((rc = launch_reduce(h, 1, false, st))) return rc;
        if ((rc = eval_epilogue(h, d_par, nullptr, 1, d_out, nullptr, st))) return rc;
    } else {
        err = "stage must be 0 .. 4";
        return SSDE_ERR_BAD_ARG;
    }
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

int ssde_check(ssde_handle* h) {
    if (!h) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned e = 0;
    CUDA_TRY(cudaMemcpy(&e, h->counters.as<unsigned>() + 2, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (e) { err = device_failure((double)e); return SSDE_ERR_NUMERIC; }
    return SSDE_OK;
}

int ssde_hvp_device(ssde_handle* h, const double* d_par, const double* d_dir, double* d_out, double* d_hv, void* stream) {
    if (!h || !d_par || !d_dir || !d_out || !d_hv) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    h->timed = (st == h->stream);
    if (h->timed) CUDA_TRY(cudaEventRecord(h->ev0, st));
    int rc = run_hvp(h, d_par, d_dir, d_out, d_hv, st);
    if (rc) return rc;
    if (h->timed) CUDA_TRY(cudaEventRecord(h->ev1, st));
    return SSDE_OK;
}

int ssde_hess_cols_device(ssde_handle* h, const double* d_par, int first, int count, double* d_out, double* d_hess, void* stream) {
    if (!h || !d_par || !d_out || !d_hess) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    if (first < 0 || count < 0 || first + count > h->npar) { err = "Hessian columns out of range"; return SSDE_ERR_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = tangent_setup(h);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    h->timed = (st == h->stream);
    if (h->timed) CUDA_TRY(cudaEventRecord(h->ev0, st));
    int launches = 0;
    for (int j = 0; j < count; ++j) {
        unit_vector_kernel<<<(h->npar + 255) / 256, 256, 0, st>>>(h->t_dir.as<double>(), h->npar, first + j);
        if ((rc = run_hvp(h, d_par, h->t_dir.as<double>(), d_out, d_hess + (size_t)j * h->npar, st))) return rc;
        launches += h->last_launches + 1;
    }
    h->last_launches = launches;
    if (h->timed) CUDA_TRY(cudaEventRecord(h->ev1, st));
    return SSDE_OK;
}

int ssde_hess_theta_device(ssde_handle* h, const double* d_par, double* d_hess, void* stream) {
    if (!h || !d_par || !d_hess) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    if (is_kalman(h->model) || h->n_dec > 0 || !h->sde_stream) {
        err = "the one-pass X'WX Hessian exists for BM / OU without decay terms and designs with uniform warp-tiles "
              "(<= 32 slots, <= 12 per SDE parameter); use ssde_hess_cols_device (tangent passes) otherwise";
        return SSDE_ERR_UNSUPPORTED;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int p = h->p_fe + h->p_re;
    h->last_launches = 0;
    h->pcount = 0;
    int rc;
    if ((rc = eval_prologue(h, d_par, nullptr, 0, st))) return rc;
    CUDA_TRY(cudaMemsetAsync(d_hess, 0, sizeof(double) * (size_t)p * p, st));
    SdeArgs a;
    a.X = design_of(h);
    a.theta = Theta{h->theta.as<double>(), nullptr};
    a.obs = h->obs.as<double>(); a.dt = h->dt.as<double>(); a.flags = h->flags.as<uint8_t>();
    a.want_grad = 0; a.grad_theta = nullptr; a.p_theta = p; a.block_llk = nullptr; a.ntiles = h->ntiles_lp;
    mark(h, st, "sde_hess");
    auto launch = [&](auto kern, size_t smem) {
        kern<<<h->grid_hess, SDE_NT, smem, st>>>(a, h->hess_hot, d_hess);
    };
    if (h->model == SSDE_BM) {
        if (h->n_dim == 1) launch(sde_hess_kernel<MODEL_BM, 1>, sizeof(SdeHessSmem<2>));
        else if (h->n_dim == 2) launch(sde_hess_kernel<MODEL_BM, 2>, sizeof(SdeHessSmem<3>));
        else launch(sde_hess_kernel<MODEL_BM, 3>, sizeof(SdeHessSmem<4>));
    } else {
        if (h->n_dim == 1) launch(sde_hess_kernel<MODEL_OU, 1>, sizeof(SdeHessSmem<3>));
        else launch(sde_hess_kernel<MODEL_OU, 2>, sizeof(SdeHessSmem<4>));
    }
    CUDA_TRY(cudaGetLastError());
    // smoothing penalty: lambda_i S_i on the coeff_re block (nllk_sde.hpp:118-119), added by the shard that owns it
    if (h->has_smooth && h->add_penalty && h->include_penalty) {
        mark(h, st, "hess_penalty");
        hess_penalty_kernel<<<std::max((h->p_re + 127) / 128, 1), 128, 0, st>>>(
            d_par, h->o_ll, h->n_s, h->sm_off.as<int32_t>(), h->S_rowptr.as<uint32_t>(), h->S_col.as<uint32_t>(),
            h->S_val.as<double>(), h->p_fe, h->p_re, d_hess);
        CUDA_TRY(cudaGetLastError());
    }
    return SSDE_OK;
}

int ssde_hvp(ssde_handle* h, const double* par, int n_dir, const double* dirs, double* nllk, double* grad, double* hv) {
    if (!h || !par || !dirs || !hv || n_dir < 1) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = tangent_setup(h);
    if (rc) return rc;
    const int np = h->npar;
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->par.p, par, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    std::vector<double> out((size_t)np + 2);
    h->timed = true;
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    int launches = 0;
    for (int j = 0; j < n_dir; ++j) {
        CUDA_TRY(cudaMemcpyAsync(h->t_dir.p, dirs + (size_t)j * np, sizeof(double) * np, cudaMemcpyHostToDevice, st));
        if ((rc = run_hvp(h, h->par.as<double>(), h->t_dir.as<double>(), h->t_out.as<double>(), h->t_out.as<double>() + np + 2, st))) return rc;
        launches += h->last_launches;
        CUDA_TRY(cudaMemcpyAsync(hv + (size_t)j * np, h->t_out.as<double>() + np + 2, sizeof(double) * np, cudaMemcpyDeviceToHost, st));
    }
    h->last_launches = launches;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(out.data(), h->t_out.p, sizeof(double) * (np + 2), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (nllk) *nllk = out[0];
    if (grad) std::memcpy(grad, out.data() + 1, sizeof(double) * np);
    if (out[1 + np] != 0.0) { err = device_failure(out[1 + np]); return SSDE_ERR_NUMERIC; }
    return SSDE_OK;
}

int ssde_eval(ssde_handle* h, const double* par, int order, double* nllk, double* grad, double* hess) {
    if (!h || !par || !nllk) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    if (order < 0 || order > 2) { err = "order must be 0, 1 or 2"; return SSDE_ERR_BAD_ARG; }
    if (order >= 1 && !grad) { err = "grad buffer required for order >= 1"; return SSDE_ERR_BAD_ARG; }
    if (order == 2 && !hess) { err = "hess buffer required for order 2"; return SSDE_ERR_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    double* hp = h->h_pinned;
    double* ho = h->h_pinned + h->npar;
    std::memcpy(hp, par, sizeof(double) * h->npar);
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->par.p, hp, sizeof(double) * h->npar, cudaMemcpyHostToDevice, st));
    if (order == 2) {
        // joint Hessian (what obj$he returns, R/sde.R:1363): npar tangent passes, column j = H e_j
        const int np = h->npar;
        int rc = tangent_setup(h);
        if (rc) return rc;
        if (!h->t_hess.p && (rc = dev_alloc<double>(h->t_hess, (size_t)np * np, err))) return rc;
        if ((rc = ssde_hess_cols_device(h, h->par.as<double>(), 0, np, h->t_out.as<double>(), h->t_hess.as<double>(), nullptr))) return rc;
        CUDA_TRY(cudaMemcpyAsync(ho, h->t_out.p, sizeof(double) * (np + 2), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(hess, h->t_hess.p, sizeof(double) * (size_t)np * np, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        *nllk = ho[0];
        std::memcpy(grad, ho + 1, sizeof(double) * np);
        for (int i = 0; i < np; ++i)                    // symmetrise (columns come from independent passes)
            for (int j = i + 1; j < np; ++j) {
                const double m = 0.5 * (hess[(size_t)j * np + i] + hess[(size_t)i * np + j]);
                hess[(size_t)j * np + i] = hess[(size_t)i * np + j] = m;
            }
        if (ho[1 + np] != 0.0) { err = device_failure(ho[1 + np]); return SSDE_ERR_NUMERIC; }
        return SSDE_OK;
    }
    h->timed = true;
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    int rc = run_eval(h, h->par.as<double>(), order, h->out.as<double>(), st, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    const size_t nout = (order >= 1) ? (size_t)h->npar + 2 : 1;
    CUDA_TRY(cudaMemcpyAsync(ho, h->out.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, st));
    if (order == 0) CUDA_TRY(cudaMemcpyAsync(ho + 1 + h->npar, h->out.as<double>() + 1 + h->npar, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *nllk = ho[0];
    if (order >= 1) std::memcpy(grad, ho + 1, sizeof(double) * h->npar);
    if (ho[1 + h->npar] != 0.0) { err = device_failure(ho[1 + h->npar]); return SSDE_ERR_NUMERIC; }
    return SSDE_OK;
}

int ssde_report(ssde_handle* h, double* aest_all) {
    if (!h || !aest_all) return SSDE_ERR_BAD_ARG;
    std::string& err = h->err;
    if (!is_kalman(h->model)) { err = "REPORT(aest_all) exists for the Kalman models only"; return SSDE_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(h->device));
    const int m = state_means(h->model, h->n_dim);
    if (!h->aest.p) { int rc = dev_alloc<double>(h->aest, (size_t)h->n_pad * m, err); if (rc) return rc; }
    int rc = run_eval(h, h->par.as<double>(), 0, h->out.as<double>(), h->stream, h->aest.as<double>());
    if (rc) return rc;
    std::vector<double> tmp((size_t)h->n * m);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->aest.p, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < h->n; ++i)
        for (int c = 0; c < m; ++c) aest_all[(size_t)c * h->n + i] = tmp[(size_t)i * m + c];
    return SSDE_OK;
}

int ssde_simulate_ctcrw(int device, int64_t n_tracks, int64_t n_steps, const double* d_times,
                        const double* d_tau, const double* d_nu, const double* d_mu,
                        const double* d_e1, const double* d_e2, double* d_z, void* stream) {
    std::string& err = g_create_error;
    err.clear();
    if (n_tracks < 1 || n_steps < 1 || !d_times || !d_tau || !d_nu || !d_e1 || !d_e2 || !d_z) { err = "bad argument"; return SSDE_ERR_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    const int nt = 32;
    ctcrw_sim_kernel<<<(unsigned)((n_tracks + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
        n_tracks, n_steps, d_times, d_tau, d_nu, d_mu, d_e1, d_e2, d_z);
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

int ssde_simulate_ou(int device, int64_t n_tracks, int64_t n_steps, const double* d_times, const double* d_mu,
                     const double* d_tau, const double* d_kappa, const double* d_e, double* d_z, void* stream) {
    std::string& err = g_create_error;
    err.clear();
    if (n_tracks < 1 || n_steps < 1 || !d_times || !d_mu || !d_tau || !d_kappa || !d_e || !d_z) { err = "bad argument"; return SSDE_ERR_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    const int nt = 32;
    ou_sim_kernel<<<(unsigned)((n_tracks + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
        n_tracks, n_steps, d_times, d_mu, d_tau, d_kappa, d_e, d_z);
    CUDA_TRY(cudaGetLastError());
    return SSDE_OK;
}

double ssde_last_eval_ms(ssde_handle* h) {
    if (!h || !h->timed) return -1.0;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

int ssde_last_eval_launches(const ssde_handle* h) { return h ? h->last_launches : 0; }

int ssde_launch_info(const ssde_handle* h, int32_t info[8]) {
    if (!h || !info) return SSDE_ERR_BAD_ARG;
    info[0] = h->num_sms; info[1] = h->grid_f; info[2] = h->grid_b; info[3] = h->grid_lp;
    info[4] = h->ntiles_f; info[5] = h->ntiles_b; info[6] = (int32_t)h->ntiles_lp; info[7] = h->n_tracks;
    return SSDE_OK;
}

int ssde_set_profile(ssde_handle* h, int on) {
    if (!h) return SSDE_ERR_BAD_ARG;
    h->profile = on != 0;
    h->pcount = 0;
    return SSDE_OK;
}

int ssde_last_kernel_times(ssde_handle* h, int cap, float* ms, const char** names) {
    if (!h || !h->profile || h->pcount < 2) return 0;
    if (cudaEventSynchronize(h->pev[h->pcount - 1]) != cudaSuccess) return 0;
    int k = 0;
    for (; k + 1 < h->pcount && k < cap; ++k) {
        float t = 0.f;
        cudaEventElapsedTime(&t, h->pev[k], h->pev[k + 1]);
        ms[k] = t;
        if (names) names[k] = h->pnames[k];
    }
    return k;
}

}  // extern "C"
