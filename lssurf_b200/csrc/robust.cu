// Robust statistics of the edit loop: RDE by exact radix selection, sigma_extra by bounded Brent,
// the three-sigma edit mask and the data-row weights.
//
// Replaces RDE (LSsurf/RDE.py:10-18), calc_sigma_extra (LSsurf/calc_sigma_extra.py:28-44, which calls
// scipy.optimize.minimize_scalar(method='bounded')), and the mask / weight updates of iterate_fit
// (LSsurf/smooth_fit.py:123-135, 159-163).
//
// RDE needs only the four order statistics np.interp touches.  They are found
// exactly by an MSD radix select over the order-preserving 64-bit image of the fp64 values:
// 8 passes of 8 bits; each pass is one streaming histogram kernel over the points (the scaled
// residual r/sqrt(s^2+sigma^2) is recomputed on the fly with correctly rounded fp64 operations, so
// it is bit-identical to NumPy's) followed by a one-block kernel that picks each target's bin.
// When points are sharded over GPUs the histograms are summed with one NCCL all-reduce per pass.
#include <math.h>
#include <stdlib.h>

#include <cmath>

#include "common.cuh"

namespace lss {

constexpr int SEL_MAX = 4;       // order statistics selected together
constexpr int SEL_BINS = 256;
constexpr int SEL_BLOCK = 256;

struct SelectState {
    unsigned long long prefix[SEL_MAX];   // key bits decided so far (left aligned value >> remaining)
    unsigned long long rank[SEL_MAX];     // rank of the target inside its prefix bucket
    int slot[SEL_MAX];                    // histogram slot of each target (targets sharing a prefix share a slot)
    int n_slots;
    int n_targets;
    unsigned long long slot_prefix[SEL_MAX];
};

__device__ __forceinline__ unsigned long long f64_key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// v = r / sqrt(s1^2 + sigma^2) with the operation order of calc_sigma_extra.py:39 (each op rounded).
__device__ __forceinline__ double scaled_value(const double *__restrict__ r, const double *__restrict__ sigma, double s1sq, int64_t i) {
    const double ri = r[i];
    if (!sigma) return ri;
    const double sg = sigma[i];
    return __ddiv_rn(ri, __dsqrt_rn(__dadd_rn(s1sq, __dmul_rn(sg, sg))));
}

// Count of selected finite values (pass 0 only) and per-slot digit histograms of the current pass.
// With `keys` the population is the candidate array keys[0 .. *n_keys) (already order-preserving images) instead of r / sigma / mask.
__global__ void __launch_bounds__(SEL_BLOCK)
select_hist_kernel(const double *__restrict__ r, const double *__restrict__ sigma, const unsigned char *__restrict__ mask,
                   double s1sq, int64_t n, int pass, const SelectState *__restrict__ state,
                   unsigned long long *__restrict__ hist /*[SEL_MAX][SEL_BINS]*/, unsigned long long *__restrict__ count,
                   const unsigned long long *__restrict__ keys = nullptr, const unsigned long long *__restrict__ n_keys = nullptr) {
    __shared__ unsigned int h[SEL_MAX][SEL_BINS];
    __shared__ unsigned long long sp[SEL_MAX];
    __shared__ int ns;
    for (int i = threadIdx.x; i < SEL_MAX * SEL_BINS; i += SEL_BLOCK) (&h[0][0])[i] = 0;
    if (threadIdx.x < SEL_MAX) sp[threadIdx.x] = state->slot_prefix[threadIdx.x];
    if (threadIdx.x == 0) ns = pass == 0 ? 1 : state->n_slots;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    unsigned long long cnt = 0;
    if (keys) n = (int64_t)*n_keys;
    // warp-uniform trip count so that the match below always sees the full warp
    for (int64_t i0 = (int64_t)blockIdx.x * SEL_BLOCK + (threadIdx.x & ~31); i0 < n; i0 += (int64_t)gridDim.x * SEL_BLOCK) {
        const int64_t i = i0 + lane_id();
        bool use = i < n && (keys || !mask || mask[i]);
        double v = 0.0;
        if (use && !keys) {
            v = scaled_value(r, sigma, s1sq, i);
            use = isfinite(v);   // RDE drops non-finite values (RDE.py:12)
        }
        int slot = -1;
        unsigned int digit = 0;
        if (use) {
            ++cnt;
            const unsigned long long k = keys ? keys[i] : f64_key(v);
            digit = (unsigned int)(k >> shift) & 0xffu;
            if (pass == 0) slot = 0;
            else {
                const unsigned long long pre = k >> (shift + 8);
                for (int sidx = 0; sidx < ns; ++sidx)
                    if (pre == sp[sidx]) slot = sidx;
            }
        }
        // warp-aggregated shared-memory histogram: values cluster (same sign/exponent) in the top digits
        const unsigned int tag = slot < 0 ? 0xffffffffu : ((unsigned int)slot << 8) | digit;
        const unsigned peers = __match_any_sync(0xffffffffu, tag);
        if (slot >= 0 && lane_id() == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[slot][digit], (unsigned int)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ns * SEL_BINS; i += SEL_BLOCK) {
        const unsigned int c = (&h[0][0])[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
    if (pass == 0 && !keys) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane_id() == 0 && cnt) atomicAdd(count, cnt);
    }
}

// Pass 0 only: turn the global count n into the target ranks (RDE.py:15-16).
// targets: rank list for p = 0.16 then 0.84, each either one rank (clamped ends) or (j, j+1).
struct RdePlan {
    unsigned long long n;
    int n_targets;
    int lo_idx[2], hi_idx[2];  // indices into the target list for s[j], s[j+1] of each percentile
    double frac[2];            // u - (j + 0.5)
    int clamp[2];              // 1 => percentile is s[lo_idx] itself
};

__global__ void select_plan_kernel(const unsigned long long *__restrict__ count, SelectState *__restrict__ state, RdePlan *__restrict__ plan) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long n = *count;
    plan->n = n;
    int nt = 0;
    if (n >= 2) {
        const double pcts[2] = {0.16, 0.84};
        for (int q = 0; q < 2; ++q) {
            const double u = __dmul_rn(pcts[q], (double)n);      // np.array([0.16, 0.84]) * n
            if (u <= 0.5) {                                       // left of the first abscissa -> fp[0]
                plan->clamp[q] = 1; plan->lo_idx[q] = plan->hi_idx[q] = nt; state->rank[nt++] = 0; plan->frac[q] = 0.0;
            } else if (u >= (double)n - 0.5) {                    // at / right of the last abscissa -> fp[-1]
                plan->clamp[q] = 1; plan->lo_idx[q] = plan->hi_idx[q] = nt; state->rank[nt++] = n - 1; plan->frac[q] = 0.0;
            } else {
                const unsigned long long j = (unsigned long long)floor(u - 0.5);
                plan->clamp[q] = 0;
                plan->frac[q] = __dsub_rn(u, (double)j + 0.5);   // x - xp[j]
                plan->lo_idx[q] = nt; state->rank[nt++] = j;
                plan->hi_idx[q] = nt; state->rank[nt++] = j + 1;
            }
        }
    }
    plan->n_targets = nt;
    state->n_targets = nt;
    state->n_slots = 1;
    state->slot_prefix[0] = 0;
    for (int t = 0; t < SEL_MAX; ++t) { state->prefix[t] = 0; state->slot[t] = 0; }
}

// One block: for every target find the bin of its slot's histogram holding its rank, extend its prefix,
// then regroup targets by prefix into slots for the next pass.  Clears the histograms.
__global__ void __launch_bounds__(SEL_BINS)
select_step_kernel(SelectState *__restrict__ state, unsigned long long *__restrict__ hist) {
    __shared__ unsigned long long wsum[SEL_BINS / 32];
    const int nt = state->n_targets;
    const int b = threadIdx.x;
    for (int t = 0; t < nt; ++t) {
        const int sl = state->slot[t];
        const unsigned long long rk = state->rank[t];
        const unsigned long long pre = state->prefix[t];
        // exclusive scan of the slot's 256-bin histogram
        const unsigned long long c = hist[sl * SEL_BINS + b];
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane_id() >= (unsigned)o) inc += up;
        }
        if (lane_id() == 31) wsum[b >> 5] = inc;
        __syncthreads();                       // also orders the reads of state above before the write below
        unsigned long long base = 0;
        for (int w = 0; w < (b >> 5); ++w) base += wsum[w];
        const unsigned long long lo = base + inc - c, hi = base + inc;
        if (rk >= lo && rk < hi) {
            state->rank[t] = rk - lo;
            state->prefix[t] = (pre << 8) | (unsigned long long)b;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < SEL_MAX * SEL_BINS; i += SEL_BINS) hist[i] = 0;
    if (threadIdx.x == 0) {
        int ns = 0;
        for (int t = 0; t < nt; ++t) {
            int found = -1;
            for (int s2 = 0; s2 < ns; ++s2)
                if (state->slot_prefix[s2] == state->prefix[t]) found = s2;
            if (found < 0) { found = ns; state->slot_prefix[ns++] = state->prefix[t]; }
            state->slot[t] = found;
        }
        state->n_slots = ns;
    }
}

__global__ void select_finish_kernel(const SelectState *__restrict__ state, const RdePlan *__restrict__ plan, double *__restrict__ out /*[2]: rde, n*/) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    out[1] = (double)plan->n;
    if (plan->n < 2) { out[0] = nan(""); return; }
    double q[2];
    for (int k = 0; k < 2; ++k) {
        const double lo = key_f64(state->prefix[plan->lo_idx[k]]);
        if (plan->clamp[k]) q[k] = lo;
        else {
            const double hi = key_f64(state->prefix[plan->hi_idx[k]]);
            // np.interp: slope = (fp[j+1]-fp[j]) / (xp[j+1]-xp[j]) with unit spacing; slope*(x-xp[j]) + fp[j]
            const double slope = __ddiv_rn(__dsub_rn(hi, lo), 1.0);
            q[k] = __dadd_rn(__dmul_rn(slope, plan->frac[k]), lo);
        }
    }
    out[0] = __ddiv_rn(__dsub_rn(q[1], q[0]), 2.0);   // (LH[1]-LH[0])/2.   RDE.py:18
}

// ---- edit mask / weights ---------------------------------------------------------------------

__global__ void __launch_bounds__(256)
edit_mask_kernel(const double *__restrict__ r, const double *__restrict__ sigma, const double *__restrict__ sx, int64_t n,
                 double n_sigma, unsigned char *__restrict__ in_tse, unsigned long long *__restrict__ counts /*[0]=in,[1]=near*/,
                 unsigned long long *__restrict__ min_bits) {
    unsigned long long c_in = 0, c_near = 0;
    double mn = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double sg = sigma[i], se = sx ? sx[i] : 0.0;
        // sigma_aug = sqrt(sigma**2 + sigma_extra**2); in_TSE = abs(r/sigma_aug) < 3.0   (smooth_fit.py:159-163)
        const double aug = __dsqrt_rn(__dadd_rn(__dmul_rn(sg, sg), __dmul_rn(se, se)));
        const double a = fabs(__ddiv_rn(r[i], aug));
        const bool in = a < n_sigma;            // NaN -> false, as in NumPy
        in_tse[i] = in ? 1 : 0;
        if (in) { ++c_in; mn = fmin(mn, sg); }
        if (fabs(a - n_sigma) < 1e-6) ++c_near;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_in += __shfl_xor_sync(0xffffffffu, c_in, o);
        c_near += __shfl_xor_sync(0xffffffffu, c_near, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane_id() == 0) {
        if (c_in) atomicAdd(counts, c_in);
        if (c_near) atomicAdd(counts + 1, c_near);
        if (mn < INFINITY) atomicMin(min_bits, f64_key(mn));   // order-preserving image of the double
    }
}

__global__ void __launch_bounds__(256)
data_weights_kernel(const double *__restrict__ sigma, const double *__restrict__ sx, const unsigned char *__restrict__ mask, int64_t n,
                    double *__restrict__ w) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double sg = sigma[i];
        double e2 = sg * sg;
        if (sx) e2 += sx[i] * sx[i];               // sigma_extra_relax on the last pass (smooth_fit.py:124-127)
        w[i] = (!mask || mask[i]) ? 1.0 / e2 : 0.0;
    }
}


// ---- two-pass candidate selection --------------------------------------------------------------
// A full 8-pass radix select streams all points eight times per RDE evaluation.  Instead: ONE pass builds a histogram
// over a monotone, magnitude-resolving binning of the values (sign, 32 binary exponents around 1, 6 mantissa bits:
// 4098 bins, ~0.4 % of a normal sample per bin near the 16th / 84th percentiles); a second pass appends the keys of the
// points inside the (at most 4) bins that hold the target ranks to a candidate array; the exact 8-pass select then runs
// on the candidates (~10^-2 of the points) with ranks taken relative to them.  Binning is monotone in the value, so
// every point outside the target bins is on a known side of every target: the result is the same order statistic.
constexpr int BIN_EXP_LO = -20, BIN_EXP_N = 32, BIN_MANT_BITS = 6;
constexpr int BIN_HALF = BIN_EXP_N << BIN_MANT_BITS;          // 2048 magnitude classes per sign
constexpr int BIN_COUNT = 2 * BIN_HALF + 2;                   // negatives (descending magnitude), tiny-, tiny+, positives

__device__ __forceinline__ int value_bin(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    const bool neg = (u >> 63) != 0;
    const int e = (int)((u >> 52) & 0x7ff) - 1023 - BIN_EXP_LO;              // 0 .. BIN_EXP_N - 1 inside the resolved range
    int mag;                                                                 // -1: tiny, 0 .. BIN_HALF - 1
    if (e < 0) mag = -1;
    else if (e >= BIN_EXP_N) mag = BIN_HALF - 1;
    else mag = (e << BIN_MANT_BITS) | (int)((u >> (52 - BIN_MANT_BITS)) & ((1u << BIN_MANT_BITS) - 1));
    return neg ? (BIN_HALF - 1 - mag) : (BIN_HALF + 2 + mag);                // tiny-: BIN_HALF, tiny+: BIN_HALF + 1
}

struct BinPlan {
    int n_bins;                    // distinct target bins (ascending)
    int bins[SEL_MAX];
};

__global__ void __launch_bounds__(SEL_BLOCK)
select_binhist_kernel(const double *__restrict__ r, const double *__restrict__ sigma, const unsigned char *__restrict__ mask,
                      double s1sq, int64_t n, unsigned long long *__restrict__ hist /*[BIN_COUNT]*/, unsigned long long *__restrict__ count) {
    __shared__ unsigned int h[BIN_COUNT];
    for (int i = threadIdx.x; i < BIN_COUNT; i += SEL_BLOCK) h[i] = 0;
    __syncthreads();
    unsigned long long cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * SEL_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * SEL_BLOCK) {
        if (mask && !mask[i]) continue;
        const double v = scaled_value(r, sigma, s1sq, i);
        if (!isfinite(v)) continue;
        ++cnt;
        atomicAdd(&h[value_bin(v)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BIN_COUNT; i += SEL_BLOCK) {
        const unsigned int c = h[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane_id() == 0 && cnt) atomicAdd(count, cnt);
}

// One block: bins of the target ranks; ranks become ranks inside the union of the target bins; clears the
// candidate counter and the 8-bit histograms of the exact select that follows.
__global__ void __launch_bounds__(SEL_BLOCK)
select_binplan_kernel(SelectState *__restrict__ state, const unsigned long long *__restrict__ hist, BinPlan *__restrict__ bp,
                      unsigned long long *__restrict__ n_cand, unsigned long long *__restrict__ hist8) {
    __shared__ unsigned long long start[BIN_COUNT + 1];
    for (int i = threadIdx.x; i < SEL_MAX * SEL_BINS; i += SEL_BLOCK) hist8[i] = 0;
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < BIN_COUNT; ++b) { start[b] = acc; acc += hist[b]; }
        start[BIN_COUNT] = acc;
        const int nt = state->n_targets;
        int nb = 0;
        unsigned long long before = 0;          // candidates in target bins below the current one
        for (int t = 0; t < nt; ++t) {          // target ranks ascend (0.16 pair, then 0.84 pair)
            const unsigned long long rk = state->rank[t];
            int lo = 0, hi = BIN_COUNT;         // last bin with start <= rk
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (start[mid] <= rk) lo = mid; else hi = mid; }
            while (lo + 1 < BIN_COUNT && start[lo + 1] <= rk) ++lo;
            if (nb == 0 || bp->bins[nb - 1] != lo) {
                if (nb > 0) before += hist[bp->bins[nb - 1]];
                bp->bins[nb++] = lo;
            }
            state->rank[t] = rk - start[lo] + before;
        }
        bp->n_bins = nb;
        *n_cand = 0;
    }
}

__global__ void __launch_bounds__(SEL_BLOCK)
select_compact_kernel(const double *__restrict__ r, const double *__restrict__ sigma, const unsigned char *__restrict__ mask,
                      double s1sq, int64_t n, const BinPlan *__restrict__ bp, unsigned long long *__restrict__ cand,
                      unsigned long long *__restrict__ n_cand) {
    const int nb = bp->n_bins;
    int tb[SEL_MAX];
#pragma unroll
    for (int t = 0; t < SEL_MAX; ++t) tb[t] = t < nb ? bp->bins[t] : -1;
    for (int64_t i0 = (int64_t)blockIdx.x * SEL_BLOCK + (threadIdx.x & ~31); i0 < n; i0 += (int64_t)gridDim.x * SEL_BLOCK) {
        const int64_t i = i0 + lane_id();
        bool take = false;
        double v = 0.0;
        if (i < n && (!mask || mask[i])) {
            v = scaled_value(r, sigma, s1sq, i);
            if (isfinite(v)) {
                const int b = value_bin(v);
                take = (b == tb[0]) | (b == tb[1]) | (b == tb[2]) | (b == tb[3]);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m) {
            unsigned long long base = 0;
            if (lane_id() == 0) base = atomicAdd(n_cand, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (take) cand[base + __popc(m & ((1u << lane_id()) - 1u))] = f64_key(v);
        }
    }
}

struct RdeWork {
    Scratch state, plan, hist, count, out, binhist, binplan, cand, n_cand;
    int64_t cand_cap = 0;
    int init(cudaStream_t st) {
        LSS_CUDA(state.alloc(sizeof(SelectState), st));
        LSS_CUDA(plan.alloc(sizeof(RdePlan), st));
        LSS_CUDA(hist.alloc(sizeof(unsigned long long) * SEL_MAX * SEL_BINS, st));
        LSS_CUDA(count.alloc(sizeof(unsigned long long), st));
        LSS_CUDA(out.alloc(sizeof(double) * 2, st));
        LSS_CUDA(binhist.alloc(sizeof(unsigned long long) * BIN_COUNT, st));
        LSS_CUDA(binplan.alloc(sizeof(BinPlan), st));
        LSS_CUDA(n_cand.alloc(sizeof(unsigned long long), st));
        return LSS_OK;
    }
};

constexpr int64_t kCandidateMin = 1 << 16;    // below this many points the plain 8-pass select is as fast

// One RDE evaluation; result (rde, n) copied to the host and the stream synchronised.
static int rde_eval(RdeWork &wk, const double *r, const double *sigma, const unsigned char *mask, double s1sq, int64_t n,
                    lss_comm *comm, double *rde_h, int64_t *n_h, cudaStream_t st) {
    LSS_CUDA(cudaMemsetAsync(wk.hist.p, 0, sizeof(unsigned long long) * SEL_MAX * SEL_BINS, st));
    LSS_CUDA(cudaMemsetAsync(wk.count.p, 0, sizeof(unsigned long long), st));
    LSS_CUDA(cudaMemsetAsync(wk.state.p, 0, sizeof(SelectState), st));
    const int grid = stride_grid(n, SEL_BLOCK, 8);
    static int two_pass = -1;
    if (two_pass < 0) two_pass = getenv("LSS_SELECT_FULL") ? 0 : 1;
    // with several ranks the choice must not depend on the rank's own point count (the collectives have to match)
    if (two_pass && (n >= kCandidateMin || (comm && comm->n_ranks > 1))) {
        if (wk.cand_cap < n) {
            if (wk.cand.p) { cudaFreeAsync(wk.cand.p, st); wk.cand.p = nullptr; }
            LSS_CUDA(wk.cand.alloc(sizeof(unsigned long long) * (size_t)n, st));
            wk.cand_cap = n;
        }
        LSS_CUDA(cudaMemsetAsync(wk.binhist.p, 0, sizeof(unsigned long long) * BIN_COUNT, st));
        select_binhist_kernel<<<grid, SEL_BLOCK, 0, st>>>(r, sigma, mask, s1sq, n, wk.binhist.as<unsigned long long>(),
                                                          wk.count.as<unsigned long long>());
        LSS_LAUNCH_CHECK();
        LSS_TRY(comm_allreduce(comm, wk.binhist.p, BIN_COUNT, 2, 0, st));
        LSS_TRY(comm_allreduce(comm, wk.count.p, 1, 2, 0, st));
        select_plan_kernel<<<1, 32, 0, st>>>(wk.count.as<unsigned long long>(), wk.state.as<SelectState>(), wk.plan.as<RdePlan>());
        LSS_LAUNCH_CHECK();
        select_binplan_kernel<<<1, SEL_BLOCK, 0, st>>>(wk.state.as<SelectState>(), wk.binhist.as<unsigned long long>(), wk.binplan.as<BinPlan>(),
                                                       wk.n_cand.as<unsigned long long>(), wk.hist.as<unsigned long long>());
        LSS_LAUNCH_CHECK();
        select_compact_kernel<<<grid, SEL_BLOCK, 0, st>>>(r, sigma, mask, s1sq, n, wk.binplan.as<BinPlan>(), wk.cand.as<unsigned long long>(),
                                                          wk.n_cand.as<unsigned long long>());
        LSS_LAUNCH_CHECK();
        const int cgrid = 2 * kSMs;          // the candidate count stays on the device
        for (int pass = 0; pass < 8; ++pass) {
            select_hist_kernel<<<cgrid, SEL_BLOCK, 0, st>>>(nullptr, nullptr, nullptr, 0.0, 0, pass, wk.state.as<SelectState>(),
                                                            wk.hist.as<unsigned long long>(), wk.count.as<unsigned long long>(),
                                                            wk.cand.as<unsigned long long>(), wk.n_cand.as<unsigned long long>());
            LSS_LAUNCH_CHECK();
            LSS_TRY(comm_allreduce(comm, wk.hist.p, SEL_MAX * SEL_BINS, 2, 0, st));
            select_step_kernel<<<1, SEL_BINS, 0, st>>>(wk.state.as<SelectState>(), wk.hist.as<unsigned long long>());
            LSS_LAUNCH_CHECK();
        }
    } else {
        for (int pass = 0; pass < 8; ++pass) {
            if (n > 0) {
                select_hist_kernel<<<grid, SEL_BLOCK, 0, st>>>(r, sigma, mask, s1sq, n, pass, wk.state.as<SelectState>(),
                                                               wk.hist.as<unsigned long long>(), wk.count.as<unsigned long long>());
                LSS_LAUNCH_CHECK();
            }
            LSS_TRY(comm_allreduce(comm, wk.hist.p, SEL_MAX * SEL_BINS, 2, 0, st));
            if (pass == 0) {
                LSS_TRY(comm_allreduce(comm, wk.count.p, 1, 2, 0, st));
                select_plan_kernel<<<1, 32, 0, st>>>(wk.count.as<unsigned long long>(), wk.state.as<SelectState>(), wk.plan.as<RdePlan>());
                LSS_LAUNCH_CHECK();
            }
            select_step_kernel<<<1, SEL_BINS, 0, st>>>(wk.state.as<SelectState>(), wk.hist.as<unsigned long long>());
            LSS_LAUNCH_CHECK();
        }
    }
    select_finish_kernel<<<1, 32, 0, st>>>(wk.state.as<SelectState>(), wk.plan.as<RdePlan>(), wk.out.as<double>());
    LSS_LAUNCH_CHECK();
    double h[2];
    LSS_CUDA(cudaMemcpyAsync(h, wk.out.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    *rde_h = h[0];
    if (n_h) *n_h = (int64_t)h[1];
    return LSS_OK;
}

// min / max of sigma over the selected points with a finite residual (out[0] = min, out[1] = max; initialised to +-inf)
__global__ void __launch_bounds__(SEL_BLOCK)
sigma_range_kernel(const double *__restrict__ r, const double *__restrict__ sigma, const unsigned char *__restrict__ mask, int64_t n,
                   double *__restrict__ out) {
    double lo = INFINITY, hi = -INFINITY;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * SEL_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * SEL_BLOCK) {
        if (mask && !mask[i]) continue;
        if (!isfinite(r[i])) continue;
        const double sg = sigma[i];
        if (!(sg == sg)) bad = true;
        lo = fmin(lo, sg);
        hi = fmax(hi, sg);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (__any_sync(0xffffffffu, bad)) { lo = -INFINITY; hi = INFINITY; }      // a NaN sigma: never "constant"
    if (lane_id() == 0) {
        // fp64 min / max through the ordered integer image of the value
        atomicMin(reinterpret_cast<unsigned long long *>(out), f64_key(lo));
        atomicMax(reinterpret_cast<unsigned long long *>(out) + 1, f64_key(hi));
    }
}

static double host_key_f64(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &u, sizeof(v));
    return v;
}

// libm pow through a volatile exponent: NumPy evaluates scalar `x**2` with pow(), which is not always
// the correctly rounded x*x; the compiler must not fold it.
static double py_square(double x) {
    volatile double two = 2.0;
    return pow(x, two);
}

}  // namespace lss

using namespace lss;

extern "C" int lss_rde(const double *r, const double *sigma, const uint8_t *mask, double s1, int64_t n, lss_comm_t *comm,
                       double *rde_h, int64_t *n_used_h, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!rde_h) return fail(LSS_EINVAL, "rde: null output");
    RdeWork wk;
    LSS_TRY(wk.init(st));
    return rde_eval(wk, r, sigma, mask, py_square(s1), n, comm, rde_h, n_used_h, st);
}

extern "C" int lss_sigma_extra(const double *r, const double *sigma, const uint8_t *mask, int64_t n, lss_comm_t *comm,
                               double *sigma_extra_h, int32_t *nfev_h, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!sigma_extra_h) return fail(LSS_EINVAL, "sigma_extra: null output");
    *sigma_extra_h = 0.0;
    if (nfev_h) *nfev_h = 0;
    RdeWork wk;
    LSS_TRY(wk.init(st));
    // sigma_hat = RDE(this_r); fewer than 10 selected points -> sigma_extra stays 0 (calc_sigma_extra.py:34-38).
    // The count test in the reference is on the mask (finite or not); r is finite on this path.
    double sigma_hat = 0.0;
    int64_t n_used = 0;
    LSS_TRY(rde_eval(wk, r, nullptr, mask, 0.0, n, comm, &sigma_hat, &n_used, st));
    if (n_used < 10) return LSS_OK;
    if (!(sigma_hat == sigma_hat) || isinf(sigma_hat)) return LSS_OK;   // scipy raises -> reference swallows, keeps 0
    int nfev = 0, rc = LSS_OK;
    // Constant sigma over the selected points (one value per sensor / data set is common): x = r / sqrt(s1^2 + sigma^2) is then
    // a monotone image of r for every s1, so its order statistics are the SAME elements the evaluation above selected from r.
    // Every further RDE costs four divisions on the host -- with the operations and roundings of the device code -- instead
    // of passes over the points.  (One range pass over sigma decides; any spread, NaN or non-positive value keeps the general path.)
    bool const_sigma = false;
    double sg_const = 0.0, rv[SEL_MAX] = {0, 0, 0, 0};
    RdePlan plan_h;
    memset(&plan_h, 0, sizeof(plan_h));
    if (!getenv("LSS_SIGMA_EXTRA_GENERAL")) {
        Scratch range;
        LSS_CUDA(range.alloc(sizeof(unsigned long long) * 2, st));
        const unsigned long long init[2] = {~0ull, 0ull};
        LSS_CUDA(cudaMemcpyAsync(range.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        sigma_range_kernel<<<stride_grid(n, SEL_BLOCK, 8), SEL_BLOCK, 0, st>>>(r, sigma, mask, n, range.as<double>());
        LSS_LAUNCH_CHECK();
        unsigned long long keys[2];
        SelectState st_h;
        LSS_CUDA(cudaMemcpyAsync(keys, range.p, sizeof(keys), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaMemcpyAsync(&st_h, wk.state.p, sizeof(st_h), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaMemcpyAsync(&plan_h, wk.plan.p, sizeof(plan_h), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
        double lo = host_key_f64(keys[0]), hi = host_key_f64(keys[1]);
        if (comm && comm->n_ranks > 1) {           // the same decision on every rank
            Scratch mm;
            LSS_CUDA(mm.alloc(sizeof(double) * 2, st));
            const double v[2] = {lo, -hi};
            LSS_CUDA(cudaMemcpyAsync(mm.p, v, sizeof(v), cudaMemcpyHostToDevice, st));
            LSS_TRY(comm_allreduce(comm, mm.p, 2, 0, 1, st));
            double w[2];
            LSS_CUDA(cudaMemcpyAsync(w, mm.p, sizeof(w), cudaMemcpyDeviceToHost, st));
            LSS_CUDA(cudaStreamSynchronize(st));
            lo = w[0]; hi = -w[1];
        }
        if (lo == hi && lo > 0.0 && std::isfinite(lo) && plan_h.n >= 2) {
            const_sigma = true;
            sg_const = lo;
            for (int t = 0; t < SEL_MAX && t < plan_h.n_targets; ++t) rv[t] = host_key_f64(st_h.prefix[t]);
        }
    }
    auto rde_const = [&](double s1sq) -> double {
        // scaled_value() and select_finish_kernel on the host: every operation rounded on its own (volatile: no contraction)
        volatile double sq = sg_const * sg_const;
        volatile double sum = s1sq + sq;
        volatile double c = sqrt(sum);
        double q[2];
        for (int k = 0; k < 2; ++k) {
            volatile double lo = rv[plan_h.lo_idx[k]] / c;
            if (plan_h.clamp[k]) { q[k] = lo; continue; }
            volatile double hi = rv[plan_h.hi_idx[k]] / c;
            volatile double d = hi - lo;
            volatile double slope = d / 1.0;
            volatile double prod = slope * plan_h.frac[k];
            volatile double qq = prod + lo;
            q[k] = qq;
        }
        volatile double diff = q[1] - q[0];
        return diff / 2.0;
    };
    auto f = [&](double s1) -> double {
        double v = 0.0;
        if (const_sigma) {
            v = rde_const(py_square(s1));
        } else {
            int e = rde_eval(wk, r, sigma, mask, py_square(s1), n, comm, &v, nullptr, st);
            if (e != LSS_OK) rc = e;
        }
        ++nfev;
        return py_square(v - 1.0);
    };
    // Bounded Brent (Forsythe-Malcolm-Moler FMIN / MATLAB fminbnd) with scipy's constants
    // (scipy/optimize/_optimize.py `_minimize_scalar_bounded`: xatol=1e-5, maxiter=500).
    const double xatol = 1e-5;
    const int maxfun = 500;
    const double seps = sqrt(2.2e-16), gold = 0.5 * (3.0 - sqrt(5.0));
    double a = 0.0, b = sigma_hat;
    double xf = a + gold * (b - a), w2 = xf, v3 = xf;   // best, second best (nfc), third best (fulc)
    double step = 0.0, prev = 0.0;
    double fx = f(xf), fw = fx, fv = fx;
    double mid = 0.5 * (a + b);
    double t1 = seps * fabs(xf) + xatol / 3.0, t2 = 2.0 * t1;
    while (rc == LSS_OK && fabs(xf - mid) > (t2 - 0.5 * (b - a))) {
        bool golden = true;
        if (fabs(prev) > t1) {
            golden = false;
            double rr = (xf - w2) * (fx - fv);
            double q = (xf - v3) * (fx - fw);
            double p = (xf - v3) * q - (xf - w2) * rr;
            q = 2.0 * (q - rr);
            if (q > 0.0) p = -p;
            q = fabs(q);
            rr = prev;
            prev = step;
            if (fabs(p) < fabs(0.5 * q * rr) && p > q * (a - xf) && p < q * (b - xf)) {
                step = (p + 0.0) / q;
                const double x = xf + step;
                if ((x - a) < t2 || (b - x) < t2) {
                    const double d = mid - xf;
                    const double sgn = (d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0)) + (d == 0.0 ? 1.0 : 0.0);
                    step = t1 * sgn;
                }
            } else {
                golden = true;
            }
        }
        if (golden) {
            prev = (xf >= mid) ? a - xf : b - xf;
            step = gold * prev;
        }
        const double sgn = (step > 0.0 ? 1.0 : (step < 0.0 ? -1.0 : 0.0)) + (step == 0.0 ? 1.0 : 0.0);
        const double x = xf + sgn * fmax(fabs(step), t1);
        const double fu = f(x);
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            v3 = w2; fv = fw;
            w2 = xf; fw = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if (fu <= fw || w2 == xf) {
                v3 = w2; fv = fw;
                w2 = x; fw = fu;
            } else if (fu <= fv || v3 == xf || v3 == w2) {
                v3 = x; fv = fu;
            }
        }
        mid = 0.5 * (a + b);
        t1 = seps * fabs(xf) + xatol / 3.0;
        t2 = 2.0 * t1;
        if (nfev >= maxfun) break;
    }
    if (rc != LSS_OK) return rc;
    *sigma_extra_h = xf;
    if (nfev_h) *nfev_h = nfev;
    return LSS_OK;
}

extern "C" int lss_edit_mask(const double *r, const double *sigma, const double *sigma_extra, int64_t n, double n_sigma,
                             uint8_t *in_TSE, double *stats_h, lss_comm_t *comm, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Scratch acc;
    LSS_CUDA(acc.alloc(3 * sizeof(unsigned long long), st));
    unsigned long long init[3] = {0, 0, ~0ull};
    LSS_CUDA(cudaMemcpyAsync(acc.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (n > 0) {
        edit_mask_kernel<<<stride_grid(n, 256), 256, 0, st>>>(r, sigma, sigma_extra, n, n_sigma, in_TSE, acc.as<unsigned long long>(),
                                                             acc.as<unsigned long long>() + 2);
        LSS_LAUNCH_CHECK();
    }
    LSS_TRY(comm_allreduce(comm, acc.p, 2, 2, 0, st));
    LSS_TRY(comm_allreduce(comm, acc.as<unsigned long long>() + 2, 1, 2, 1, st));
    unsigned long long h[3];
    LSS_CUDA(cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    if (stats_h) {
        stats_h[0] = (double)h[0];
        stats_h[2] = (double)h[1];
        if (h[2] == ~0ull) stats_h[1] = INFINITY;
        else {
            unsigned long long u = (h[2] >> 63) ? (h[2] & 0x7fffffffffffffffull) : ~h[2];
            double d;
            memcpy(&d, &u, sizeof(d));
            stats_h[1] = d;
        }
    }
    return LSS_OK;
}

extern "C" int lss_data_weights(const double *sigma, const double *sigma_extra, const uint8_t *mask, int64_t n, double *w,
                                void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return LSS_OK;
    data_weights_kernel<<<stride_grid(n, 256), 256, 0, st>>>(sigma, sigma_extra, mask, n, w);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}
