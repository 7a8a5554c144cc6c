// Device-wide exclusive scan, stable LSD radix sort, COO -> canonical CSR and CSR transpose.
//
// Replaces scipy's coo_tocsr + csr_sum_duplicates reached from lin_op.toCSR (LSsurf/lin_op.py:745-753)
// and csr.T / tocsc used for G^T products (LSsurf/smooth_fit.py:142, 339-345).  All of it is
// HBM-bound integer work: coalesced tile loads, warp-aggregated shared-memory histograms
// (__match_any_sync), stable warp-level multi-split ranking, no global atomics on the data path.
#include <stdlib.h>
#include <math.h>

#include "common.cuh"

namespace lss {

// ============================================================================ exclusive scan
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *smem /*[32]*/) {
    // inclusive warp scan
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane_id() >= (unsigned)o) inc += t;
    }
    const int w = threadIdx.x >> 5;
    if (lane_id() == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane_id() < (blockDim.x >> 5) ? smem[lane_id()] : 0;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane_id() >= (unsigned)o) si += t;
        }
        smem[lane_id()] = si - s;  // exclusive warp offsets
        if (lane_id() == 31) smem[32] = si;
    }
    __syncthreads();
    uint32_t r = smem[w] + inc - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

// pass 1: per-tile sums
__global__ void __launch_bounds__(SCAN_BLOCK) scan_reduce_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k * SCAN_BLOCK + threadIdx.x;
        if (i < n) acc += in[i];
    }
    uint32_t total;
    block_exclusive_scan(acc, &total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// pass 3 (or only pass when one tile): scan inside the tile, add the tile offset
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_apply_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t n, const uint32_t *__restrict__ offsets,
                  uint32_t *__restrict__ total_out) {
    __shared__ uint32_t sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;  // blocked: thread owns 8 consecutive
    uint32_t v[SCAN_ITEMS];
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        acc += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(acc, &total, sm) + (offsets ? offsets[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_BLOCK - 1) *total_out = ex;
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, cudaStream_t stream) {
    if (n <= 0) {
        if (total) LSS_CUDA(cudaMemsetAsync(total, 0, sizeof(uint32_t), stream));
        return LSS_OK;
    }
    int tiles = ceil_div(n, SCAN_TILE);
    if (tiles == 1) {
        scan_apply_kernel<<<1, SCAN_BLOCK, 0, stream>>>(in, out, n, nullptr, total);
        LSS_LAUNCH_CHECK();
        return LSS_OK;
    }
    Scratch sums;
    LSS_CUDA(sums.alloc(sizeof(uint32_t) * tiles, stream));
    scan_reduce_kernel<<<tiles, SCAN_BLOCK, 0, stream>>>(in, n, sums.as<uint32_t>());
    LSS_LAUNCH_CHECK();
    LSS_TRY(exclusive_scan_u32(sums.as<uint32_t>(), sums.as<uint32_t>(), tiles, nullptr, stream));
    scan_apply_kernel<<<tiles, SCAN_BLOCK, 0, stream>>>(in, out, n, sums.as<uint32_t>(), total);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

// ============================================================================ radix sort
constexpr int RS_BLOCK = 256;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_ITEMS = 16;                     // keys per thread
constexpr int RS_TILE = RS_BLOCK * RS_ITEMS;     // 4096 keys per CTA
constexpr int RS_BINS = 256;

// Tile-local digit histogram -> counts[digit * n_tiles + tile].
__global__ void __launch_bounds__(RS_BLOCK)
rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t mask, uint32_t *__restrict__ counts, int n_tiles) {
    __shared__ uint32_t h[RS_BINS];
    for (int i = threadIdx.x; i < RS_BINS; i += RS_BLOCK) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + k * RS_BLOCK + threadIdx.x;
        bool in = i < n;
        uint32_t d = in ? (uint32_t)(keys[i] >> shift) & mask : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (in && lane_id() == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RS_BINS; i += RS_BLOCK) counts[(int64_t)i * n_tiles + blockIdx.x] = h[i];
}

// Stable scatter.  Arrangement: warp w owns the contiguous chunk [w*512, (w+1)*512) of the tile and
// reads it in 16 coalesced rounds of 32 keys, so tile order == (warp, round, lane) order.
__global__ void __launch_bounds__(RS_BLOCK)
rs_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
                  uint32_t *__restrict__ vals_out, int64_t n, int shift, uint32_t mask, const uint32_t *__restrict__ offsets,
                  int n_tiles, int iota_vals) {
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS];   // per-warp running digit counts -> exclusive warp offsets
    __shared__ uint32_t gbase[RS_BINS];
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_BLOCK) (&cnt[0][0])[i] = 0;
    for (int i = threadIdx.x; i < RS_BINS; i += RS_BLOCK) gbase[i] = offsets[(int64_t)i * n_tiles + blockIdx.x];
    __syncthreads();
    const int w = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + k * 32 + lane_id();
        bool in = i < n;
        key[k] = in ? keys_in[i] : 0;
        uint32_t d = in ? (uint32_t)(key[k] >> shift) & mask : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (in && lane_id() == leader) {
            old = cnt[w][d];
            cnt[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & ((1u << lane_id()) - 1u));
        __syncwarp();  // the next round's leader may be another lane: order the shared-memory counters
    }
    __syncthreads();
    // exclusive scan over warps, per digit (256 threads == 256 digits)
    {
        uint32_t run = 0;
        const int d = threadIdx.x;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            uint32_t c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + k * 32 + lane_id();
        if (i < n) {
            uint32_t d = (uint32_t)(key[k] >> shift) & mask;
            uint32_t pos = gbase[d] + cnt[w][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = iota_vals ? (uint32_t)i : vals_in[i];
        }
    }
}

int radix_sort_pairs(uint64_t *keys_a, uint32_t *vals_a, uint64_t *keys_b, uint32_t *vals_b, int64_t n, int bit_lo,
                     int n_bits, int iota_first, uint64_t **keys_out, uint32_t **vals_out, cudaStream_t stream) {
    *keys_out = keys_a;
    *vals_out = vals_a;
    if (n <= 0 || n_bits <= 0) return LSS_OK;
    if (n >= (1ll << 32)) return fail(LSS_ERANGE, "radix sort: more than 2^32 items");
    const int n_tiles = ceil_div(n, RS_TILE);
    Scratch counts;
    LSS_CUDA(counts.alloc(sizeof(uint32_t) * (size_t)RS_BINS * n_tiles, stream));
    const int passes = (n_bits + 7) / 8;
    int done = 0;
    uint64_t *ka = keys_a, *kb = keys_b;
    uint32_t *va = vals_a, *vb = vals_b;
    for (int p = 0; p < passes; ++p) {
        int bits = (n_bits - done) / (passes - p);  // spread the bits evenly over the passes
        if ((n_bits - done) % (passes - p)) ++bits;
        const uint32_t mask = (1u << bits) - 1u;
        rs_hist_kernel<<<n_tiles, RS_BLOCK, 0, stream>>>(ka, n, bit_lo + done, mask, counts.as<uint32_t>(), n_tiles);
        LSS_LAUNCH_CHECK();
        LSS_TRY(exclusive_scan_u32(counts.as<uint32_t>(), counts.as<uint32_t>(), (int64_t)RS_BINS * n_tiles, nullptr, stream));
        rs_scatter_kernel<<<n_tiles, RS_BLOCK, 0, stream>>>(ka, va, kb, vb, n, bit_lo + done, mask, counts.as<uint32_t>(),
                                                            n_tiles, iota_first && p == 0);
        LSS_LAUNCH_CHECK();
        done += bits;
        uint64_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    *keys_out = ka;
    *vals_out = va;
    return LSS_OK;
}

// ============================================================================ COO -> CSR
static int bits_for(int64_t max_value) {  // bits needed to represent values in [0, max_value]
    int b = 1;
    while (b < 63 && (max_value >> b)) ++b;
    return b;
}

// key = row << 32 | col; entries with v == 0 get the sentinel row n_rows so they sort last.
// stats[0] = number kept, stats[1] = max kept row, stats[2] = max kept col (as unsigned).
__global__ void __launch_bounds__(256)
coo_keys_kernel(int64_t nnz, const int *__restrict__ rows, const int *__restrict__ cols, const double *__restrict__ vals,
                int64_t n_rows, uint64_t *__restrict__ keys, unsigned long long *__restrict__ stats) {
    unsigned long long good = 0, mr = 0, mc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const bool keep = vals[i] != 0.0;   // NaN is kept, -0.0 dropped, as NumPy's v != 0 (lin_op.py:750)
        const uint64_t r = keep ? (uint64_t)(uint32_t)rows[i] : (uint64_t)n_rows;
        const uint64_t c = keep ? (uint64_t)(uint32_t)cols[i] : 0;
        keys[i] = (r << 32) | c;
        if (keep) {
            ++good;
            mr = r > mr ? r : mr;
            mc = c > mc ? c : mc;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        good += __shfl_xor_sync(0xffffffffu, good, o);
        unsigned long long a = __shfl_xor_sync(0xffffffffu, mr, o), b = __shfl_xor_sync(0xffffffffu, mc, o);
        mr = a > mr ? a : mr;
        mc = b > mc ? b : mc;
    }
    if (lane_id() == 0 && good) {
        atomicAdd(stats, good);
        atomicMax(stats + 1, mr);
        atomicMax(stats + 2, mc);
    }
}

__global__ void __launch_bounds__(256)
head_flags_kernel(const uint64_t *__restrict__ keys, int64_t n, uint32_t *__restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// One thread per sorted entry; run heads sum their duplicates left to right (scipy csr_sum_duplicates
// order after a stable sort) and write the CSR entry; row starts fill indptr for every row they open.
__global__ void __launch_bounds__(256)
csr_emit_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx, const double *__restrict__ vals,
                int64_t n, const uint32_t *__restrict__ pos, int64_t n_rows, int *__restrict__ indptr,
                int *__restrict__ indices, double *__restrict__ data, const uint32_t *__restrict__ n_out) {
    const uint64_t cmask = 0xffffffffull;
    const int col_bits = 32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        const bool head = (i == 0) || (keys[i - 1] != k);
        if (!head) continue;
        double acc = vals[idx[i]];
        int64_t j = i + 1;
        for (; j < n && keys[j] == k; ++j) acc += vals[idx[j]];
        const uint32_t p = pos[i];
        indices[p] = (int)(k & cmask);
        data[p] = acc;
        const int64_t row = (int64_t)(k >> col_bits);
        const int64_t prev_row = (i == 0) ? -1 : (int64_t)(keys[i - 1] >> col_bits);
        for (int64_t r = prev_row + 1; r <= row; ++r) indptr[r] = (int)p;
        if (j == n) {  // this run is the last one: close every remaining row
            const int total = (int)(*n_out);
            for (int64_t r = row + 1; r <= n_rows; ++r) indptr[r] = total;
        }
    }
}

__global__ void fill_i32_kernel(int *p, int64_t n, int v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ============================================================================ transpose
__global__ void __launch_bounds__(256)
csr_expand_keys_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices,
                       uint64_t *__restrict__ keys, uint32_t *__restrict__ src_row) {
    // one warp per row: keys[k] = column of entry k, payload = source row
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const int b = indptr[r], e = indptr[r + 1];
        for (int k = b + (int)lane_id(); k < e; k += 32) {
            keys[k] = (uint64_t)(uint32_t)indices[k];
            src_row[k] = (uint32_t)r;
        }
    }
}

// After sorting entry ids by column: t_indices = source row, t_data = value, t_indptr from key changes.
__global__ void __launch_bounds__(256)
transpose_emit_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ent, const uint32_t *__restrict__ src_row,
                      const double *__restrict__ data, int64_t nnz, int64_t n_cols, int *__restrict__ t_indptr,
                      int *__restrict__ t_indices, double *__restrict__ t_data) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t e = ent[i];
        t_indices[i] = (int)src_row[e];
        t_data[i] = data[e];
        const int64_t c = (int64_t)keys[i];
        const int64_t pc = (i == 0) ? -1 : (int64_t)keys[i - 1];
        for (int64_t r = pc + 1; r <= c; ++r) t_indptr[r] = (int)i;
        if (i == nnz - 1)
            for (int64_t r = c + 1; r <= n_cols; ++r) t_indptr[r] = (int)nnz;
    }
}

__global__ void __launch_bounds__(256) copy_order_kernel(const uint32_t *__restrict__ in, int *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int)in[i];
}


// ============================================================================ dense ranks / group medians
__global__ void __launch_bounds__(256)
rank_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ pos,
                    int64_t n, int *__restrict__ rank, uint64_t *__restrict__ uniq) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool head = (i == 0) || (keys[i] != keys[i - 1]);
        // pos = exclusive scan of the head flags: rank of a non-head is pos - 1
        const uint32_t r = head ? pos[i] : pos[i] - 1;
        rank[idx[i]] = (int)r;
        if (head && uniq) uniq[r] = keys[i];
    }
}

__device__ __forceinline__ unsigned long long f64_sort_key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256)
f64_keys_kernel(const double *__restrict__ x, int64_t n, double nan_value, uint64_t *__restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v != v) v = nan_value;
        if (v == 0.0) v = 0.0;               // -0.0 and +0.0 compare equal in NumPy
        // a NaN that is still a NaN sorts last whatever its sign bit (NumPy's order); its payload is not kept
        keys[i] = (v != v) ? 0xfff8000000000000ull : f64_sort_key(v);
    }
}

// keys (id << 0) for the second, stable pass of the (id, value) sort
__global__ void __launch_bounds__(256)
gather_id_keys_kernel(const int *__restrict__ ids, const uint32_t *__restrict__ order, int64_t n, uint64_t *__restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = (uint64_t)(uint32_t)ids[order[i]];
}

// one thread per group: [lo, hi) = its segment in the (id, value)-sorted order; NaNs sort last inside a group
__global__ void __launch_bounds__(128)
group_median_kernel(const uint64_t *__restrict__ id_keys, const uint32_t *__restrict__ order, const double *__restrict__ vals,
                    int64_t n, int n_groups, double *__restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int64_t lo = 0, hi = n;                       // lower_bound(g), lower_bound(g + 1)
    {
        int64_t a = 0, b = n;
        while (a < b) { int64_t m = (a + b) >> 1; if (id_keys[m] < (uint64_t)g) a = m + 1; else b = m; }
        lo = a;
        b = n;
        while (a < b) { int64_t m = (a + b) >> 1; if (id_keys[m] < (uint64_t)g + 1) a = m + 1; else b = m; }
        hi = a;
    }
    while (hi > lo && vals[order[hi - 1]] != vals[order[hi - 1]]) --hi;      // drop trailing NaNs (nanmedian)
    const int64_t m = hi - lo;
    if (m <= 0) { out[g] = nan(""); return; }
    const double a = vals[order[lo + (m - 1) / 2]], b = vals[order[lo + m / 2]];
    out[g] = (m & 1) ? a : (a + b) / 2.0;                                   // np.median: mean of the two middle values
}


// ---- fast paths: small-integer dense ranks, groups of equal values, row-regular COO blocks ----------------------

// stats[0] = min, stats[1] = max of the order keys of x (NaN -> nan_value, -0 -> +0); stats[2] != 0 when some value is
// not an integer (or not finite): the counting path below does not apply then.
__global__ void __launch_bounds__(256)
int_range_kernel(const double *__restrict__ x, int64_t n, double nan_value, unsigned long long *__restrict__ stats) {
    unsigned long long mn = ~0ull, mx = 0ull;
    unsigned int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v != v) v = nan_value;
        if (v == 0.0) v = 0.0;
        if (!(v == floor(v)) || !(fabs(v) < 4.0e15)) { bad = 1; continue; }
        const unsigned long long k = f64_sort_key(v);
        mn = k < mn ? k : mn;
        mx = k > mx ? k : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane_id() == 0) {
        if (mn != ~0ull) { atomicMin(stats, mn); atomicMax(stats + 1, mx); }
        if (bad) atomicOr(stats + 2, 1ull);
    }
}

__global__ void __launch_bounds__(256)
int_presence_kernel(const double *__restrict__ x, int64_t n, double nan_value, double lo, uint32_t *__restrict__ present) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v != v) v = nan_value;
        present[(int64_t)(v - lo)] = 1u;
    }
}

__global__ void __launch_bounds__(256)
int_rank_kernel(const double *__restrict__ x, int64_t n, double nan_value, double lo, const uint32_t *__restrict__ pos, int *__restrict__ rank) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v != v) v = nan_value;
        rank[i] = (int)pos[(int64_t)(v - lo)];
    }
}

__global__ void __launch_bounds__(256)
int_unique_kernel(int64_t n_bins, double lo, const uint32_t *__restrict__ present, const uint32_t *__restrict__ pos, double *__restrict__ uniq) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x)
        if (present[b]) uniq[pos[b]] = lo + (double)b;
}

// per-group minimum / maximum order key (warp-aggregated: neighbouring points mostly share their group); flag[0] != 0
// when a NaN was met or an id is out of range
__global__ void __launch_bounds__(256)
group_minmax_kernel(const int *__restrict__ ids, const double *__restrict__ vals, int64_t n, int n_groups,
                    unsigned long long *__restrict__ gmin, unsigned long long *__restrict__ gmax, unsigned int *__restrict__ flag) {
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane_id();
        int g = -1;
        unsigned long long k = 0;
        if (i < n) {
            g = ids[i];
            const double v = vals[i];
            if (v != v || g < 0 || g >= n_groups) { atomicOr(flag, 1u); g = -1; }
            else k = f64_sort_key(v == 0.0 ? 0.0 : v);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, g);
        unsigned long long mn = k, mx = k;
        // reduce over the peer set with shuffles from each peer lane
        for (unsigned m = peers; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const unsigned long long ko = __shfl_sync(peers, k, src);
            mn = ko < mn ? ko : mn;
            mx = ko > mx ? ko : mx;
        }
        if (g >= 0 && lane_id() == (unsigned)(__ffs(peers) - 1)) {
            atomicMin(gmin + g, mn);
            atomicMax(gmax + g, mx);
        }
    }
}

__global__ void __launch_bounds__(256)
group_const_kernel(int n_groups, const unsigned long long *__restrict__ gmin, const unsigned long long *__restrict__ gmax,
                   double *__restrict__ out, unsigned int *__restrict__ flag) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const unsigned long long a = gmin[g], b = gmax[g];
    if (a == ~0ull && b == 0ull) { out[g] = nan(""); return; }          // empty group: nanmedian of nothing
    if (a != b) { atomicOr(flag, 2u); return; }
    const unsigned long long u = (a >> 63) ? (a & 0x7fffffffffffffffull) : ~a;
    out[g] = __longlong_as_double((long long)u);
}

// Row-regular COO blocks -> canonical CSR without sorting.  Block b holds k[b] entries per row, row-major; the rows of
// the result take their entries block after block, zero values dropped.  This is already canonical when the kept
// columns of every row ascend strictly (interpolation corners in np.mgrid order, one grid after the other): rows that
// violate it are counted in *unordered and the caller falls back to the sort.
constexpr int BLK_MAX = 8;
struct CooBlocks {
    int n_blocks;
    int k[BLK_MAX];
    const int *cols[BLK_MAX];
    const double *vals[BLK_MAX];
};

template <bool FILL>
__global__ void __launch_bounds__(256)
blocks_rows_kernel(int64_t n_rows, int64_t n_cols, CooBlocks B, uint32_t *__restrict__ counts, const int *__restrict__ indptr,
                   int *__restrict__ indices, double *__restrict__ data, unsigned long long *__restrict__ unordered) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int cnt = 0, prev = -1;
    bool bad = false;
    int64_t o = FILL ? indptr[r] : 0;
    for (int b = 0; b < B.n_blocks; ++b) {
        const int k = B.k[b];
        const int *c = B.cols[b] + r * k;
        const double *v = B.vals[b] + r * k;
        for (int j = 0; j < k; ++j) {
            const double val = v[j];
            if (val == 0.0) continue;
            const int col = c[j];
            if (FILL) {
                indices[o] = col;
                data[o] = val;
                ++o;
            } else {
                if (col <= prev || col >= n_cols) bad = true;
                prev = col;
                ++cnt;
            }
        }
    }
    if (!FILL) {
        counts[r] = (uint32_t)cnt;
        if (bad) atomicAdd(unordered, 1ull);
    }
}

__global__ void __launch_bounds__(256) set_last_kernel(int *__restrict__ indptr, int64_t n_rows, const uint32_t *__restrict__ total) {
    if (blockIdx.x == 0 && threadIdx.x == 0) indptr[n_rows] = (int)*total;
}

}  // namespace lss

using namespace lss;

extern "C" int lss_blocks_to_csr(int64_t n_rows, int64_t n_cols, int32_t n_blocks, const int32_t *k_h, const int32_t *const *cols_h,
                                 const double *const *vals_h, int32_t *indptr, int32_t *indices, double *data, int64_t *nnz_out_h,
                                 int64_t *unordered_h, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_blocks < 1 || n_blocks > BLK_MAX || n_rows < 0 || !k_h || !cols_h || !vals_h || !nnz_out_h || !unordered_h)
        return fail(LSS_EINVAL, "blocks_to_csr: bad arguments (1..%d blocks)", BLK_MAX);
    CooBlocks B;
    memset(&B, 0, sizeof(B));
    B.n_blocks = n_blocks;
    int64_t per_row = 0;
    for (int b = 0; b < n_blocks; ++b) {
        if (k_h[b] < 1 || k_h[b] > 64) return fail(LSS_EINVAL, "blocks_to_csr: %d entries per row in block %d", k_h[b], b);
        B.k[b] = k_h[b]; B.cols[b] = cols_h[b]; B.vals[b] = vals_h[b];
        per_row += k_h[b];
    }
    if (per_row * n_rows >= (1ll << 31)) return fail(LSS_ERANGE, "blocks_to_csr: exceeds int32 index range");
    *nnz_out_h = 0;
    *unordered_h = 0;
    if (n_rows == 0) {
        LSS_CUDA(cudaMemsetAsync(indptr, 0, sizeof(int), st));
        LSS_CUDA(cudaStreamSynchronize(st));
        return LSS_OK;
    }
    Scratch counts, acc;
    LSS_CUDA(counts.alloc(sizeof(uint32_t) * n_rows, st));
    LSS_CUDA(acc.alloc(2 * sizeof(unsigned long long), st));
    LSS_CUDA(cudaMemsetAsync(acc.p, 0, 2 * sizeof(unsigned long long), st));
    const int grid = ceil_div(n_rows, 256);
    blocks_rows_kernel<false><<<grid, 256, 0, st>>>(n_rows, n_cols, B, counts.as<uint32_t>(), nullptr, nullptr, nullptr, acc.as<unsigned long long>());
    LSS_LAUNCH_CHECK();
    uint32_t *total = reinterpret_cast<uint32_t *>(acc.as<unsigned long long>() + 1);
    LSS_TRY(exclusive_scan_u32(counts.as<uint32_t>(), reinterpret_cast<uint32_t *>(indptr), n_rows, total, st));
    set_last_kernel<<<1, 32, 0, st>>>(indptr, n_rows, total);
    LSS_LAUNCH_CHECK();
    unsigned long long h[2] = {0, 0};
    LSS_CUDA(cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    *unordered_h = (int64_t)h[0];
    if (h[0]) return LSS_OK;                  // not canonical by construction: the caller sorts
    blocks_rows_kernel<true><<<grid, 256, 0, st>>>(n_rows, n_cols, B, nullptr, indptr, indices, data, nullptr);
    LSS_LAUNCH_CHECK();
    *nnz_out_h = (int64_t)(uint32_t)h[1];
    return LSS_OK;
}

extern "C" int lss_coo_to_csr(int64_t nnz, const int32_t *rows, const int32_t *cols, const double *vals, int64_t n_rows,
                              int64_t n_cols, int32_t *indptr, int32_t *indices, double *data, int64_t *nnz_out_h,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz < 0 || n_rows < 0 || n_cols < 0) return fail(LSS_EINVAL, "coo_to_csr: negative size");
    if (nnz >= (1ll << 31) || n_rows >= (1ll << 31) - 1 || n_cols >= (1ll << 31)) return fail(LSS_ERANGE, "coo_to_csr: exceeds int32 index range");
    if (nnz_out_h) *nnz_out_h = 0;
    unsigned long long stats[3] = {0, 0, 0};
    Scratch ka, kb, va, vb, cnt, flags;
    if (nnz > 0) {
        LSS_CUDA(ka.alloc(sizeof(uint64_t) * nnz, st));
        LSS_CUDA(cnt.alloc(4 * sizeof(unsigned long long), st));
        LSS_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * sizeof(unsigned long long), st));
        coo_keys_kernel<<<stride_grid(nnz, 256), 256, 0, st>>>(nnz, rows, cols, vals, n_rows, ka.as<uint64_t>(), cnt.as<unsigned long long>());
        LSS_LAUNCH_CHECK();
        LSS_CUDA(cudaMemcpyAsync(stats, cnt.p, sizeof(stats), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
    }
    const int64_t n = (int64_t)stats[0];
    if (n == 0) {
        fill_i32_kernel<<<stride_grid(n_rows + 1, 256), 256, 0, st>>>(indptr, n_rows + 1, 0);
        LSS_LAUNCH_CHECK();
        LSS_CUDA(cudaStreamSynchronize(st));
        return LSS_OK;
    }
    // scipy raises for indices outside the matrix (coo -> csr validation)
    if ((int64_t)stats[1] >= n_rows) return fail(LSS_EINVAL, "coo_to_csr: row index %llu exceeds matrix dimension %lld", stats[1], (long long)n_rows);
    if ((int64_t)stats[2] >= n_cols) return fail(LSS_EINVAL, "coo_to_csr: column index %llu exceeds matrix dimension %lld", stats[2], (long long)n_cols);
    LSS_CUDA(kb.alloc(sizeof(uint64_t) * nnz, st));
    LSS_CUDA(va.alloc(sizeof(uint32_t) * nnz, st));
    LSS_CUDA(vb.alloc(sizeof(uint32_t) * nnz, st));
    // LSD: column bits first (payload = original position, written by the first pass), then row bits
    // (the sentinel row n_rows of dropped entries is included so they sort to the tail).
    uint64_t *k1, *ks;
    uint32_t *v1, *vs;
    const int cbits = bits_for((int64_t)stats[2]);
    const int rbits = bits_for(n_rows);
    LSS_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), nnz, 0, cbits, 1, &k1, &v1, st));
    uint64_t *k2 = (k1 == ka.as<uint64_t>()) ? kb.as<uint64_t>() : ka.as<uint64_t>();
    uint32_t *v2 = (v1 == va.as<uint32_t>()) ? vb.as<uint32_t>() : va.as<uint32_t>();
    LSS_TRY(radix_sort_pairs(k1, v1, k2, v2, nnz, 32, rbits, 0, &ks, &vs, st));
    uint32_t *n_out_dev = reinterpret_cast<uint32_t *>(cnt.as<unsigned long long>() + 3);
    LSS_CUDA(flags.alloc(sizeof(uint32_t) * n, st));
    head_flags_kernel<<<stride_grid(n, 256), 256, 0, st>>>(ks, n, flags.as<uint32_t>());
    LSS_LAUNCH_CHECK();
    LSS_TRY(exclusive_scan_u32(flags.as<uint32_t>(), flags.as<uint32_t>(), n, n_out_dev, st));
    csr_emit_kernel<<<stride_grid(n, 256, 16), 256, 0, st>>>(ks, vs, vals, n, flags.as<uint32_t>(), n_rows, indptr, indices, data, n_out_dev);
    LSS_LAUNCH_CHECK();
    uint32_t n_out = 0;
    LSS_CUDA(cudaMemcpyAsync(&n_out, n_out_dev, sizeof(n_out), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    if (nnz_out_h) *nnz_out_h = (int64_t)n_out;
    return LSS_OK;
}

extern "C" int lss_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *indptr, const int32_t *indices,
                                 const double *data, int32_t *t_indptr, int32_t *t_indices, double *t_data, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz >= (1ll << 31) || n_rows >= (1ll << 31) || n_cols >= (1ll << 31) - 1) return fail(LSS_ERANGE, "csr_transpose: exceeds int32 index range");
    if (nnz == 0) {
        fill_i32_kernel<<<stride_grid(n_cols + 1, 256), 256, 0, st>>>(t_indptr, n_cols + 1, 0);
        LSS_LAUNCH_CHECK();
        return LSS_OK;
    }
    Scratch ka, kb, va, vb, src;
    LSS_CUDA(ka.alloc(sizeof(uint64_t) * nnz, st));
    LSS_CUDA(kb.alloc(sizeof(uint64_t) * nnz, st));
    LSS_CUDA(va.alloc(sizeof(uint32_t) * nnz, st));
    LSS_CUDA(vb.alloc(sizeof(uint32_t) * nnz, st));
    LSS_CUDA(src.alloc(sizeof(uint32_t) * nnz, st));
    csr_expand_keys_kernel<<<stride_grid(n_rows * 32, 256, 16), 256, 0, st>>>(n_rows, indptr, indices, ka.as<uint64_t>(), src.as<uint32_t>());
    LSS_LAUNCH_CHECK();
    uint64_t *ks;
    uint32_t *vs;
    LSS_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), nnz, 0,
                             bits_for(n_cols > 0 ? n_cols - 1 : 0), 1, &ks, &vs, st));
    transpose_emit_kernel<<<stride_grid(nnz, 256, 16), 256, 0, st>>>(ks, vs, src.as<uint32_t>(), data, nnz, n_cols, t_indptr, t_indices, t_data);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

extern "C" int lss_argsort_u64(int64_t n, const uint64_t *keys, int32_t key_bits, int32_t *order, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n >= (1ll << 31)) return fail(LSS_ERANGE, "argsort: exceeds int32 index range");
    if (key_bits < 1 || key_bits > 64) return fail(LSS_EINVAL, "argsort: key_bits must be 1..64");
    if (n == 0) return LSS_OK;
    Scratch ka, kb, va, vb;
    LSS_CUDA(ka.alloc(sizeof(uint64_t) * n, st));
    LSS_CUDA(kb.alloc(sizeof(uint64_t) * n, st));
    LSS_CUDA(va.alloc(sizeof(uint32_t) * n, st));
    LSS_CUDA(vb.alloc(sizeof(uint32_t) * n, st));
    LSS_CUDA(cudaMemcpyAsync(ka.p, keys, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, st));
    uint64_t *ks;
    uint32_t *vs;
    LSS_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, 0, key_bits, 1, &ks, &vs, st));
    copy_order_kernel<<<stride_grid(n, 256), 256, 0, st>>>(vs, order, n);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

extern "C" int lss_dense_rank_f64(const double *x, int64_t n, double nan_value, int32_t *rank, double *unique_h,
                                  int64_t unique_cap, int64_t *n_unique_h, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n >= (1ll << 31)) return fail(LSS_ERANGE, "dense_rank: exceeds int32 index range");
    if (n_unique_h) *n_unique_h = 0;
    if (n == 0) return LSS_OK;
    {
        // small-integer values (track / cycle numbers, combined ranks): presence table + scan instead of a 64-bit sort
        Scratch stats;
        LSS_CUDA(stats.alloc(3 * sizeof(unsigned long long), st));
        const unsigned long long init[3] = {~0ull, 0ull, 0ull};
        LSS_CUDA(cudaMemcpyAsync(stats.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        int_range_kernel<<<stride_grid(n, 256), 256, 0, st>>>(x, n, nan_value, stats.as<unsigned long long>());
        LSS_LAUNCH_CHECK();
        unsigned long long h[3];
        LSS_CUDA(cudaMemcpyAsync(h, stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
        auto decode = [](unsigned long long k) { unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k; double d; memcpy(&d, &u, sizeof(d)); return d; };
        if (h[2] == 0 && h[0] != ~0ull && !getenv("LSS_RANK_SORT")) {
            const double lo = decode(h[0]), hi = decode(h[1]);
            const double span = hi - lo;
            if (span >= 0 && span < (double)(1 << 22)) {
                const int64_t n_bins = (int64_t)span + 1;
                Scratch present, pos, tot, uq;
                LSS_CUDA(present.alloc(sizeof(uint32_t) * n_bins, st));
                LSS_CUDA(pos.alloc(sizeof(uint32_t) * n_bins, st));
                LSS_CUDA(tot.alloc(sizeof(uint32_t), st));
                LSS_CUDA(uq.alloc(sizeof(double) * n_bins, st));
                LSS_CUDA(cudaMemsetAsync(present.p, 0, sizeof(uint32_t) * n_bins, st));
                int_presence_kernel<<<stride_grid(n, 256), 256, 0, st>>>(x, n, nan_value, lo, present.as<uint32_t>());
                LSS_LAUNCH_CHECK();
                LSS_TRY(exclusive_scan_u32(present.as<uint32_t>(), pos.as<uint32_t>(), n_bins, tot.as<uint32_t>(), st));
                int_rank_kernel<<<stride_grid(n, 256), 256, 0, st>>>(x, n, nan_value, lo, pos.as<uint32_t>(), rank);
                LSS_LAUNCH_CHECK();
                int_unique_kernel<<<stride_grid(n_bins, 256), 256, 0, st>>>(n_bins, lo, present.as<uint32_t>(), pos.as<uint32_t>(), uq.as<double>());
                LSS_LAUNCH_CHECK();
                uint32_t nu = 0;
                LSS_CUDA(cudaMemcpyAsync(&nu, tot.p, sizeof(nu), cudaMemcpyDeviceToHost, st));
                LSS_CUDA(cudaStreamSynchronize(st));
                if (n_unique_h) *n_unique_h = nu;
                if (unique_h) {
                    if ((int64_t)nu > unique_cap) return fail(LSS_EINVAL, "dense_rank: %u unique values exceed the capacity %lld", nu, (long long)unique_cap);
                    LSS_CUDA(cudaMemcpyAsync(unique_h, uq.p, sizeof(double) * nu, cudaMemcpyDeviceToHost, st));
                    LSS_CUDA(cudaStreamSynchronize(st));
                }
                return LSS_OK;
            }
        }
    }
    Scratch ka, kb, va, vb, flags, uniq, tot;
    LSS_CUDA(ka.alloc(sizeof(uint64_t) * n, st));
    LSS_CUDA(kb.alloc(sizeof(uint64_t) * n, st));
    LSS_CUDA(va.alloc(sizeof(uint32_t) * n, st));
    LSS_CUDA(vb.alloc(sizeof(uint32_t) * n, st));
    LSS_CUDA(flags.alloc(sizeof(uint32_t) * n, st));
    LSS_CUDA(uniq.alloc(sizeof(uint64_t) * n, st));
    LSS_CUDA(tot.alloc(sizeof(uint32_t), st));
    f64_keys_kernel<<<stride_grid(n, 256), 256, 0, st>>>(x, n, nan_value, ka.as<uint64_t>());
    LSS_LAUNCH_CHECK();
    uint64_t *ks;
    uint32_t *vs;
    LSS_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, 0, 64, 1, &ks, &vs, st));
    head_flags_kernel<<<stride_grid(n, 256), 256, 0, st>>>(ks, n, flags.as<uint32_t>());
    LSS_LAUNCH_CHECK();
    LSS_TRY(exclusive_scan_u32(flags.as<uint32_t>(), flags.as<uint32_t>(), n, tot.as<uint32_t>(), st));
    rank_scatter_kernel<<<stride_grid(n, 256), 256, 0, st>>>(ks, vs, flags.as<uint32_t>(), n, rank, uniq.as<uint64_t>());
    LSS_LAUNCH_CHECK();
    uint32_t nu = 0;
    LSS_CUDA(cudaMemcpyAsync(&nu, tot.p, sizeof(nu), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    if (n_unique_h) *n_unique_h = nu;
    if (unique_h) {
        if ((int64_t)nu > unique_cap) return fail(LSS_EINVAL, "dense_rank: %u unique values exceed the capacity %lld", nu, (long long)unique_cap);
        LSS_CUDA(cudaMemcpyAsync(unique_h, uniq.p, sizeof(uint64_t) * nu, cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < nu; ++i) {   // decode the order-preserving keys back to doubles
            uint64_t k;
            memcpy(&k, &unique_h[i], sizeof(k));
            uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
            memcpy(&unique_h[i], &u, sizeof(u));
        }
    }
    return LSS_OK;
}

extern "C" int lss_group_median(const int32_t *ids, const double *vals, int64_t n, int32_t n_groups, double *out,
                                void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n >= (1ll << 31)) return fail(LSS_ERANGE, "group_median: exceeds int32 index range");
    if (n_groups <= 0) return LSS_OK;
    if (n > 0 && !getenv("LSS_RANK_SORT")) {
        // every group holds one distinct value (a constant sigma_corr per track, say): its median is that value
        Scratch gmin, gmax, flag;
        LSS_CUDA(gmin.alloc(sizeof(unsigned long long) * n_groups, st));
        LSS_CUDA(gmax.alloc(sizeof(unsigned long long) * n_groups, st));
        LSS_CUDA(flag.alloc(sizeof(unsigned int), st));
        LSS_CUDA(cudaMemsetAsync(gmin.p, 0xff, sizeof(unsigned long long) * n_groups, st));
        LSS_CUDA(cudaMemsetAsync(gmax.p, 0, sizeof(unsigned long long) * n_groups, st));
        LSS_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(unsigned int), st));
        group_minmax_kernel<<<stride_grid(n, 256), 256, 0, st>>>(ids, vals, n, n_groups, gmin.as<unsigned long long>(), gmax.as<unsigned long long>(),
                                                                flag.as<unsigned int>());
        LSS_LAUNCH_CHECK();
        group_const_kernel<<<ceil_div(n_groups, 256), 256, 0, st>>>(n_groups, gmin.as<unsigned long long>(), gmax.as<unsigned long long>(), out,
                                                                   flag.as<unsigned int>());
        LSS_LAUNCH_CHECK();
        unsigned int f = 0;
        LSS_CUDA(cudaMemcpyAsync(&f, flag.p, sizeof(f), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
        if (f == 0) return LSS_OK;
    }
    Scratch ka, kb, va, vb;
    LSS_CUDA(ka.alloc(sizeof(uint64_t) * (n + 1), st));
    LSS_CUDA(kb.alloc(sizeof(uint64_t) * (n + 1), st));
    LSS_CUDA(va.alloc(sizeof(uint32_t) * (n + 1), st));
    LSS_CUDA(vb.alloc(sizeof(uint32_t) * (n + 1), st));
    uint64_t *ks = ka.as<uint64_t>();
    uint32_t *vs = va.as<uint32_t>();
    if (n > 0) {
        // sort by value (NaN last), then stably by group id
        f64_keys_kernel<<<stride_grid(n, 256), 256, 0, st>>>(vals, n, nan(""), ka.as<uint64_t>());
        LSS_LAUNCH_CHECK();
        uint64_t *k1;
        uint32_t *v1;
        LSS_TRY(radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, 0, 64, 1, &k1, &v1, st));
        uint64_t *k2 = (k1 == ka.as<uint64_t>()) ? kb.as<uint64_t>() : ka.as<uint64_t>();
        uint32_t *v2 = (v1 == va.as<uint32_t>()) ? vb.as<uint32_t>() : va.as<uint32_t>();
        gather_id_keys_kernel<<<stride_grid(n, 256), 256, 0, st>>>(ids, v1, n, k1);
        LSS_LAUNCH_CHECK();
        LSS_TRY(radix_sort_pairs(k1, v1, k2, v2, n, 0, bits_for(n_groups), 0, &ks, &vs, st));
    }
    group_median_kernel<<<ceil_div(n_groups, 128), 128, 0, st>>>(ks, vs, vals, n, n_groups, out);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}
