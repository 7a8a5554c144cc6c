// Grid-structured normal-matrix product q += N p (include/lssurf_b200.h: lss_solver_set_structured).
//
// One launch, everything accumulated into a zeroed q with fp64 REDs; what is left of N (bias columns) goes through the
// nnz-tiled CSR kernel in accumulate mode.
#include <stdlib.h>

#include "structured.cuh"

namespace lss {

namespace {

constexpr int SN_BLOCK = 256;
constexpr int SN_ROUNDS = 8;          // structured.py MAX_ROUNDS: z0-dz pairs of one dz node <= SN_ROUNDS * S
constexpr int SN_STAGES = 2;          // ring of node stages in shared memory (2 stages: 4 CTAs per SM; 3 stages / 3 CTAs measured 2 % slower)
constexpr int SN_HEAD = 128;          // bytes in front of the ring: mbarriers + per-stage descriptors

struct SnArgs {
    int T, S, nn, n0, n_diag, dd_stride, node_len, rec_len, stage_len;
    long long z0c0, dzc0;
    const double *node_val, *zd_val, *zz_val;
    const int *zd_ptr, *zz_off;
    // remainder as a flat entry list sorted by row: row of every entry, column, value
    const int *rem_row, *rem_col;
    const double *rem_val;
    long long rem_nnz;
    unsigned long long *trace;      // LSS_SN_TRACE: cycles of CTA phases, summed over thread 0 of all CTAs [8]
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned sn_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sn_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(sn_smem(bar)) : "memory");
}
__device__ __forceinline__ void sn_mbar_expect(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(sn_smem(bar)) : "memory");
}
__device__ __forceinline__ bool sn_mbar_try(unsigned long long *bar, unsigned parity) {
    unsigned done;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done) : "r"(sn_smem(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void sn_bulk(void *dst, const void *src, unsigned bytes, unsigned long long *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(sn_smem(dst)), "l"(src), "r"(bytes), "r"(sn_smem(bar)), "l"(pol) : "memory");
}

// Persistent CTAs, one dz node per iteration (node = blockIdx.x + r * gridDim.x).  Everything N holds for the node -- its
// block of neighbour / self diagonals and its z0-dz pair records -- is two contiguous byte ranges of the value buffer, which
// the last thread of the CTA requests as 1-D bulk copies (TMA) into a ring of SN_STAGES shared-memory stages, SN_STAGES nodes
// ahead: the matrix stream costs no registers, no LSU instructions and never waits for the consumers.  The gathers of p for
// node i + 1 (ids arrived with its stage) are in flight while node i is worked on.  Per node: thread <-> (slot, t) forms the
// neighbour / self products, thread <-> pair sums the z0 side of its record in registers and leaves the dz-side products in
// place, then the rows (node, t) are summed over pairs and slots: ONE fp64 RED per row of q and node.  After its nodes a
// CTA takes a share of the z0-z0 diagonals (one thread per z0 row).  TT: epochs per node at compile time (0: run time).
template <int TT>
__global__ void __launch_bounds__(SN_BLOCK)
sn_product_kernel(SnArgs a, const double *__restrict__ p, double *__restrict__ q, const double *__restrict__ stop,
                  double *__restrict__ dot_slots) {
    extern __shared__ __align__(128) unsigned char sn_ring[];
    __shared__ double accs[SN_BLOCK];                   // dz-side partial sums per (slot, t)
    __shared__ double pbs[2][2 * SN_BLOCK];             // p_dz of the neighbour nodes [slot][t], double-buffered by node parity
    __shared__ double pas[2][32];                       // p_dz of the node itself
    if (stop && *stop != 0.0) return;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sn_ring);          // [0, 32): mbarriers
    int *desc = reinterpret_cast<int *>(sn_ring + 32);                                    // 4 ints per stage: node, zb0, zb1
    double *ring = reinterpret_cast<double *>(sn_ring + SN_HEAD);
    const uint64_t once = l2_policy_evict_first();
    const int tid = threadIdx.x;
    const int T = TT ? TT : a.T, S = TT ? SN_BLOCK / TT : a.S;
    const int rec_len = TT ? (TT + 2) / 2 * 2 : a.rec_len;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SN_STAGES; ++s) sn_mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n0 = blockIdx.x, dn = gridDim.x;
    const bool producer = tid == SN_BLOCK - 1;
    // producer: pair range of the next node to request, loaded one iteration before it is needed
    int d_zb0 = 0, d_zb1 = 0;
    auto load_desc = [&](int node) {
        if (node < a.nn) {
            d_zb0 = __ldg(a.zd_ptr + node);
            d_zb1 = __ldg(a.zd_ptr + node + 1);
        }
    };
    auto issue = [&](int node, int stage) {
        if (node >= a.nn) return;
        double *base = ring + (size_t)stage * a.stage_len;
        int *dd = desc + stage * 4;
        dd[0] = node; dd[1] = d_zb0; dd[2] = d_zb1;
        const unsigned b_node = (unsigned)a.node_len * 8u, b_rec = (unsigned)(d_zb1 - d_zb0) * (unsigned)rec_len * 8u;
        sn_mbar_expect(full + stage, b_node + b_rec);
        sn_bulk(base, a.node_val + (size_t)node * a.node_len, b_node, full + stage, once);
        if (b_rec) sn_bulk(base + a.node_len, a.zd_val + (size_t)d_zb0 * rec_len, b_rec, full + stage, once);
    };
    if (producer) {
#pragma unroll
        for (int s = 0; s < SN_STAGES; ++s) {
            load_desc(n0 + s * dn);
            issue(n0 + s * dn, s);
        }
        load_desc(n0 + SN_STAGES * dn);
    }
    const int slot = tid / T, t = tid - slot * T;
    const bool valid = slot < S;
    const int ds = a.dd_stride, self_off = 3 * ds * T + (ds + 1) / 2;
    const bool two = ds > S;                    // a second round of neighbour slots (uniform)
    double dot = 0.0;
    long long tr[6] = {0, 0, 0, 0, 0, 0}, c0 = clock64();
#ifdef LSS_SN_TRACE_BUILD
#define SN_T(k) do { if (a.trace) { const long long c1 = clock64(); tr[k] += c1 - c0; c0 = c1; } } while (0)
#else
#define SN_T(k) do { } while (0)
#endif
    // gathers of p for one node, into registers: the node's own epochs (tid < T), the neighbour rows (thread <-> (slot, t),
    // two rounds), the z0 value of pair tid
    double g_pa = 0.0, g_pb[2] = {0.0, 0.0}, g_pk = 0.0;
    long long g_k = -1;
    int g_np = 0;
    auto gather = [&](int node, int it) {
        const int stage = it % SN_STAGES;
        const unsigned parity = (unsigned)(it / SN_STAGES) & 1u;
        while (!sn_mbar_try(full + stage, parity)) {}
        const double *blk = ring + stage * a.stage_len;
        const double *rec = blk + a.node_len;
        const int *ids = reinterpret_cast<const int *>(blk + 3 * ds * T);
        g_np = desc[stage * 4 + 2] - desc[stage * 4 + 1];
        g_pa = tid < T ? __ldg(p + a.dzc0 + (int64_t)node * T + tid) : 0.0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = slot + r * S;
            g_pb[r] = 0.0;
            if (r == 1 && !two) break;
            if (valid && j < ds) {
                const int b = ids[j];
                if (b >= 0) g_pb[r] = __ldg(p + a.dzc0 + (int64_t)b * T + t);
            }
        }
        g_k = tid < g_np ? reinterpret_cast<const long long *>(rec)[tid * rec_len + T] : -1;
        g_pk = g_k >= 0 ? __ldg(p + g_k) : 0.0;
    };
    auto publish = [&](int buf) {        // gathered values -> shared memory (read after the next CTA barrier)
        if (tid < T) pas[buf][tid] = g_pa;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = slot + r * S;
            if (r == 1 && !two) break;
            if (valid && j < ds) pbs[buf][j * T + t] = g_pb[r];
        }
    };
    if (n0 < a.nn) {
        gather(n0, 0);
        publish(0);
    }
    __syncthreads();
    int i = 0;
    for (int node = n0; node < a.nn; node += dn, ++i) {
        const int stage = i % SN_STAGES, buf = i & 1;
        double *blk = ring + stage * a.stage_len;
        double *rec = blk + a.node_len;
        const int *ids = reinterpret_cast<const int *>(blk + 3 * ds * T);
        const int np = g_np;
        const long long kz = g_k;
        const double pkz = g_pk;
        const int64_t col0 = a.dzc0 + (int64_t)node * T;
        SN_T(0);
        // ---- gathers for the next node: in flight during the work on this one
        const bool more = node + dn < a.nn;
        if (more) gather(node + dn, i + 1);
        SN_T(1);
        // ---- phase 1a: neighbour and self diagonals, thread <-> (slot, t)
        double acc = 0.0;
        if (valid) {
            if (slot < 5) {
                const int ts = t + slot - 2;                                 // self band slot - 2
                if (ts >= 0 && ts < T) acc = blk[self_off + slot * T + t] * pas[buf][ts];
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int j = slot + r * S;
                if (r == 1 && !two) break;
                if (j < ds && ids[j] >= 0) {
                    const double *pb = pbs[buf] + j * T + t;
                    acc += blk[(ds + j) * T + t] * pb[0];
                    if (t > 0) acc += blk[j * T + t] * pb[-1];
                    if (t + 1 < T) acc += blk[(2 * ds + j) * T + t] * pb[1];
                }
            }
        }
        // ---- phase 1b: z0-dz records, thread <-> pair: z0 side summed in registers, dz-side products written back in place
        if (kz >= 0) {
            double2 *R2 = reinterpret_cast<double2 *>(rec + tid * rec_len);
            double s0 = 0.0, s1 = 0.0;
            const int half = rec_len >> 1;
#pragma unroll
            for (int h = 0; h < (TT ? (TT + 2) / 2 : 1); ++h) {
                if (TT) {
                    double2 v = R2[h];
                    if (2 * h < T) { s0 += v.x * pas[buf][2 * h]; v.x *= pkz; }
                    if (2 * h + 1 < T) { s1 += v.y * pas[buf][2 * h + 1]; v.y *= pkz; }
                    R2[h] = v;
                }
            }
            if (!TT) {
                for (int h = 0; h < half; ++h) {
                    double2 v = R2[h];
                    if (2 * h < T) { s0 += v.x * pas[buf][2 * h]; v.x *= pkz; }
                    if (2 * h + 1 < T) { s1 += v.y * pas[buf][2 * h + 1]; v.y *= pkz; }
                    R2[h] = v;
                }
            }
            atomicAdd(q + kz, s0 + s1);
            dot += (s0 + s1) * pkz;
        }
        __syncthreads();
        SN_T(2);
        // ---- phase 2: dz side: row (node, t) = diagonals + sum over the pairs, first per slot, then over the slots
        if (valid) {
            double a0 = 0.0, a1 = 0.0;
            int pair = slot;
            for (; pair + S < np; pair += 2 * S) {
                a0 += rec[pair * rec_len + t];
                a1 += rec[(pair + S) * rec_len + t];
            }
            if (pair < np) a0 += rec[pair * rec_len + t];
            acc += a0 + a1;
        }
        accs[tid] = valid ? acc : 0.0;
        if (more) publish(buf ^ 1);
        __syncthreads();                   // also: the stage may be overwritten from here on
        SN_T(3);
        if (producer) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic accesses to the stage before the bulk copy into it
            issue(node + SN_STAGES * dn, stage);
            load_desc(node + (SN_STAGES + 1) * dn);
        }
        if (tid < T) {
            double sd0 = 0.0, sd1 = 0.0;
            int sl = 0;
            for (; sl + 1 < S; sl += 2) {
                sd0 += accs[sl * T + tid];
                sd1 += accs[(sl + 1) * T + tid];
            }
            if (sl < S) sd0 += accs[sl * T + tid];
            atomicAdd(q + col0 + tid, sd0 + sd1);
            dot += (sd0 + sd1) * pas[buf][tid];
        }
        SN_T(4);
    }
    if (a.trace && tid == 0)
        for (int k = 0; k < 6; ++k) atomicAdd(a.trace + k, (unsigned long long)tr[k]);
    // ---- zz rows: one thread per z0 node, shared out over the CTAs
    for (int64_t i0 = (int64_t)blockIdx.x * SN_BLOCK + tid; i0 < a.n0; i0 += (int64_t)gridDim.x * SN_BLOCK) {
        double acc = 0.0;
        for (int d = 0; d < a.n_diag; ++d) {
            const int64_t j = i0 + __ldg(a.zz_off + d);
            if (j >= 0 && j < a.n0) acc += ldg_stream(a.zz_val + (size_t)d * a.n0 + i0, once) * __ldg(p + a.z0c0 + j);
        }
        atomicAdd(q + a.z0c0 + i0, acc);
        dot += acc * __ldg(p + a.z0c0 + i0);
    }
    // ---- remainder (bias columns ...): every warp takes 128 consecutive entries at a time (4 independent loads per lane, then
    // 4 gathers), folds equal rows with a segmented warp scan and adds one value per (warp, row) with a RED
    {
        const int lane = tid & 31;
        const long long warp = ((long long)blockIdx.x * SN_BLOCK + tid) >> 5, n_warps = ((long long)gridDim.x * SN_BLOCK) >> 5;
        for (long long base = warp * 128; base < a.rem_nnz; base += n_warps * 128) {
            int key[4], col[4];
            double av[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long k = base + 32 * j + lane;
                const bool in = k < a.rem_nnz;
                key[j] = in ? ldg_stream(a.rem_row + k, once) : -1;
                col[j] = in ? ldg_stream(a.rem_col + k, once) : 0;
                av[j] = in ? ldg_stream(a.rem_val + k, once) : 0.0;
            }
            double val[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) val[j] = key[j] >= 0 ? av[j] * __ldg(p + col[j]) : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double v = val[j];
                const int ky = key[j];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double v2 = __shfl_up_sync(0xffffffffu, v, o);
                    const int k2 = __shfl_up_sync(0xffffffffu, ky, o);
                    if (lane >= o && k2 == ky) v += v2;
                }
                const int kn = __shfl_down_sync(0xffffffffu, ky, 1);
                if (ky >= 0 && (lane == 31 || kn != ky)) {
                    atomicAdd(q + ky, v);
                    dot += v * __ldg(p + ky);
                }
            }
        }
    }
    if (dot_slots) {
        dot = warp_sum_d(dot);
        if ((tid & 31) == 0 && dot != 0.0) atomicAdd(dot_slots + ((blockIdx.x * 8u + (tid >> 5)) & 255u), dot);
    }
}

static const void *sn_kernel_for(int T) {
    return T == 21 ? (const void *)sn_product_kernel<21> : (const void *)sn_product_kernel<0>;
}

// rem_row[k] = row of entry k of the remainder CSR (one warp per compact row)
__global__ void __launch_bounds__(256)
sn_expand_rows_kernel(int n_rows, const int *__restrict__ ptr, const int *__restrict__ rows, int *__restrict__ out) {
    const int r = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
    if (r >= n_rows) return;
    const int row = rows[r];
    for (int k = ptr[r] + (threadIdx.x & 31); k < ptr[r + 1]; k += 32) out[k] = row;
}

// ids that travel inside the value buffer: neighbour ids of every node block, z0 column of every pair record
__global__ void __launch_bounds__(256)
sn_ids_kernel(int nn, int ds, int T, int node_len, const int *__restrict__ dd_b, int n_zd, int rec_len, const int *__restrict__ zd_k,
              double *__restrict__ node_val, double *__restrict__ zd_val) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < (int64_t)nn * ds) {
        const int node = (int)(i / ds), j = (int)(i - (int64_t)node * ds);
        reinterpret_cast<int *>(node_val + (size_t)node * node_len + 3 * ds * T)[j] = dd_b[i];
    }
    if (ds & 1) {                        // the odd id of the last pair of ints
        if (i < nn) reinterpret_cast<int *>(node_val + (size_t)i * node_len + 3 * ds * T)[ds] = -1;
    }
    if (i < n_zd) reinterpret_cast<long long *>(zd_val)[(size_t)i * rec_len + T] = (long long)zd_k[i];
}

__global__ void __launch_bounds__(256)
sn_refresh_kernel(int64_t nnz, const int *__restrict__ map, const double *__restrict__ data, double *__restrict__ vals) {
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * 256) {
        const int m = map[k];
        if (m >= 0) vals[m] = data[k];
    }
}

__global__ void __launch_bounds__(256)
sn_gather_kernel(int64_t n, const int *__restrict__ src, const double *__restrict__ data, double *__restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) out[k] = data[src[k]];
}

}  // namespace

int structured_create(StructuredN **out, const lss_structured_t *desc, cudaStream_t stream) {
    if (!desc || desc->nt < 2 || desc->nt > 32 || desc->n_vals <= 0 || !desc->map || desc->dd_stride > 2 * (SN_BLOCK / desc->nt) ||
        desc->max_pairs > SN_ROUNDS * (SN_BLOCK / desc->nt) || (desc->node_len & 1) || (desc->rec_len & 1) || (desc->zd_base & 1) || desc->nn < 1)
        return fail(LSS_EINVAL, "structured normal matrix: bad description");
    StructuredN *s = new StructuredN();
    memset(s, 0, sizeof(*s));
    s->d = *desc;
    s->stream = stream;
    Scratch::configure_pool();
    if (cudaMallocAsync((void **)&s->vals, sizeof(double) * (size_t)desc->n_vals, stream) != cudaSuccess ||
        cudaMallocAsync((void **)&s->rem_data, sizeof(double) * (size_t)(desc->rem_nnz > 0 ? desc->rem_nnz : 1), stream) != cudaSuccess) {
        structured_destroy(s);
        return fail(LSS_ECUDA, "structured normal matrix: cudaMallocAsync failed");
    }
    cudaMemsetAsync(s->vals, 0, sizeof(double) * (size_t)desc->n_vals, stream);
    {
        const int64_t n_ids = (int64_t)desc->nn * desc->dd_stride;
        const int64_t n = n_ids > desc->n_zd ? n_ids : desc->n_zd;
        sn_ids_kernel<<<ceil_div((n > desc->nn ? n : desc->nn) + 1, 256), 256, 0, stream>>>(desc->nn, desc->dd_stride, desc->nt, desc->node_len, desc->dd_b, desc->n_zd,
                                                                                       desc->rec_len, desc->zd_k, s->vals, s->vals + desc->zd_base);
    }
    // ring geometry: a stage holds one node block and the records of the largest node
    s->stage_len = (desc->node_len + desc->max_pairs * desc->rec_len + 15) & ~15;
    s->smem = SN_HEAD + (size_t)SN_STAGES * s->stage_len * sizeof(double);
    if (cudaFuncSetAttribute(sn_kernel_for(desc->nt), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem) != cudaSuccess) {
        structured_destroy(s);
        return fail(LSS_EINVAL, "structured normal matrix: %zu bytes of shared memory per CTA do not fit", s->smem);
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sn_kernel_for(desc->nt), SN_BLOCK, s->smem);
    if (per_sm < 1) { structured_destroy(s); return fail(LSS_EINVAL, "structured normal matrix: the product kernel does not fit on an SM"); }
    s->grid = kSMs * per_sm < desc->nn ? kSMs * per_sm : desc->nn;
    if (desc->rem_nnz > 0) {
        if (cudaMallocAsync((void **)&s->rem_row, sizeof(int) * (size_t)desc->rem_nnz, stream) != cudaSuccess) {
            structured_destroy(s);
            return fail(LSS_ECUDA, "structured normal matrix: cudaMallocAsync failed");
        }
        sn_expand_rows_kernel<<<ceil_div((int64_t)desc->rem_n_rows * 32, 256), 256, 0, stream>>>(desc->rem_n_rows, desc->rem_indptr, desc->rem_rows, s->rem_row);
    }
    *out = s;
    return LSS_OK;
}

void structured_destroy(StructuredN *s) {
    if (!s) return;
    if (s->vals) cudaFreeAsync(s->vals, s->stream);
    if (s->rem_data) cudaFreeAsync(s->rem_data, s->stream);
    if (s->rem_row) cudaFreeAsync(s->rem_row, s->stream);
    delete s;
}

int structured_refresh(StructuredN *s, const double *csr_data) {
    sn_refresh_kernel<<<stride_grid(s->d.nnz, 256, 16), 256, 0, s->stream>>>(s->d.nnz, s->d.map, csr_data, s->vals);
    LSS_LAUNCH_CHECK();
    if (s->d.rem_nnz > 0) {
        sn_gather_kernel<<<stride_grid(s->d.rem_nnz, 256, 16), 256, 0, s->stream>>>(s->d.rem_nnz, s->d.rem_src, csr_data, s->rem_data);
        LSS_LAUNCH_CHECK();
    }
    return LSS_OK;
}

int structured_apply(StructuredN *s, const double *p, double *q, const double *stop, int64_t *launches, double *dot_slots) {
    const lss_structured_t &d = s->d;
    SnArgs a;
    a.T = d.nt; a.S = SN_BLOCK / d.nt; a.nn = d.nn; a.n0 = d.n0; a.n_diag = d.n_diag; a.dd_stride = d.dd_stride;
    a.node_len = d.node_len; a.rec_len = d.rec_len; a.stage_len = s->stage_len;
    a.z0c0 = d.z0_col0; a.dzc0 = d.dz_col0;
    a.node_val = s->vals; a.zd_val = s->vals + d.zd_base; a.zz_val = s->vals + d.zz_base;
    a.zd_ptr = d.zd_ptr; a.zz_off = d.zz_off;
    a.rem_row = s->rem_row; a.rem_col = d.rem_indices; a.rem_val = s->rem_data; a.rem_nnz = d.rem_nnz;
    static unsigned long long *trace_d = nullptr;
    static int trace_calls = 0;
    if (getenv("LSS_SN_TRACE") && !trace_d) {
        cudaMalloc((void **)&trace_d, 64);
        cudaMemset(trace_d, 0, 64);
    }
    a.trace = trace_d;
    {
        void *args[] = {&a, &p, &q, &stop, &dot_slots};
        LSS_CUDA(cudaLaunchKernel(sn_kernel_for(d.nt), dim3(s->grid), dim3(SN_BLOCK), args, s->smem, s->stream));
        ++lss::g_launches;
    }
    if (trace_d && ++trace_calls == 200) {
        unsigned long long h[8];
        cudaMemcpy(h, trace_d, 64, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[lss sn trace] cycles per node (thread 0): top %.0f gather-issue %.0f phase1 %.0f phase2 %.0f rows %.0f (%.0f)\n",
                h[0] / 200.0 / d.nn, h[1] / 200.0 / d.nn, h[2] / 200.0 / d.nn, h[3] / 200.0 / d.nn, h[4] / 200.0 / d.nn, h[5] / 200.0 / d.nn);
    }
    if (launches) *launches += 1;
    return LSS_OK;
}

}  // namespace lss
