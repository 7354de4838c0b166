// Multi-GPU reduce of the near-field accumulators over NVLink peer memory (SURVEY section 8 row e).
//
// Every rank bins its block of launch rows into a private full-size grid.  Instead of an NCCL
// all-reduce of the three planes followed by the normalisation kernel, each rank maps the
// buffers of all ranks of the node into its address space (CUDA IPC) and ONE kernel per rank
//   * reads its slab of cells from every peer's re / im / count planes (P2P loads),
//   * sums them and the OPL statistics (phase reference = mean OPL over ALL ranks' rays),
//   * normalises to the complex field (bins_to_field_stats_kernel's arithmetic) and
//   * stores the slab into the field buffer of every destination rank (P2P stores),
// i.e. reduce-scatter + consumer + all-gather in one pass: 3 x 4 B read per cell and peer,
// 8 B written per cell and destination, nothing staged.  The two stream-ordered barriers around
// it ("all traces finished", "all slabs landed") are the caller's (dist.py: a one-element
// NCCL all-reduce each).
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace sosat {

constexpr int kMaxPeers = 16;

struct PeerArgs {
    const char *base[kMaxPeers];  // flat buffers of the peers (identical layout)
    char *out[kMaxPeers];         // buffers that receive the field slab
    int n_peers, n_out;
    long long re_off, im_off, cnt_off, stats_off, out_off;  // byte offsets from a base (field 0)
    // stack of (receiver, frequency) fields, blockIdx.y = irx * n_freq + f: plane `cells` long
    int n_freq;
    long long cells;
    double k_over_2e3[SOSAT_MAX_FREQS];
};

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float2 norm_cell(float r, float i, float c, float fcs, float fsn) {
    float2 v = make_float2(0.f, 0.f);
    if (c > 0.f) {
        const float inv = 1.0f / c, x = r * inv, y = i * inv;
        v.x = x * fcs + y * fsn;  // (x + i y) * exp(-i phi0)
        v.y = y * fcs - x * fsn;
    }
    return v;
}

__global__ void __launch_bounds__(256)
bins_reduce_field_kernel(const __grid_constant__ PeerArgs P, long long begin, long long end,
                         double opl_ref, double *__restrict__ stats_sum) {
    const int fld = blockIdx.y, irx = fld / P.n_freq, ifr = fld - irx * P.n_freq;
    const long long re_off = P.re_off + 4 * fld * P.cells, im_off = P.im_off + 4 * fld * P.cells;
    const long long cnt_off = P.cnt_off + 4 * irx * P.cells, out_off = P.out_off + 8 * fld * P.cells;
    const long long stats_off = P.stats_off + 8 * SOSAT_NUM_STATS * irx;
    const double k_over_2e3 = P.k_over_2e3[ifr];
    // statistics of the whole bundle: sum of the peers' rows (8 doubles each).  ONE thread per
    // statistic and block fetches them (every thread reading the same remote sectors would
    // serialise on the peers' L2: 1 ms at 8 GPUs, measured)
    __shared__ double s_st[SOSAT_NUM_STATS];
    if (threadIdx.x < SOSAT_NUM_STATS) {
        double v = 0.0;
        for (int p = 0; p < P.n_peers; ++p)
            v += __ldcg((const double *)(P.base[p] + stats_off) + threadIdx.x);
        s_st[threadIdx.x] = v;
        if (stats_sum && blockIdx.x == 0 && ifr == 0) stats_sum[SOSAT_NUM_STATS * irx + threadIdx.x] = v;
    }
    __syncthreads();
    double st[SOSAT_NUM_STATS];
#pragma unroll
    for (int q = 0; q < SOSAT_NUM_STATS; ++q) st[q] = s_st[q];
    const double nv = st[SOSAT_STAT_N_VALID];
    const double mean_opl = nv > 0 ? st[SOSAT_STAT_SUM_OPL] / nv : 0.0;
    double sn, cs;
    sincos(k_over_2e3 * (mean_opl - opl_ref), &sn, &cs);
    const float fcs = (float)cs, fsn = (float)sn;

    // 4 cells per thread and iteration: 3 x n_peers independent 16-byte P2P loads in flight
    const long long n4 = (end - begin) >> 2;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n4;
         g += (long long)gridDim.x * blockDim.x) {
        const long long i = begin + 4 * g;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f), m = r, c = r;
#pragma unroll 4
        for (int p = 0; p < P.n_peers; ++p) {
            r = add4(r, __ldcg((const float4 *)(P.base[p] + re_off) + (i >> 2)));
            m = add4(m, __ldcg((const float4 *)(P.base[p] + im_off) + (i >> 2)));
            c = add4(c, __ldcg((const float4 *)(P.base[p] + cnt_off) + (i >> 2)));
        }
        const float2 v0 = norm_cell(r.x, m.x, c.x, fcs, fsn), v1 = norm_cell(r.y, m.y, c.y, fcs, fsn);
        const float2 v2 = norm_cell(r.z, m.z, c.z, fcs, fsn), v3 = norm_cell(r.w, m.w, c.w, fcs, fsn);
        const float4 lo = make_float4(v0.x, v0.y, v1.x, v1.y), hi = make_float4(v2.x, v2.y, v3.x, v3.y);
        for (int d = 0; d < P.n_out; ++d) {
            float4 *o = (float4 *)(P.out[d] + out_off) + (i >> 1);
            o[0] = lo;
            o[1] = hi;
        }
    }
    // tail (slab length not a multiple of 4)
    for (long long i = begin + 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (long long)gridDim.x * blockDim.x) {
        float r = 0.f, m = 0.f, c = 0.f;
        for (int p = 0; p < P.n_peers; ++p) {
            r += __ldcg((const float *)(P.base[p] + re_off) + i);
            m += __ldcg((const float *)(P.base[p] + im_off) + i);
            c += __ldcg((const float *)(P.base[p] + cnt_off) + i);
        }
        const float2 v = norm_cell(r, m, c, fcs, fsn);
        for (int d = 0; d < P.n_out; ++d) ((float2 *)(P.out[d] + out_off))[i] = v;
    }
}

// ---- barrier over the ranks of the node through flags in peer-mapped memory --------------------
// Every rank owns an array of 64-bit slots in its exported buffer.  Rank r announces "epoch e
// reached" by storing e into slot r of EVERY peer's array (st.release.sys after a system fence:
// the kernels this rank launched before, in stream order, have completed, so their writes are
// performed) and waits until all slots of its own array hold >= e (ld.acquire.sys).  One tiny
// kernel per barrier, stream-ordered like the one-element NCCL all-reduce it replaces (8 us
// against 40 us at 8 GPUs).  Epochs only grow, so a peer that is already one barrier ahead does no
// harm.  A rank that waits longer than timeout_ns gives up, records it in slot 63 and returns --
// wrong data later (the callers check the slot) instead of a hung GPU.
constexpr int kBarrierErrSlot = 63;

__global__ void peer_barrier_kernel(PeerArgs P, int rank, long long flags_off,
                                    unsigned long long epoch, unsigned long long timeout_ns) {
    const int t = threadIdx.x;
    if (t >= P.n_peers) return;
    __threadfence_system();
    unsigned long long *remote = (unsigned long long *)(const_cast<char *>(P.base[t]) + flags_off) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    const unsigned long long *mine =
        (const unsigned long long *)(P.base[rank] + flags_off) + t;
    unsigned long long t0, now, seen = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            unsigned long long *err = (unsigned long long *)(const_cast<char *>(P.base[rank]) + flags_off) + kBarrierErrSlot;
            *err = epoch;
            break;
        }
        __nanosleep(200);
    }
}

}  // namespace sosat

using namespace sosat;

static_assert(sizeof(cudaIpcMemHandle_t) == SOSAT_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" int sosat_peer_alloc(int64_t bytes, void **d_ptr) {
    SOSAT_REQUIRE(d_ptr && bytes > 0, "sosat_peer_alloc: null pointer or non-positive size");
    // a cudaMalloc of its own: IPC handles name whole allocations, and memory handed out by a
    // caching or virtual-memory allocator cannot be exported
    SOSAT_CUDA_OK(cudaMalloc(d_ptr, (size_t)bytes));
    SOSAT_CUDA_OK(cudaMemset(*d_ptr, 0, (size_t)bytes));
    return SOSAT_OK;
}

extern "C" int sosat_peer_free(void *d_ptr) {
    if (d_ptr) SOSAT_CUDA_OK(cudaFree(d_ptr));
    return SOSAT_OK;
}

extern "C" int sosat_peer_export(const void *d_ptr, unsigned char *handle) {
    SOSAT_REQUIRE(d_ptr && handle, "sosat_peer_export: null pointer");
    cudaIpcMemHandle_t h;
    SOSAT_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, sizeof h);
    return SOSAT_OK;
}

extern "C" int sosat_peer_open(const unsigned char *handle, void **d_peer) {
    SOSAT_REQUIRE(handle && d_peer, "sosat_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    SOSAT_CUDA_OK(cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess));
    return SOSAT_OK;
}

extern "C" int sosat_peer_close(void *d_peer) {
    if (d_peer) SOSAT_CUDA_OK(cudaIpcCloseMemHandle(d_peer));
    return SOSAT_OK;
}

extern "C" int sosat_bins_reduce_fields(const void *const *h_peer_base, int n_peers,
                                        int64_t re_off, int64_t im_off, int64_t cnt_off,
                                        int64_t stats_off, int64_t n_cells, int n_rx, int n_freq,
                                        int64_t cell_begin, int64_t cell_end,
                                        const double *h_k_over_2e3, double opl_ref,
                                        void *const *h_out_base, int n_out, int64_t out_off,
                                        double *d_stats_sum, void *stream) {
    SOSAT_REQUIRE(h_peer_base && h_out_base && h_k_over_2e3, "sosat_bins_reduce_fields: null pointer");
    SOSAT_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && n_out >= 1 && n_out <= kMaxPeers,
                  "sosat_bins_reduce_fields: between 1 and %d peers / destinations", kMaxPeers);
    SOSAT_REQUIRE(n_rx >= 1 && n_freq >= 1 && n_freq <= SOSAT_MAX_FREQS &&
                      (long long)n_rx * n_freq <= 65535,
                  "n_rx x n_freq = %d x %d out of range", n_rx, n_freq);
    SOSAT_REQUIRE(cell_begin >= 0 && cell_end >= cell_begin && cell_end <= n_cells, "bad cell range");
    // (an empty slab -- more ranks than 4-cell units -- still launches: it sums the statistics)
    SOSAT_REQUIRE(cell_begin % 4 == 0 || cell_begin == cell_end, "cell_begin must be a multiple of 4");
    SOSAT_REQUIRE(re_off % 16 == 0 && im_off % 16 == 0 && cnt_off % 16 == 0 && out_off % 16 == 0 &&
                      stats_off % 8 == 0,
                  "sosat_bins_reduce_fields: misaligned plane offset");
    SOSAT_REQUIRE(n_rx * n_freq == 1 || n_cells % 4 == 0,
                  "a stack of fields needs a cell count per plane that is a multiple of 4 "
                  "(16-byte aligned planes); got %lld", (long long)n_cells);
    PeerArgs P;
    memset(&P, 0, sizeof P);
    for (int p = 0; p < n_peers; ++p) {
        SOSAT_REQUIRE(h_peer_base[p] && ((uintptr_t)h_peer_base[p] & 15) == 0, "peer base null or misaligned");
        P.base[p] = (const char *)h_peer_base[p];
    }
    for (int d = 0; d < n_out; ++d) {
        SOSAT_REQUIRE(h_out_base[d] && ((uintptr_t)h_out_base[d] & 15) == 0, "destination base null or misaligned");
        P.out[d] = (char *)h_out_base[d];
    }
    P.n_peers = n_peers;
    P.n_out = n_out;
    P.re_off = re_off; P.im_off = im_off; P.cnt_off = cnt_off; P.stats_off = stats_off;
    P.out_off = out_off;
    P.n_freq = n_freq;
    P.cells = n_cells;
    for (int f = 0; f < n_freq; ++f) P.k_over_2e3[f] = h_k_over_2e3[f];
    const int fields = n_rx * n_freq;
    const long long n = cell_end - cell_begin;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    long long cap = (long long)num_sms() * 8 / fields;
    if (cap < 4) cap = 4;
    if (blocks > cap) blocks = cap;
    bins_reduce_field_kernel<<<dim3((unsigned)blocks, (unsigned)fields), 256, 0, (cudaStream_t)stream>>>(
        P, cell_begin, cell_end, opl_ref, d_stats_sum);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_bins_reduce_field(const void *const *h_peer_base, int n_peers,
                                       int64_t re_off, int64_t im_off, int64_t cnt_off,
                                       int64_t stats_off, int64_t cell_begin, int64_t cell_end,
                                       double k_over_2e3, double opl_ref,
                                       void *const *h_out_base, int n_out, int64_t out_off,
                                       double *d_stats_sum, void *stream) {
    return sosat_bins_reduce_fields(h_peer_base, n_peers, re_off, im_off, cnt_off, stats_off,
                                    cell_end, 1, 1, cell_begin, cell_end, &k_over_2e3, opl_ref,
                                    h_out_base, n_out, out_off, d_stats_sum, stream);
}

extern "C" int sosat_peer_barrier(const void *const *h_peer_base, int n_peers, int rank,
                                  int64_t flags_off, uint64_t epoch, double timeout_s, void *stream) {
    SOSAT_REQUIRE(h_peer_base, "sosat_peer_barrier: null pointer");
    SOSAT_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && rank >= 0 && rank < n_peers,
                  "sosat_peer_barrier: rank %d of %d peers", rank, n_peers);
    SOSAT_REQUIRE(flags_off >= 0 && flags_off % 8 == 0 && epoch > 0, "bad flag offset or epoch");
    PeerArgs P;
    memset(&P, 0, sizeof P);
    for (int p = 0; p < n_peers; ++p) {
        SOSAT_REQUIRE(h_peer_base[p], "peer base null");
        P.base[p] = (const char *)h_peer_base[p];
    }
    P.n_peers = n_peers;
    const unsigned long long ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, rank, flags_off,
                                                            (unsigned long long)epoch, ns);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}
