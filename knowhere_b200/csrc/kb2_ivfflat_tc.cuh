// kb2_ivfflat_tc.cuh — list-major tensor-core engine of the IVF_FLAT scan (batched queries).
//
// Reference path being replaced: IVFFlatScanner::scan_codes (exact distance of ONE query to every row of a probed
// list, K/IndexIVFFlat.cpp:139-232) under IndexIVF::search_preassigned (F/IndexIVF.cpp:401-768): the reference streams
// each probed list once per (query, list) pair — C2: 16.0 MB per query.
//
// Here the query x list distance is what it is, a dense contraction: the (query, probe) pairs are grouped by list
// (same plan as the IVF_PQ engine), and every list is read ONCE per batch — a [128 rows] x [N queries of the list] x d
// tile product on tcgen05 with the operands brought by TMA.
//   * A operand: 128 consecutive rows of the list, fp32, straight from the list-order vector store (TMA, 128-byte swizzle).
//   * B operand: the item's queries, gathered pair-major beforehand and already split hi/lo (two TMA tiles).
//   * fp32 fidelity on the tf32 pipe: converter warps split the A tile into hi = tf32(x), lo = tf32(x - hi); the MMA warp
//     issues hi*hi + hi*lo + lo*hi (3 x kind::tf32, M=128, N=16..128, K=8), error ~2^-21 relative (kb2_gemm_tc.cuh).
//   * accumulators: two 128-column TMEM buffers; the epilogue warps of tile i run under the MMAs of tile i+1.
//   * epilogue: key = |q|^2 + |x|^2 - 2 acc (L2) / -acc (IP); keys within the per-query admission bound (exact k-th best
//     key of the query's nearest probed lists, from the query-major kernel, + a 3e-5 relative slack for the tf32 split)
//     are logged as survivors; finalize_kernel re-ranks the k+16 best of them EXACTLY from the fp32 rows, so the result
//     is the exact scan's (same guarantee as FLAT).
// Work item = (list, <=128 of the queries probing it).  One persistent CTA per SM, 320 threads:
//   warp 0 TMA producer | warp 1 MMA issuer + TMEM owner | warps 2-5 converters | warps 6-9 epilogue.
#pragma once
#include "kb2_gemm_tc.cuh"
#include "kb2_ivfpq_tc.cuh"

namespace kb2 {
namespace fltc {

constexpr int TM = 128;         // list rows per tile (UMMA M)
constexpr int NQ_ITEM = 128;    // most queries per item (UMMA N max)
constexpr int BK = 32;          // floats per k-block (one 128-byte swizzle row)
constexpr int TILE_BYTES = 128 * BK * 4;        // 16 KB: 128 rows of one k-block
constexpr int THREADS = 320;
// BROWS = query rows per item the instance is built for (its B tiles are BROWS x 128 B): 32 -> 40 KB stages, 5 in flight
// (few queries per list: C2 has ~31); 128 -> 64 KB stages, 3 in flight.  The kernel is bound by the latency of the
// TMA -> convert -> MMA -> release chain of a stage, so the number of stages in flight sets the HBM rate it reaches.
template <int BROWS>
struct FlCfg {
    static constexpr int B_TILE_BYTES = BROWS * BK * 4;
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * B_TILE_BYTES;     // A_hi | A_lo | B_hi | B_lo
    static constexpr int STAGES = BROWS <= 32 ? 5 : 3;
    static constexpr int OFF_META = STAGES * STAGE_BYTES;                     // thr[128] | base[128] | qidx[128]
    static constexpr int OFF_BAR = OFF_META + 3 * NQ_ITEM * 4;
    static constexpr size_t SMEM_BYTES = OFF_BAR + 256 + 384 /*item ring*/ + 1024 /*alignment slack*/;
    static_assert(SMEM_BYTES <= 227 * 1024, "IVF_FLAT tensor-core kernel shared memory");
};
constexpr float kSlack = 3e-5f;   // 3xTF32 contraction error, relative to |q|^2 + |x|^2 (measured 5e-6, tests/test_gemm_tc_gpu.py)

struct Params {
    int metric, d;
    const int32_t* n_items;
    int32_t* ticket;              // optional work counter (see pqtc::Params::ticket)
    const int32_t* item_list;
    const int32_t* item_q0;       // first pair of the item
    const int32_t* item_nq;
    const int32_t* pair_q;        // [pairs] query index, grouped by list
    const float* qnorm2;          // [nq] |q|^2
    const float* bound;           // [nq] admission bound on the key (+inf: none)
    const int64_t* list_off;
    const int32_t* list_len;
    const float* xnorm2;          // [npad] |x|^2 per position
    const uint8_t* bitset;
    const int32_t* rows;
    uint4* log;                   // [gridDim.x][log_cap] survivors {query, position, key bits, 0}
    uint32_t* log_cnt;            // [gridDim.x] entries; [gridDim.x] = 1 when a log overflowed
    uint32_t log_cap;
    unsigned long long* counters; // [0] rows scanned (pairs x rows)
};

// pair-major copy of the queries, split for the 3xTF32 contraction: hi = tf32(q), lo = tf32(q - hi)   (warp per pair)
__global__ void __launch_bounds__(256)
gather_split_queries_kernel(const float* __restrict__ q, const int32_t* __restrict__ pair_q, int64_t npairs, int64_t npairs_pad, int d,
                            float* __restrict__ hi, float* __restrict__ lo) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= npairs_pad) return;
    const int32_t qi = i < npairs ? pair_q[i] : -1;
    for (int j = lane; j < d; j += kWarp) {
        const float v = qi >= 0 ? q[(int64_t)qi * d + j] : 0.f;
        const float h = tc::tf32_rn(v);
        hi[i * d + j] = h;
        lo[i * d + j] = tc::tf32_rn(v - h);
    }
}

// bound[q] = key of entry k-1 of the query's sorted phase-A row (+inf when it holds fewer than k entries)
__global__ void
extract_bound_kernel(const uint64_t* __restrict__ partial, int64_t stride, int k, int64_t nq, float* __restrict__ bound) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t e = partial[q * stride + k - 1];
    bound[q] = (e == kEmpty) ? INFINITY : unpack_key(e);
}

// plan for 128-query items (the IVF_PQ engine's plan kernel cuts at 256)
__global__ void __launch_bounds__(1024)
plan_kernel(const int32_t* __restrict__ lcount, int nlist, int item_cap, int32_t* __restrict__ lstart, int32_t* __restrict__ item_list,
            int32_t* __restrict__ item_q0, int32_t* __restrict__ item_nq, int32_t* __restrict__ n_items) {
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp_a, tmp_b;
    __shared__ int carry_a, carry_b;
    if (threadIdx.x == 0) carry_a = carry_b = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nlist; b0 += 1024) {
        const int l = b0 + threadIdx.x;
        const int c = l < nlist ? lcount[l] : 0;
        const int nch = (c + item_cap - 1) / item_cap;
        int ex_a, ex_b, tot_a, tot_b;
        Scan(tmp_a).ExclusiveSum(c, ex_a, tot_a);
        Scan(tmp_b).ExclusiveSum(nch, ex_b, tot_b);
        const int ca = carry_a, cb = carry_b;
        if (l < nlist) {
            lstart[l] = ca + ex_a;
            if (nch > 0) {
                int per = ((c + nch - 1) / nch + 15) & ~15;
                if ((nch - 1) * per >= c || per > item_cap) per = item_cap;
                for (int ch = 0; ch < nch; ch++) {
                    const int i = cb + ex_b + ch;
                    item_list[i] = l;
                    item_q0[i] = ca + ex_a + ch * per;
                    item_nq[i] = max(0, min(per, c - ch * per));
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry_a = ca + tot_a;
            carry_b = cb + tot_b;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_items = carry_b;
}

__device__ __forceinline__ uint32_t
make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

template <int METRIC, int BROWS>
__global__ void __launch_bounds__(THREADS, 1)
ivfflat_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_qhi,
                  const __grid_constant__ CUtensorMap tmap_qlo, Params p) {
    using C = FlCfg<BROWS>;
    constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_TILE_BYTES = C::B_TILE_BYTES, OFF_META = C::OFF_META,
                  OFF_BAR = C::OFF_BAR;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = tc::smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_dyn + (base - raw);
    const uint32_t bars = base + OFF_BAR;
    // barriers: full_raw[S] full_conv[S] empty[S] acc_full[2] acc_empty[2] | tmem slot
    auto bar_full_raw = [&](int s) { return bars + 8u * s; };
    auto bar_full_conv = [&](int s) { return bars + 8u * (STAGES + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto bar_acc_full = [&](int i) { return bars + 8u * (3 * STAGES + i); };
    auto bar_acc_empty = [&](int i) { return bars + 8u * (3 * STAGES + 2 + i); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(sm + OFF_BAR + 8 * (3 * STAGES + 4));
    uint32_t* log_cursor = (uint32_t*)(sm + OFF_BAR + 8 * (3 * STAGES + 5));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = *p.n_items;
    const int nkb = p.d / BK;

    // item sequence of this CTA, drawn from a global counter and shared by the four roles (same scheme as the IVF_PQ filter
    // kernel; the roles are < 16 items apart: the producer leads by at most STAGES k-blocks, the epilogue trails by 2 tiles)
    constexpr int SCHED_R = 32;
    int* sch_claim = (int*)(sm + OFF_BAR + 256);
    int* sch_item = sch_claim + SCHED_R;
    volatile int* sch_ready = (volatile int*)(sch_item + SCHED_R);
    if (threadIdx.x < SCHED_R) {
        sch_claim[threadIdx.x] = (int)threadIdx.x - SCHED_R;
        sch_ready[threadIdx.x] = -1;
    }
    auto item_at_thread = [&](int seq) -> int {
        if (!p.ticket) return (int)blockIdx.x + seq * (int)gridDim.x;
        const int sl = seq & (SCHED_R - 1);
        if (sch_ready[sl] != seq) {
            if (atomicCAS(sch_claim + sl, seq - SCHED_R, seq) == seq - SCHED_R) {
                const int t = atomicAdd(p.ticket, 1);
                ((volatile int*)sch_item)[sl] = t;
                __threadfence_block();
                sch_ready[sl] = seq;
            } else {
                while (sch_ready[sl] != seq) {}
            }
        }
        __threadfence_block();
        return ((volatile int*)sch_item)[sl];
    };
    auto item_at = [&](int seq) -> int {   // warp-uniform call
        int v = 0;
        if (lane == 0) v = item_at_thread(seq);
        return __shfl_sync(0xffffffffu, v, 0);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            tc::mbar_init(bar_full_raw(s), 1);
            tc::mbar_init(bar_full_conv(s), 128);
            tc::mbar_init(bar_empty(s), 1);
        }
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(bar_acc_full(i), 1);
            tc::mbar_init(bar_acc_empty(i), 128);
        }
        *log_cursor = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer: per (item, tile, k-block) one stage = raw A tile + B_hi + B_lo =================
        if (lane == 0) {
            uint32_t it = 0;
            int seq = 0;
            for (int item = item_at_thread(0); item < n_items; item = item_at_thread(++seq)) {
                const int l = p.item_list[item];
                const int q0 = p.item_q0[item];
                const int64_t off = p.list_off[l];
                const int ntiles = (p.list_len[l] + TM - 1) / TM;
                for (int t = 0; t < ntiles; t++) {
                    for (int kb = 0; kb < nkb; kb++, it++) {
                        const int s = it % STAGES;
                        pqtc::mbar_wait_g(bar_empty(s), ((it / STAGES) & 1u) ^ 1u);
                        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
                        tc::mbar_expect_tx(bar_full_raw(s), TILE_BYTES + 2 * B_TILE_BYTES);
                        tc::tma_load_2d(st, &tmap_x, kb * BK, (int)(off + (int64_t)t * TM), bar_full_raw(s));
                        tc::tma_load_2d(st + 2 * TILE_BYTES, &tmap_qhi, kb * BK, q0, bar_full_raw(s));
                        tc::tma_load_2d(st + 2 * TILE_BYTES + B_TILE_BYTES, &tmap_qlo, kb * BK, q0, bar_full_raw(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        uint32_t it = 0, g = 0;
        int seq = 0;
        for (int item = item_at(0); item < n_items; item = item_at(++seq)) {
            const int l = p.item_list[item];
            const int nmma = (p.item_nq[item] + 15) & ~15;
            const int ntiles = (p.list_len[l] + TM - 1) / TM;
            const uint32_t idesc = make_idesc_tf32(nmma);
            for (int t = 0; t < ntiles; t++, g++) {
                const int buf = g & 1;
                pqtc::mbar_wait_g(bar_acc_empty(buf), ((g >> 1) & 1u) ^ 1u);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    pqtc::mbar_wait_g(bar_full_conv(s), (it / STAGES) & 1u);
                    tc::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
                        const uint32_t d_t = tmem_base + (uint32_t)buf * 128u;
#pragma unroll
                        for (int kk = 0; kk < BK / 8; kk++) {
                            const uint32_t ko = (uint32_t)kk * 32u;
                            const uint64_t a_hi = tc::make_desc(st + ko);
                            const uint64_t a_lo = tc::make_desc(st + TILE_BYTES + ko);
                            const uint64_t b_hi = tc::make_desc(st + 2 * TILE_BYTES + ko);
                            const uint64_t b_lo = tc::make_desc(st + 2 * TILE_BYTES + B_TILE_BYTES + ko);
                            tc::tc_mma_tf32(d_t, a_hi, b_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
                            tc::tc_mma_tf32(d_t, a_hi, b_lo, idesc, 1u);
                            tc::tc_mma_tf32(d_t, a_lo, b_hi, idesc, 1u);
                        }
                        tc::tc_commit(bar_empty(s));
                        if (kb == nkb - 1) tc::tc_commit(bar_acc_full(buf));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 6) {
        // ================= converters: split the raw A tile into hi (in place) and lo =================
        const int t128 = threadIdx.x - 64;   // 0..127
        uint32_t it = 0;
        int seq = 0;
        for (int item = item_at(0); item < n_items; item = item_at(++seq)) {
            const int l = p.item_list[item];
            const int ntiles = (p.list_len[l] + TM - 1) / TM;
            for (int t = 0; t < ntiles; t++) {
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    pqtc::mbar_wait_g(bar_full_raw(s), (it / STAGES) & 1u);
                    float4* hi = reinterpret_cast<float4*>(sm + (size_t)s * STAGE_BYTES);
                    float4* lo = reinterpret_cast<float4*>(sm + (size_t)s * STAGE_BYTES + TILE_BYTES);
#pragma unroll 4
                    for (int i = t128; i < TILE_BYTES / 16; i += 128) {
                        const float4 v = hi[i];
                        float4 h, w;
                        h.x = tc::tf32_rn(v.x); w.x = tc::tf32_rn(v.x - h.x);
                        h.y = tc::tf32_rn(v.y); w.y = tc::tf32_rn(v.y - h.y);
                        h.z = tc::tf32_rn(v.z); w.z = tc::tf32_rn(v.z - h.z);
                        h.w = tc::tf32_rn(v.w); w.w = tc::tf32_rn(v.w - h.w);
                        hi[i] = h;
                        lo[i] = w;
                    }
                    tc::fence_proxy_async();
                    tc::mbar_arrive(bar_full_conv(s));
                }
            }
        }
    } else {
        // ================= epilogue: thread = list row, columns = the item's queries =================
        const int e = threadIdx.x - 192;     // 0..127
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access (warps 6..9 -> 2,3,0,1)
        const int row = quarter * 32 + lane;
        float* m_thr = (float*)(sm + OFF_META);
        float* m_base = m_thr + NQ_ITEM;
        int* m_q = (int*)(m_base + NQ_ITEM);
        uint4* my_log = p.log + (size_t)blockIdx.x * p.log_cap;
        bool log_over = false;
        unsigned long long n_rows = 0;
        uint32_t g = 0;
        int seq = 0;
        for (int item = item_at(0); item < n_items; item = item_at(++seq)) {
            const int l = p.item_list[item];
            const int q0 = p.item_q0[item];
            const int nqi = p.item_nq[item];
            const int nmma = (nqi + 15) & ~15;
            const int nch = (nmma + 31) >> 5;
            const int len = p.list_len[l];
            const int64_t off = p.list_off[l];
            const int ntiles = (len + TM - 1) / TM;
            asm volatile("bar.sync 1, 128;" ::: "memory");   // every epilogue thread is done with the previous item's meta
            {
                // admit  <=>  key - slack * (|q|^2 + |x|^2) <= bound   (|q|^2 + |x|^2 >= 2 |q||x| bounds the 3xTF32 error scale)
                float thr = -INFINITY, bs = 0.f;
                int q = -1;
                if (e < nqi) {
                    q = p.pair_q[q0 + e];
                    bs = p.qnorm2[q];
                    const float bnd = p.bound[q];
                    thr = (bnd < INFINITY) ? bnd + kSlack * bs + 1e-30f : INFINITY;
                }
                m_thr[e] = thr;
                m_base[e] = bs;
                m_q[e] = q;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e == 0) n_rows += (unsigned long long)len * (unsigned long long)nqi;
            for (int t = 0; t < ntiles; t++, g++) {
                const int buf = g & 1;
                pqtc::mbar_wait_g(bar_acc_full(buf), (g >> 1) & 1u);
                tc::tc_fence_after();
                const int rel = t * TM + row;
                const bool row_ok = rel < len;
                float xn = 0.f;
                if (row_ok) xn = __ldg(p.xnorm2 + off + rel);
                bool alive = row_ok;
                if (alive && p.bitset) alive = !bit_is_set(p.bitset, p.rows[off + rel]);
                const float xs = (METRIC == KB2_METRIC_L2) ? xn * (1.f - kSlack) : -kSlack * xn;   // (row part of the key) - slack * |x|^2
                const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 128);
#pragma unroll 1
                for (int ci = 0; ci < nch; ci++) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                          "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr0 + (uint32_t)(ci * 32)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (ci == nch - 1) {   // the accumulator is in registers: hand it back
                        tc::tc_fence_before();
                        tc::mbar_arrive(bar_acc_empty(buf));
                    }
                    uint32_t mask = 0;
                    if (alive) {
#pragma unroll
                        for (int u = 0; u < 32; u++) {
                            const int col = ci * 32 + u;
                            const float acc = __uint_as_float(v[u]);
                            const float key = (METRIC == KB2_METRIC_L2) ? (m_base[col] + xs - 2.f * acc) : (xs - acc);
                            mask |= (col < nqi && key <= m_thr[col]) ? (1u << u) : 0u;
                        }
                    }
                    const uint32_t cnt = __popc(mask);
                    if (__any_sync(0xffffffffu, cnt != 0u)) {
                        uint32_t incl = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t tv = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += tv;
                        }
                        uint32_t wbase = 0;
                        if (lane == 31) wbase = atomicAdd(log_cursor, incl);
                        wbase = __shfl_sync(0xffffffffu, wbase, 31);
                        uint32_t slot = wbase + incl - cnt;
#pragma unroll
                        for (int u = 0; u < 32; u++) {   // static register indices (a data-dependent v[u] would spill the tile)
                            if ((mask >> u) & 1u) {
                                const int col = ci * 32 + u;
                                const float acc = __uint_as_float(v[u]);
                                const float key = (METRIC == KB2_METRIC_L2) ? (m_base[col] + xn - 2.f * acc) : -acc;
                                if (slot < p.log_cap) {
                                    uint4 o;
                                    o.x = (uint32_t)m_q[col];
                                    o.y = (uint32_t)(off + rel);
                                    o.z = __float_as_uint(key);
                                    o.w = 0u;
                                    my_log[slot] = o;
                                } else {
                                    log_over = true;
                                }
                                slot++;
                            }
                        }
                    }
                }
            }
        }
        if (log_over) p.log_cnt[gridDim.x] = 1u;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (e == 0) {
            p.log_cnt[blockIdx.x] = min(*log_cursor, p.log_cap);
            if (p.counters) atomicAdd(p.counters, n_rows);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// items that cut ONE pseudo-list of rows (the coarse quantizer's centroids) into chunks of `cap` consecutive queries
__global__ void
uniform_items_kernel(int64_t nq, int cap, int32_t* __restrict__ item_list, int32_t* __restrict__ item_q0, int32_t* __restrict__ item_nq,
                     int32_t* __restrict__ n_items, int32_t* __restrict__ pair_q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ni = (nq + cap - 1) / cap;
    if (i == 0) *n_items = (int32_t)ni;
    if (i < ni) {
        item_list[i] = 0;
        item_q0[i] = (int32_t)(i * cap);
        item_nq[i] = (int32_t)min((int64_t)cap, nq - i * cap);
    }
    if (i < nq) pair_q[i] = (int32_t)i;
}
// queries whose candidate row cannot serve the selection (fewer than `need` entries, or overflowed) -> *out += 1 each
__global__ void
check_counts_kernel(const uint32_t* __restrict__ cand_cnt, const uint32_t* __restrict__ qflag, int64_t nq, uint32_t need,
                    const uint32_t* __restrict__ log_over, unsigned long long* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && *log_over) atomicAdd(out, 1ull);
    if (i < nq && (qflag[i] || cand_cnt[i] < need)) atomicAdd(out, 1ull);
}

// survivors of all CTA logs -> per-query candidate rows as packed (key, position)   (grid = (x, number of logs))
__global__ void
scatter_kernel(const uint4* __restrict__ log, const uint32_t* __restrict__ log_cnt, uint32_t log_cap, uint64_t* __restrict__ cand,
               uint32_t* __restrict__ cand_cnt, int cap, uint32_t* __restrict__ qflag) {
    const uint32_t n = min(log_cnt[blockIdx.y], log_cap);
    const uint4* src = log + (size_t)blockIdx.y * log_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 e = src[i];
        const uint32_t slot = atomicAdd(cand_cnt + e.x, 1u);
        if (slot < (uint32_t)cap) cand[(int64_t)e.x * cap + slot] = pack_kp(__uint_as_float(e.z), e.y);
        else qflag[e.x] = 1u;
    }
}
// number of flagged queries (or every query when a log overflowed) -> *out
__global__ void
count_flags_kernel(const uint32_t* __restrict__ qflag, int64_t nq, const uint32_t* __restrict__ log_over, uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && *log_over) atomicAdd(out, 1u);
    if (i < nq && qflag[i]) atomicAdd(out, 1u);
}

}  // namespace fltc
}  // namespace kb2
