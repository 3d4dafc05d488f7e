// kb2_ivfpq_tc.cuh — list-major tensor-core engine of the IVF_PQ scan (large batches).
//
// Reference path being replaced: IVFPQScanner::scan_list_with_table (one table look-up chain per
// (query, code) pair; F/impl/pq_code_distance/IVFPQScanner_impl.h:110-185) under
// IndexIVF::search_preassigned (F/IndexIVF.cpp:401-768).
//
// Why a second engine.  The query-major LUT kernel (kb2_ivf.cuh) is bound by the shared-memory gather pipe:
// 16 wavefronts per 32 codes *per query* (profiles/r1_final_scan_kernel.md).  But the ADC inner term is a dot
// product, <q, r^(code)>, and at batch 10^4 x nprobe 64 every list is probed by ~150 queries.  Decoding a
// 128-code tile ONCE into bf16 and contracting it with all the queries of the list on tcgen05 replaces
// 16 gathers per (query, code) by 16 gathers per code + 128 x N x 128 MACs on the tensor pipe.
//
// Exactness.  The tensor-core value S' is only a FILTER.  With u = 2^-8 (bf16 round-to-nearest),
//     |S' - <q,r^>| <= (2u + u^2) |q| |r^| + (fp32 accumulation) <= 0.0085 |q| Rmax,   Rmax^2 = sum_m max_j |c_pq[m][j]|^2
// so with a per-query upper bound B_q of the K-th best key (taken from a LUT scan of the query's nearest lists,
// "phase A") every code with   key' <= B_q + |alpha| * 0.0085 |q| Rmax   is a *survivor*; survivors are
// re-evaluated with exactly the fp32 operations (and summation order) of the LUT kernel and kept when
// key <= B_q.  The final top-K is the same set with the same keys as the LUT engine returns
// (tests/test_ivf_gpu.py::test_ivfpq_tc_engine_matches_lut_engine).  Queries whose bound is missing (fewer
// than K codes in their nearest lists) or whose survivor buffer overflows are flagged and redone by the LUT kernel.
//
// sm_100a mapping (one persistent CTA per SM, 544 threads, all 512 TMEM columns, <= 219.5 KB of shared memory):
//   warps 0-7  decoders : (two groups of 4 warps, one per A buffer, so two tiles are decoded concurrently)
//                         code tile -> A operand [128 codes x K] bf16 in the no-swizzle K-major UMMA layout (16-byte
//                         sub-vector of sub-quantizer m = one core-matrix row), via a bf16 copy of the PQ codebooks in
//                         shared memory, + the admission-test K-step [-r_hi, -r_mid, -r_lo, 1, 1, 1, 0, 0]; they also stage
//                         the B operand (bf16 queries of the item, gathered by index, + [1, 1, 1, h_hi, h_mid, h_lo, 0, 0])
//                         and the per-column meta data, one item ahead.
//   warp  16   MMA      : K/16 + 1 x tcgen05.mma.kind::f16 (M=128, N=16..256, K=16) per tile into one of two 256-column
//                         TMEM accumulators; tcgen05.commit releases the A buffer and publishes the accumulator.
//   warps 8-15 epilogue : (two groups of 4 warps, one per accumulator) tcgen05.ld (32 lanes x 32 columns, double-buffered in
//                         registers); the accumulator holds D = S' + h - r, so "survives" is a clear sign bit: an AND tree
//                         over the 32 words decides "nothing passes" and only hits are expanded -> one shared-memory slot
//                         reservation per warp and tile -> the group's private survivor log in global memory (plain
//                         stores, nothing on the critical path waits for a global round trip).
//   Items are drawn from a global ticket counter in descending-cost order.
// Then: scatter_survivors_kernel groups the log by query, exact_eval_kernel recomputes the survivors' keys in fp32.
// Work item = (list, chunk of <= 256 of the queries probing it); items are laid out by a single-CTA plan kernel
// from the coarse result (count -> scan -> fill).
#pragma once
#include <cuda_bf16.h>

#include <cub/block/block_scan.cuh>

#include "kb2_gemm_tc.cuh"
#include "kb2_ivf.cuh"

namespace kb2 {
namespace pqtc {

constexpr int TM = 128;        // codes per tile (UMMA M)
constexpr int NQT = 256;       // queries per item (UMMA N max)
constexpr int THREADS = 544;          // warps 0-7 decoders (2 groups), 8-15 epilogue (2 groups), 16 MMA
constexpr int GROUP_THREADS = 128;
constexpr int MMA_WARP = 16;
constexpr int META_BYTES = 3 * NQT * 4;   // h | base | qidx
// geometry of one engine instance: G groups of 16 sub-quantizers of DSUB dimensions (K = 16 G DSUB).
// Instances: <1, 8> (m = 16, d = 128: BASELINE C3) and <3, 2> (m = 48, d = 96: BASELINE C5).
template <int G, int DSUB>
struct TcCfg {
    static constexpr int M = 16 * G;
    static constexpr int KD = M * DSUB;               // dimensions
    static_assert(KD % 16 == 0 && (DSUB == 2 || DSUB == 4 || DSUB == 8), "unsupported PQ geometry");
    static constexpr int XCHUNK = KD / 8;             // 16-byte chunk that carries the admission test (row term / column threshold)
    static constexpr int KSTEPS = KD / 16 + 1;        // the last K-step holds the test chunk + a zero chunk
    static constexpr int CHUNKS = 2 * KSTEPS;         // 16-byte chunks per operand row
    static constexpr int GRP_BYTES = CHUNKS * 128;    // one 8-row group of an operand: core matrices of 128 B (UMMA SBO)
    static constexpr int TAB_BYTES = M * 256 * DSUB * 2;   // bf16 codebooks
    static constexpr int A_BYTES = (TM / 8) * GRP_BYTES, B_BYTES = (NQT / 8) * GRP_BYTES;
    static constexpr int OFF_TAB = 0;
    static constexpr int OFF_A = TAB_BYTES;
    static constexpr int OFF_B = OFF_A + 2 * A_BYTES;
    static constexpr int OFF_META = OFF_B + B_BYTES;
    static constexpr int OFF_BAR = OFF_META + 2 * META_BYTES;
    static constexpr size_t SMEM_BYTES = OFF_BAR + 256 + 128 /*alignment slack*/;   // barriers 0..135, scheduler ring 144..239
    static_assert(SMEM_BYTES <= 227 * 1024, "filter kernel shared memory");
};
constexpr int KD = 128;                   // (phase-A kernels below are specific to the <1, 8> geometry)
constexpr float kErrCoef = 0.0085f;       // (2u + u^2) for bf16 operands + fp32 accumulation slack
constexpr float kAccCoef = 4e-5f;         // fp32 accumulation of the contraction incl. the threshold terms (x (|h| + max|r|))

struct Params {
    int metric;
    int nq, nprobe;
    const float* queries;          // [nq][128] fp32
    const __nv_bfloat16* qb16;     // [nq][128]
    const float* qnorm;            // [nq]
    // plan
    const int32_t* n_items;        // device scalar
    int32_t* ticket;               // optional work counter (zeroed before the launch): CTAs draw items from it in order, so a
                                   // CTA that got short items simply draws more (NULL: item = blockIdx.x + seq * gridDim.x)
    const int32_t* item_list;      // [items]
    const int32_t* item_q0;        // [items] first pair of the chunk
    const int32_t* item_nq;        // [items]
    const int32_t* pair_q;         // [pairs] query index, grouped by list
    const float* pair_base;        // [pairs] key base: L2 |q-c|^2, IP -<q,c>
    // per-query bound from phase A
    const float* bound;            // [nq] upper bound of the k_need-th best key (+inf: none -> the LUT kernel redoes the query)
    int k_need;
    float margin_coef;             // |alpha| * kErrCoef * Rmax  (multiplied by |q|)
    float rmax;                    // max over the index of the row term |t1| / 2 (L2; 0 for IP)
    // index
    const int64_t* list_off;
    const int32_t* list_len;
    const uint4* codes;            // [G][npad] 16 code bytes per group, rotated by pos % 16 (kb2_ivf.cuh)
    const uint4* codes_plain;      // [G][npad] un-rotated copy (byte b = sub-quantizer 16 g + b); used by the decode when DSUB < 8
    int64_t npad;
    const float* t1;               // [npad] (L2)
    const float* pqc;              // [16][256][8] fp32
    const uint4* pqc16;            // [16][256] x 8 bf16
    const uint8_t* bitset;
    const int32_t* rows;
    // output
    uint4* log;                    // [2*gridDim.x + 1][log_cap] survivors {query, position, key base bits, 0}: one log per
                                   // epilogue group, the last one shared by all for tiles that overflow the smem queue
    uint32_t* log_cnt;             // [2*gridDim.x] entries per private log; [2G] cursor of the shared log;
                                   // [2G + 1] = 1 when any log overflowed; [2G + 2..3] diagnostics
    uint32_t log_cap;              // entries per private log
    uint32_t shared_cap;           // entries of the shared log
    uint32_t* qflag;               // [nq] 1: redo this query with the LUT kernel
    unsigned long long* counters;  // [0] codes scanned (pairs x codes), [2] survivors re-evaluated, [3] flagged
};

__device__ __forceinline__ bool
mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must end the launch with an error, never hang the GPU
__device__ __forceinline__ void
mbar_wait_g(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 6000000000ll) __trap();
    }
}
__device__ __forceinline__ void
bar_sync_epi() {
    asm volatile("bar.sync 1, 256;" ::: "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous);
// LBO = distance between the two core matrices of one K=16 step (128 B), SBO = distance between 8-row groups (GRP_BYTES)
__device__ __forceinline__ uint64_t
make_desc_ns(uint32_t smem_addr, uint32_t grp_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(grp_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D=f32, A=B=bf16, K-major, M=128, N=n
__device__ __forceinline__ uint32_t
make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void
mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// x = hi + mid + lo with three bf16 terms (24 significant bits: exact for finite fp32 up to 2^-27 |x|); +-inf -> (+-inf, 0, 0)
__device__ __forceinline__ void
split3_bf16(float x, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = (uint32_t)__bfloat16_as_ushort(h);
    if (!(fabsf(x) < INFINITY)) { mid = 0u; lo = 0u; return; }
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    mid = (uint32_t)__bfloat16_as_ushort(m);
    lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(r2));
}

template <int METRIC, int G, int DSUB>
__global__ void __launch_bounds__(THREADS, 1)
ivfpq_tc_filter_kernel(Params p) {
    using C = TcCfg<G, DSUB>;
    constexpr int KD = C::KD, XCHUNK = C::XCHUNK, KSTEPS = C::KSTEPS, GRP_BYTES = C::GRP_BYTES, TAB_BYTES = C::TAB_BYTES;
    constexpr int A_BYTES = C::A_BYTES, OFF_TAB = C::OFF_TAB, OFF_A = C::OFF_A, OFF_B = C::OFF_B, OFF_META = C::OFF_META,
                  OFF_BAR = C::OFF_BAR;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = tc::smem_u32(smem_dyn);
    const uint32_t base = (raw + 127u) & ~127u;
    unsigned char* sm = smem_dyn + (base - raw);
    const uint32_t bars = base + OFF_BAR;
    // barriers: a_full[2] a_empty[2] acc_full[2] acc_empty[2] b_full b_free meta_full[2] meta_free[2]
    auto bar_a_full = [&](int i) { return bars + 8u * i; };
    auto bar_a_empty = [&](int i) { return bars + 8u * (2 + i); };
    auto bar_acc_full = [&](int i) { return bars + 8u * (4 + i); };
    auto bar_acc_empty = [&](int i) { return bars + 8u * (6 + i); };
    const uint32_t bar_b_full = bars + 8u * 8, bar_b_free = bars + 8u * 9;
    auto bar_meta_full = [&](int i) { return bars + 8u * (10 + i); };
    auto bar_meta_free = [&](int i) { return bars + 8u * (12 + i); };
    const uint32_t tmem_slot = bars + 8u * 14;
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(sm + OFF_BAR + 8 * 14);
    uint32_t* qcnt = (uint32_t*)(sm + OFF_BAR + 8 * 15);   // survivor counters: [epilogue group][own tile parity]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = *p.n_items;

    // ---- item sequence of this CTA.  The three roles walk the same sequence independently (at most ~3 items apart), so the
    // seq-th draw is published through a small ring in shared memory: whoever needs it first claims the slot, takes a ticket
    // from the global counter and publishes it; the others read it.
    constexpr int SCHED_R = 8;
    int* sch_claim = (int*)(sm + OFF_BAR + 144);
    int* sch_item = sch_claim + SCHED_R;
    volatile int* sch_ready = (volatile int*)(sch_item + SCHED_R);
    if (threadIdx.x < SCHED_R) {
        sch_claim[threadIdx.x] = (int)threadIdx.x - SCHED_R;
        sch_ready[threadIdx.x] = -1;
    }
    auto item_at = [&](int seq) -> int {   // warp-uniform call
        if (!p.ticket) return (int)blockIdx.x + seq * (int)gridDim.x;
        int v = 0;
        if (lane == 0) {
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
            v = ((volatile int*)sch_item)[sl];
        }
        return __shfl_sync(0xffffffffu, v, 0);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(bar_a_full(i), GROUP_THREADS);
            tc::mbar_init(bar_a_empty(i), 1);
            tc::mbar_init(bar_acc_full(i), 1);
            tc::mbar_init(bar_acc_empty(i), GROUP_THREADS);
            tc::mbar_init(bar_meta_full(i), 2 * GROUP_THREADS);
            tc::mbar_init(bar_meta_free(i), 2 * GROUP_THREADS);
        }
        tc::mbar_init(bar_b_full, 2 * GROUP_THREADS);
        tc::mbar_init(bar_b_free, 1);
        qcnt[0] = qcnt[1] = qcnt[2] = qcnt[3] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // bf16 codebooks -> shared memory
    {
        uint4* tab = (uint4*)(sm + OFF_TAB);
        for (int i = threadIdx.x; i < TAB_BYTES / 16; i += THREADS) tab[i] = __ldg(p.pqc16 + i);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const float inv_alpha = (METRIC == KB2_METRIC_L2) ? 0.5f : 1.f;

    if (warp < 8) {
        // =========================== decoders: two groups of 4 warps, group d owns A buffer d (tiles g % 2 == d) =========
        const int dt = threadIdx.x;    // 0..255
        const int dg = dt >> 7;        // group
        const int tid = dt & 127;      // code row inside the tile
        const uint4* tab = (const uint4*)(sm + OFF_TAB);
        // per-column thresholds of one item -> meta buffer (it & 1), one column per decoder thread; written one item
        // AHEAD of its use so that the dependent global loads (pair -> bound, norm) stay off the critical path
        auto write_meta = [&](int item, int it) {
            const int par = it & 1;
            mbar_wait_g(bar_meta_free(par), (((uint32_t)it >> 1) & 1u) ^ 1u);
            const int q0 = p.item_q0[item];
            const int nqi = p.item_nq[item];
            float* m_h = (float*)(sm + OFF_META + par * META_BYTES);
            float* m_base = m_h + NQT;
            int* m_q = (int*)(m_base + NQT);
            const int j = dt;
            float h = -INFINITY, bs = 0.f;
            int q = -1;
            if (j < nqi) {
                q = p.pair_q[q0 + j];
                bs = p.pair_base[q0 + j];
                const float bnd = p.bound[q];
                if (!(bnd < INFINITY)) {
                    p.qflag[q] = 1u;   // no bound: the LUT kernel redoes this query
                    if (p.counters) atomicAdd(p.counters + 6, 1ull << 32);
                } else {
                    // pass  <=>  S' + h >= r  (S' bf16 contraction, r the row term); the margin covers the bf16 operand error,
                    // the extra term the fp32 accumulation of the K=144 contraction that now carries h and r as well
                    const float margin = p.margin_coef * p.qnorm[q] * 1.01f + 1e-30f;
                    const float h0 = (bnd + margin - bs) * inv_alpha;
                    h = h0 + kAccCoef * (fabsf(h0) + p.rmax) + 1e-30f;
                }
            }
            m_h[j] = h;
            m_base[j] = bs;
            m_q[j] = q;
            tc::mbar_arrive(bar_meta_full(par));
        };
        uint32_t g0 = 0;   // global tile counter at the start of the item
        int it = 0;
        int item = item_at(0);
        if (item < n_items) write_meta(item, 0);
        for (; item < n_items; it++) {
            const int item_next = item_at(it + 1);
            const int l = p.item_list[item];
            const int nqi = p.item_nq[item];
            const int nmma = (nqi + 15) & ~15;
            const int len = p.list_len[l];
            const int64_t off = p.list_off[l];
            const int ntiles = (len + TM - 1) / TM;
            const int par = it & 1;
            const int t_first = (int)((dg - (int)(g0 & 1u)) & 1);   // this group's first tile of the item
            uint4 w_next[G];
#pragma unroll
            for (int g = 0; g < G; g++) w_next[g] = make_uint4(0, 0, 0, 0);
            float t_next = 0.f;
            const uint4* code_src = (DSUB == 8) ? p.codes : p.codes_plain;
            if (t_first < ntiles && off + (int64_t)t_first * TM + tid < p.npad) {
#pragma unroll
                for (int g = 0; g < G; g++) w_next[g] = ldg_stream_u4(code_src + (int64_t)g * p.npad + off + (int64_t)t_first * TM + tid);
                if (METRIC == KB2_METRIC_L2) t_next = __ldg(p.t1 + off + (int64_t)t_first * TM + tid);
            }
            auto decode_tile = [&](int t) {
                const uint32_t g = g0 + (uint32_t)t;      // g & 1 == dg
                uint4 w[G];
#pragma unroll
                for (int gg = 0; gg < G; gg++) w[gg] = w_next[gg];
                const float tv = t_next;
                {
                    const int64_t pn = off + (int64_t)(t + 2) * TM + tid;
                    if (t + 2 < ntiles && pn < p.npad) {
#pragma unroll
                        for (int gg = 0; gg < G; gg++) w_next[gg] = ldg_stream_u4(code_src + (int64_t)gg * p.npad + pn);
                        if (METRIC == KB2_METRIC_L2) t_next = __ldg(p.t1 + pn);
                    }
                }
                // row term of the admission test, negated, as three bf16 terms; rows past the end of the list never pass
                float r = INFINITY;
                if (t * TM + tid < len) r = (METRIC == KB2_METRIC_L2) ? 0.5f * tv : 0.f;
                uint32_t rh, rm, rl;
                split3_bf16(-r, rh, rm, rl);
                mbar_wait_g(bar_a_empty(dg), ((g >> 1) & 1u) ^ 1u);
                unsigned char* A = sm + OFF_A + dg * A_BYTES + (tid >> 3) * GRP_BYTES + (tid & 7) * 16;
                if constexpr (DSUB == 8) {
                    // chunk = one sub-quantizer (8 bf16): 16 gathers per group through the rotated code bytes.  The table is
                    // laid out [code value][sub-quantizer] (16 B entries): the 8 lanes of a quarter-warp hold 8 consecutive
                    // sub-quantizers, i.e. 8 different 16-byte bank groups whatever their code values => every LDS.128 is
                    // conflict-free (a [sub-quantizer][code value] table costs ~3x the wavefronts with random codes).
#pragma unroll
                    for (int gg = 0; gg < G; gg++) {
                        const uint32_t ww[4] = {w[gg].x, w[gg].y, w[gg].z, w[gg].w};
#pragma unroll
                        for (int h0 = 0; h0 < 16; h0 += 8) {     // 8 gathers in flight, then 8 stores
                            uint4 v[8];
#pragma unroll
                            for (int s = 0; s < 8; s++) {
                                const uint32_t byte = (ww[(h0 + s) >> 2] >> (8 * ((h0 + s) & 3))) & 255u;
                                v[s] = tab[byte * (16 * G) + gg * 16 + ((h0 + s + tid) & 15)];   // pos % 16 == tid % 16
                            }
#pragma unroll
                            for (int s = 0; s < 8; s++) *reinterpret_cast<uint4*>(A + (gg * 16 + ((h0 + s + tid) & 15)) * 128) = v[s];
                        }
                    }
                } else {
                    // chunk = 8 / DSUB consecutive sub-quantizers, assembled from the un-rotated code bytes (static indices)
                    constexpr int SPC = 8 / DSUB;              // sub-quantizers per 16-byte chunk
                    constexpr int WPS = DSUB / 2;              // 32-bit words per sub-quantizer entry
                    const uint32_t* tab32 = reinterpret_cast<const uint32_t*>(tab);
#pragma unroll
                    for (int c0 = 0; c0 < XCHUNK; c0 += 4) {   // 4 chunks (16 words) in flight
                        uint32_t v[4][4];
#pragma unroll
                        for (int cc = 0; cc < 4; cc++) {
#pragma unroll
                            for (int j = 0; j < SPC; j++) {
                                const int m = (c0 + cc) * SPC + j;
                                const uint4 wg = w[m >> 4];
                                const int b = m & 15;
                                const uint32_t word = (b < 4) ? wg.x : (b < 8) ? wg.y : (b < 12) ? wg.z : wg.w;
                                const uint32_t byte = (word >> (8 * (b & 3))) & 255u;
#pragma unroll
                                for (int x = 0; x < WPS; x++) v[cc][j * WPS + x] = tab32[(m * 256 + byte) * WPS + x];
                            }
                        }
#pragma unroll
                        for (int cc = 0; cc < 4; cc++)
                            *reinterpret_cast<uint4*>(A + (c0 + cc) * 128) = make_uint4(v[cc][0], v[cc][1], v[cc][2], v[cc][3]);
                    }
                }
                // test chunk: [-r_hi, -r_mid, -r_lo, 1, 1, 1, 0, 0] (bf16 1.0 = 0x3F80); then a zero chunk
                *reinterpret_cast<uint4*>(A + XCHUNK * 128) = make_uint4(rh | (rm << 16), rl | (0x3F80u << 16), 0x3F803F80u, 0u);
                *reinterpret_cast<uint4*>(A + (XCHUNK + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
                tc::fence_proxy_async();
                tc::mbar_arrive(bar_a_full(dg));
            };
            // the first tile of each group only needs a free A buffer: decode it while the tensor pipe still works on
            // the previous item, then stage the B operand (which must wait for that item's last MMA)
            if (t_first < ntiles) decode_tile(t_first);
            // ---- B operand: the item's queries (bf16) gathered by index with cp.async, K-major no-swizzle layout,
            //      plus the threshold chunk [1, 1, 1, h_hi, h_mid, h_lo, 0, 0] of every column
            asm volatile("bar.sync 2, 256;" ::: "memory");          // meta[par] (thresholds, query indices) written by all decoders
            mbar_wait_g(bar_b_free, ((uint32_t)it & 1u) ^ 1u);
            {
                const float* m_h = (const float*)(sm + OFF_META + par * META_BYTES);
                const int* m_q = (const int*)(m_h + 2 * NQT);
                const uint32_t Bs = base + OFF_B;
                unsigned char* B = sm + OFF_B;
                const int kc = tid >> 3;        // 16-byte chunk along K (0..15)
                const int rsub = tid & 7;
                for (int blk = dg; blk < nmma / 8 && kc < XCHUNK; blk += 2) {
                    const int row = blk * 8 + rsub;
                    const int q = m_q[row];
                    const uint32_t dst = (uint32_t)(blk * GRP_BYTES + kc * 128 + rsub * 16);
                    if (q >= 0) {
                        const void* src = reinterpret_cast<const uint4*>(p.qb16 + (int64_t)q * KD) + kc;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(Bs + dst), "l"(src) : "memory");
                    } else {
                        *reinterpret_cast<uint4*>(B + dst) = make_uint4(0, 0, 0, 0);
                    }
                }
                if (dt < nmma) {   // one column per decoder thread
                    uint32_t hh, hm, hl;
                    split3_bf16(m_h[dt], hh, hm, hl);
                    unsigned char* Bc = B + (dt >> 3) * GRP_BYTES + (dt & 7) * 16;
                    *reinterpret_cast<uint4*>(Bc + XCHUNK * 128) = make_uint4(0x3F803F80u, 0x3F80u | (hh << 16), hm | (hl << 16), 0u);
                    *reinterpret_cast<uint4*>(Bc + (XCHUNK + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bar_b_full);
            // thresholds of the NEXT item (other meta buffer)
            if (item_next < n_items) write_meta(item_next, it + 1);
            // ---- this group's remaining tiles
            for (int t = t_first + 2; t < ntiles; t += 2) decode_tile(t);
            g0 += (uint32_t)ntiles;
            item = item_next;
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ===========================
        uint32_t g = 0;
        int it = 0;
        for (int item = item_at(0); item < n_items; item = item_at(++it)) {
            const int l = p.item_list[item];
            const int nqi = p.item_nq[item];
            const int nmma = (nqi + 15) & ~15;
            const int ntiles = (p.list_len[l] + TM - 1) / TM;
            const uint32_t idesc = make_idesc_bf16(nmma);
            mbar_wait_g(bar_b_full, (uint32_t)it & 1u);
            for (int t = 0; t < ntiles; t++, g++) {
                const int buf = g & 1;
                mbar_wait_g(bar_a_full(buf), (g >> 1) & 1u);
                mbar_wait_g(bar_acc_empty(buf), ((g >> 1) & 1u) ^ 1u);
                tc::tc_fence_after();
                if (lane == 0) {
                    const uint32_t a0 = base + OFF_A + buf * A_BYTES;
                    const uint32_t b0 = base + OFF_B;
                    const uint32_t d = tmem_base + (uint32_t)buf * 256u;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++)
                        mma_bf16(d, make_desc_ns(a0 + ks * 256, GRP_BYTES), make_desc_ns(b0 + ks * 256, GRP_BYTES), idesc, ks > 0 ? 1u : 0u);
                    tc::tc_commit(bar_a_empty(buf));
                    tc::tc_commit(bar_acc_full(buf));
                }
                __syncwarp();
            }
            if (lane == 0) tc::tc_commit(bar_b_free);
            __syncwarp();
        }
    } else {
        // =========================== epilogue: two groups of 4 warps, group e owns accumulator e (tiles g % 2 == e) ======
        // The accumulator holds D = S' + h_col - r_row: a (code, query) pair survives iff D >= 0, i.e. iff its sign bit is
        // clear.  "No survivor in these 32 columns" is an AND over the 32 words (LOP3 tree); survivors are rare (~0.1 %), so
        // the exact mask is built only on a hit and the entries go straight to the group's survivor log in global memory
        // (slot range reserved with one shared-memory atomic per warp and tile; plain stores, nothing waits for them).
        const int et = threadIdx.x - 256;        // 0..255
        const int eg = et >> 7;                  // group
        const int e = et & 127;                  // thread inside the group
        const int we = warp & 3;                 // TMEM lane quarter (== warp % 4)
        const int row = we * 32 + lane;          // code row inside the tile
        const uint32_t n_logs = 2u * gridDim.x;  // private logs (log n_logs, the former shared one, stays empty)
        uint4* my_log = p.log + (size_t)(2 * blockIdx.x + eg) * p.log_cap;
        uint32_t* my_cursor = qcnt + eg;
        bool log_over = false;
        unsigned long long n_codes = 0;
        uint32_t g0 = 0;
        int it = 0;
#define KB2_TMEM_LD32(V, TADDR)                                                                                          \
    asm volatile(                                                                                                        \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                        \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                                        \
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                       \
        : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), "=r"(V[8]),    \
          "=r"(V[9]), "=r"(V[10]), "=r"(V[11]), "=r"(V[12]), "=r"(V[13]), "=r"(V[14]), "=r"(V[15]), "=r"(V[16]),         \
          "=r"(V[17]), "=r"(V[18]), "=r"(V[19]), "=r"(V[20]), "=r"(V[21]), "=r"(V[22]), "=r"(V[23]), "=r"(V[24]),        \
          "=r"(V[25]), "=r"(V[26]), "=r"(V[27]), "=r"(V[28]), "=r"(V[29]), "=r"(V[30]), "=r"(V[31])                      \
        : "r"(TADDR))
        for (int item = item_at(0); item < n_items; item = item_at(++it)) {
            const int l = p.item_list[item];
            const int nqi = p.item_nq[item];
            const int nmma = (nqi + 15) & ~15;
            const int nch = (nmma + 31) >> 5;    // 32-column chunks; a 16-column tail reads 16 stale columns (masked below)
            const int len = p.list_len[l];
            const int64_t off = p.list_off[l];
            const int ntiles = (len + TM - 1) / TM;
            const int par = it & 1;
            mbar_wait_g(bar_meta_full(par), ((uint32_t)it >> 1) & 1u);
            const float* m_base = (const float*)(sm + OFF_META + par * META_BYTES) + NQT;
            const int* m_q = (const int*)(m_base + NQT);
            if (et == 0) n_codes += (unsigned long long)len * (unsigned long long)nqi;
            const int t_first = (int)((eg - (int)(g0 & 1u)) & 1);
            const uint32_t tail_mask = (nmma & 31) ? 0x0000ffffu : 0xffffffffu;   // valid columns of the last chunk
            for (int t = t_first; t < ntiles; t += 2) {
                const uint32_t g = g0 + (uint32_t)t;   // g & 1 == eg
                mbar_wait_g(bar_acc_full(eg), (g >> 1) & 1u);
                tc::tc_fence_after();
                const int rel = t * TM + row;
                const uint32_t taddr0 = tmem_base + ((uint32_t)(we * 32) << 16) + (uint32_t)(eg * 256);
                uint32_t masks[8];
                uint32_t va[32], vb[32];
                // D >= 0  <=>  sign bit clear.  Two-level test with short dependency chains (the serial OR-chain of a naive
                // mask build was the epilogue's critical path): 8 independent 4-way ANDs, a 3-deep AND tree over them for the
                // common "nothing passes" exit, and on a hit only the groups whose AND has a clear sign bit are expanded.
                auto scan_chunk = [&](const uint32_t (&v)[32]) -> uint32_t {
                    uint32_t g[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) g[i] = (v[4 * i] & v[4 * i + 1]) & (v[4 * i + 2] & v[4 * i + 3]);
                    const uint32_t all = ((g[0] & g[1]) & (g[2] & g[3])) & ((g[4] & g[5]) & (g[6] & g[7]));
                    if ((int32_t)all < 0) return 0u;   // all 32 sign bits set: nothing passes
                    uint32_t m = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if ((int32_t)g[i] >= 0) {
                            const uint32_t b0 = (~v[4 * i]) >> 31, b1 = (~v[4 * i + 1]) >> 31;
                            const uint32_t b2 = (~v[4 * i + 2]) >> 31, b3 = (~v[4 * i + 3]) >> 31;
                            m |= ((b0 | (b1 << 1)) | ((b2 << 2) | (b3 << 3))) << (4 * i);
                        }
                    }
                    return m;
                };
                // software pipeline over the chunks: the TMEM load of chunk ci+1 is in flight while chunk ci is tested
                KB2_TMEM_LD32(va, taddr0);
#pragma unroll
                for (int ci = 0; ci < 8; ci++) {
                    masks[ci] = 0;
                    if (ci < nch) {   // (the other epilogue group works on the other accumulator meanwhile)
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (ci + 1 < nch) {
                            if (ci & 1) { KB2_TMEM_LD32(va, taddr0 + (uint32_t)((ci + 1) * 32)); }
                            else { KB2_TMEM_LD32(vb, taddr0 + (uint32_t)((ci + 1) * 32)); }
                        } else {
                            // everything of this tile is in registers: hand the accumulator back before the last test
                            tc::tc_fence_before();
                            tc::mbar_arrive(bar_acc_empty(eg));
                        }
                        masks[ci] = (ci & 1) ? scan_chunk(vb) : scan_chunk(va);
                        if (ci == nch - 1) masks[ci] &= tail_mask;
                    }
                }
                // ---- survivors of this thread's row -> the group's log
                uint32_t total = 0;
#pragma unroll
                for (int ci = 0; ci < 8; ci++) total += __popc(masks[ci]);
                if (__any_sync(0xffffffffu, total != 0u)) {
                    uint32_t incl = total;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t tv = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += tv;
                    }
                    uint32_t wbase = 0;
                    if (lane == 31) wbase = atomicAdd(my_cursor, incl);
                    wbase = __shfl_sync(0xffffffffu, wbase, 31);
                    uint32_t slot = wbase + incl - total;
#pragma unroll
                    for (int ci = 0; ci < 8; ci++) {
                        uint32_t m = masks[ci];
                        while (m) {
                            const int u = __ffs(m) - 1;
                            m &= m - 1;
                            const int col = ci * 32 + u;
                            if (slot < p.log_cap) {
                                uint4 o;
                                o.x = (uint32_t)m_q[col];
                                o.y = (uint32_t)(off + rel);
                                o.z = __float_as_uint(m_base[col]);
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
            tc::mbar_arrive(bar_meta_free(par));
            g0 += (uint32_t)ntiles;
        }
#undef KB2_TMEM_LD32
        if (log_over) p.log_cnt[n_logs + 1] = 1u;
        if (eg == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
        if (e == 0) {
            const uint32_t n = min(*my_cursor, p.log_cap);
            p.log_cnt[2 * blockIdx.x + eg] = n;
            if (p.counters) atomicAdd(p.counters + 2, (unsigned long long)n);
        }
        if (et == 0 && p.counters) atomicAdd(p.counters, n_codes);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------- plan: (query, probe) pairs grouped by list
__global__ void
count_pairs_kernel(const int64_t* __restrict__ probe_ids, int64_t npairs, const int32_t* __restrict__ list_len,
                   int32_t* __restrict__ lcount) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int64_t l = probe_ids[i];
    if (l >= 0 && list_len[l] > 0) atomicAdd(lcount + l, 1);
}

// one CTA: exclusive scans over the lists -> first pair of each list, item table (list, query chunk)
__global__ void __launch_bounds__(1024)
plan_kernel(const int32_t* __restrict__ lcount, int nlist, int32_t* __restrict__ lstart, int32_t* __restrict__ item_list,
            int32_t* __restrict__ item_q0, int32_t* __restrict__ item_nq, int32_t* __restrict__ n_items) {
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp_a, tmp_b;
    __shared__ int carry_a, carry_b;
    if (threadIdx.x == 0) carry_a = carry_b = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nlist; b0 += 1024) {
        const int l = b0 + threadIdx.x;
        const int c = l < nlist ? lcount[l] : 0;
        const int nch = (c + NQT - 1) / NQT;
        int ex_a, ex_b, tot_a, tot_b;
        Scan(tmp_a).ExclusiveSum(c, ex_a, tot_a);
        Scan(tmp_b).ExclusiveSum(nch, ex_b, tot_b);
        const int ca = carry_a, cb = carry_b;
        if (l < nlist) {
            lstart[l] = ca + ex_a;
            if (nch > 0) {
                // even chunks, multiples of 16 queries (the UMMA N granularity); full chunks when rounding would leave the
                // last one empty (an item without queries would be an N = 0 MMA)
                int per = ((c + nch - 1) / nch + 15) & ~15;
                if ((nch - 1) * per >= c) per = NQT;
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

// ---- static load balancing of the persistent kernels.  CTA b processes items b, b + G, b + 2G, ...; in list order the per-CTA
// totals differ by +-30 % (list length x queries per list has a heavy tail), and the launch lasts as long as its slowest CTA.
// Items are therefore sorted by descending cost estimate and dealt out in snake order (round r forwards, round r+1
// backwards), the classic longest-processing-time deal.
__global__ void
item_cost_kernel(const int32_t* __restrict__ n_items, const int32_t* __restrict__ item_list, const int32_t* __restrict__ item_nq,
                 const int32_t* __restrict__ list_len, int64_t max_items, int tile_cost, int col_cost, uint32_t* __restrict__ key,
                 int32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_items) return;
    uint32_t k = 0xffffu;   // unused slots sort to the end; 16-bit keys = two radix passes
    if (i < *n_items) {
        const long long tiles = (list_len[item_list[i]] + TM - 1) / TM;
        const long long c = (tiles * (tile_cost + (long long)col_cost * ((item_nq[i] + 15) & ~15))) >> 5;
        k = 0xfffeu - (uint32_t)min(c, 0xfff0ll);   // ascending key = descending cost
    }
    key[i] = k;
    idx[i] = (int32_t)i;
}
__global__ void
deal_items_kernel(const int32_t* __restrict__ n_items, const int32_t* __restrict__ sorted_idx, int G, const int32_t* __restrict__ in_list,
                  const int32_t* __restrict__ in_q0, const int32_t* __restrict__ in_nq, int32_t* __restrict__ out_list,
                  int32_t* __restrict__ out_q0, int32_t* __restrict__ out_nq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // rank by descending cost
    const int n = *n_items;
    if (j >= n) return;
    int dst = j;   // G == 0: plain descending-cost order (drawn dynamically through Params::ticket)
    if (G > 0) {
        const int r = j / G, b = j % G;
        const bool full_round = (r + 1) * G <= n;
        dst = r * G + (((r & 1) && full_round) ? (G - 1 - b) : b);
    }
    const int src = sorted_idx[j];
    out_list[dst] = in_list[src];
    out_q0[dst] = in_q0[src];
    out_nq[dst] = in_nq[src];
}

__global__ void
fill_pairs_kernel(const int64_t* __restrict__ probe_ids, const float* __restrict__ probe_dis, int64_t npairs, int nprobe,
                  int metric, const int32_t* __restrict__ list_len, const int32_t* __restrict__ lstart,
                  int32_t* __restrict__ lcursor, int32_t* __restrict__ pair_q, float* __restrict__ pair_base) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int64_t l = probe_ids[i];
    if (l < 0 || list_len[l] <= 0) return;
    const int slot = lstart[l] + atomicAdd(lcursor + l, 1);
    pair_q[slot] = (int32_t)(i / nprobe);
    const float dv = probe_dis[i];
    pair_base[slot] = (metric == KB2_METRIC_L2) ? dv : -dv;
}

// bf16 copy + norm of the queries (warp per query, d % 4 == 0)
__global__ void
prepare_queries_kernel(const float* __restrict__ q, int64_t nq, int d, __nv_bfloat16* __restrict__ qb16, float* __restrict__ qnorm) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nq) return;
    float s = 0.f;
    for (int j = lane * 4; j < d; j += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(q + w * d + j));
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(qb16 + w * d + j) = o;
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = warp_sum(s);
    if (lane == 0) qnorm[w] = sqrtf(s) * 1.0001f;
}

// bf16 copy of the PQ codebooks + max_j |c[m][j]|^2 per sub-quantizer (grid = M, block = 256).  Layout: dsub == 8:
// [code value j][sub-quantizer m] (conflict-free gathers in the filter kernel's decode); dsub < 8: [m][j].
__global__ void __launch_bounds__(256)
prepare_tables_kernel(const float* __restrict__ pqc, int dsub, __nv_bfloat16* __restrict__ pqc16, float* __restrict__ maxn2) {
    const int m = blockIdx.x, j = threadIdx.x;
    const int M = gridDim.x;
    const float* c = pqc + ((size_t)m * 256 + j) * dsub;
    float n2 = 0.f;
    __nv_bfloat16* o = pqc16 + (dsub == 8 ? ((size_t)j * M + m) : ((size_t)m * 256 + j)) * dsub;
    for (int t = 0; t < dsub; t++) {
        n2 = fmaf(c[t], c[t], n2);
        o[t] = __float2bfloat16_rn(c[t]);
    }
    __shared__ float red[256];
    red[j] = n2;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (j < s) red[j] = fmaxf(red[j], red[j + s]);
        __syncthreads();
    }
    if (j == 0) maxn2[m] = red[0];
}
// un-rotated copy of the code bytes: plain[g][pos] byte b = sub-quantizer 16 g + b   (rotated: byte s = sub-quantizer (s + pos) % 16)
__global__ void
unrotate_codes_kernel(const uint8_t* __restrict__ rot, int64_t total_words /* G * npad */, int64_t npad, uint8_t* __restrict__ plain) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_words * 16) return;
    const int s = (int)(t & 15);
    const int64_t word = t >> 4;
    const int64_t pos = word % npad;
    plain[word * 16 + ((s + (int)(pos & 15)) & 15)] = rot[t];
}

// survivors of all CTA logs -> per-query rows (thread per log entry; grid = (x, number of logs))
__global__ void
scatter_survivors_kernel(const uint4* __restrict__ log, const uint32_t* __restrict__ log_cnt, uint32_t log_cap,
                         uint32_t shared_cap, uint64_t* __restrict__ cand, uint32_t* __restrict__ cand_cnt, int cap, uint32_t* __restrict__ qflag,
                         unsigned long long* __restrict__ counters) {
    const uint32_t n = min(log_cnt[blockIdx.y], blockIdx.y + 1 == gridDim.y ? shared_cap : log_cap);   // last log = shared one
    const uint4* src = log + (size_t)blockIdx.y * log_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 e = src[i];
        const uint32_t slot = atomicAdd(cand_cnt + e.x, 1u);
        if (slot < (uint32_t)cap) cand[(int64_t)e.x * cap + slot] = ((uint64_t)e.z << 32) | e.y;   // (key base bits, position)
        else {
            qflag[e.x] = 1u;
            if (counters) atomicAdd(counters + 6, 1ull);
        }
    }
}

// Per-query ADC tables for the whole batch:  lut[q][j*16 + m] = scale * <q_m, c_pq[m][j]>  (scale -2 for L2, -1 for IP),
// each entry the same 8-term fma chain the LUT kernel uses when it builds its table itself, so every consumer
// (phase A, the exact re-evaluation, the LUT kernel's copy-in mode) sees bit-identical values.
// grid = number of SMs, block = 256 (thread = code value j, its 16 sub-vectors live in registers).
template <int METRIC>
__global__ void __launch_bounds__(256, 1)
lut_build_kernel(const float* __restrict__ queries, int64_t nq, const int32_t* __restrict__ qlist, const uint32_t* __restrict__ qcount,
                 const float* __restrict__ pqc, float* __restrict__ lut) {
    // qlist != NULL: table i belongs to query qlist[i], i < *qcount (the queries this rank runs phase A for)
    __shared__ __align__(16) float s_q[2][KD];
    const int j = threadIdx.x;
    const float scale = (METRIC == KB2_METRIC_L2) ? -2.f : -1.f;
    if (qlist) nq = (int64_t)*qcount;
    // thread j keeps the 16 sub-vectors c_pq[.][j] (128 floats) in registers for all the queries of this CTA
    float4 c[32];
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const float* cp = pqc + ((size_t)m * 256 + j) * 8;
        c[2 * m] = __ldg(reinterpret_cast<const float4*>(cp));
        c[2 * m + 1] = __ldg(reinterpret_cast<const float4*>(cp + 4));
    }
    const int64_t per = (nq + gridDim.x - 1) / gridDim.x;
    const int64_t q_beg = (int64_t)blockIdx.x * per, q_end = min(nq, q_beg + per);
    auto qrow = [&](int64_t i) { return queries + (qlist ? (int64_t)qlist[i] : i) * KD; };
    if (q_beg < q_end && threadIdx.x < KD) s_q[0][threadIdx.x] = qrow(q_beg)[threadIdx.x];
    __syncthreads();
    for (int64_t q = q_beg; q < q_end; q++) {
        const int cur = (int)((q - q_beg) & 1);
        if (q + 1 < q_end && threadIdx.x < KD) s_q[cur ^ 1][threadIdx.x] = qrow(q + 1)[threadIdx.x];
        float* dst = lut + q * 4096 + j * 16;
#pragma unroll
        for (int m4 = 0; m4 < 16; m4 += 4) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int m = m4 + u;
                const float4 qa = *reinterpret_cast<const float4*>(&s_q[cur][m * 8]);
                const float4 qb = *reinterpret_cast<const float4*>(&s_q[cur][m * 8 + 4]);
                float a = 0.f;
                a = fmaf(qa.x, c[2 * m].x, a); a = fmaf(qa.y, c[2 * m].y, a); a = fmaf(qa.z, c[2 * m].z, a); a = fmaf(qa.w, c[2 * m].w, a);
                a = fmaf(qb.x, c[2 * m + 1].x, a); a = fmaf(qb.y, c[2 * m + 1].y, a); a = fmaf(qb.z, c[2 * m + 1].z, a); a = fmaf(qb.w, c[2 * m + 1].w, a);
                o[u] = a * scale;
            }
            *reinterpret_cast<float4*>(dst + m4) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
    }
}

// exact fp32 keys of the survivors: the LUT kernel's own values in its own order — every table entry is the 8-term fma
// chain of <q_m, c_pq[m][code]> times the scale, rounded once (never contracted into the sum), and the 16 entries are
// added into two interleaved accumulators over the stored byte order, key = base + (acc0 + acc1) — so both engines
// produce bit-identical keys.  The entries are recomputed from the fp32 codebook (L1/L2 resident) instead of read from
// a per-query table: no [nq][4096] table has to exist for the queries whose phase A ran on another rank.
// grid = nq, block = 128.  Survivors above the bound are dropped; with k_trim > 0 the row is cut to its k_trim best and compacted.
template <int METRIC, int G, int DSUB>
__global__ void __launch_bounds__(128)
exact_eval_kernel(const float* __restrict__ queries, const float* __restrict__ pqc, const float* __restrict__ lut,
                  const float* __restrict__ bound_of,
                  const uint4* __restrict__ codes, int64_t npad, const float* __restrict__ t1, const uint8_t* __restrict__ bitset,
                  const int32_t* __restrict__ rows, uint64_t* __restrict__ cand, uint32_t* __restrict__ cand_cnt, int cap,
                  uint32_t* __restrict__ qflag, const uint32_t* __restrict__ log_over, int k_trim = 0) {
    constexpr int KDIM = 16 * G * DSUB;
    __shared__ __align__(16) float s_q[KDIM];
    __shared__ uint32_t s_h[256];
    __shared__ float s_red[8];
    __shared__ uint32_t s_ctl[2];   // [0] crossing bin, [1] entries kept
    const int64_t q = blockIdx.x;
    if (*log_over) {
        if (threadIdx.x == 0) qflag[q] = 1u;
        return;
    }
    if (qflag[q]) return;
    const uint32_t n = min(cand_cnt[q], (uint32_t)cap);
    if (n == 0) return;
    // lut != NULL (<1, 8> geometry, single GPU): the batch's tables [nq][256][16] from lut_build_kernel hold exactly these
    // entries; otherwise they are recomputed from the codebook
    const bool use_lut = (G == 1 && DSUB == 8) && lut != nullptr;
    if (!use_lut) {
        for (int j = threadIdx.x; j < KDIM; j += 128) s_q[j] = queries[q * KDIM + j];
        __syncthreads();
    }
    const float* lq = lut + q * 4096;
    const float bound = bound_of[q];
    const float scale = (METRIC == KB2_METRIC_L2) ? -2.f : -1.f;
    uint64_t* row = cand + q * cap;
    // exact key of one logged survivor; returns the packed (key, position) entry, or kEmpty when it is above the bound / filtered
    auto eval = [&](uint64_t ent, float& key_out) -> uint64_t {
        const uint32_t pos = (uint32_t)ent;
        const float base = __uint_as_float((uint32_t)(ent >> 32));
        float acc0 = (METRIC == KB2_METRIC_L2) ? __ldg(t1 + pos) : 0.f, acc1 = 0.f;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const uint4 w = __ldg(codes + (int64_t)g * npad + pos);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const uint32_t byte = (ww[s >> 2] >> (8 * (s & 3))) & 255u;
                const uint32_t m = g * 16 + ((pos + s) & 15u);
                float v;
                if (use_lut) {
                    v = __ldg(lq + byte * 16 + m);
                } else {
                    const float* c = pqc + ((size_t)m * 256 + byte) * DSUB;
                    const float* qs = &s_q[m * DSUB];
                    float a = 0.f;
#pragma unroll
                    for (int t = 0; t < DSUB; t++) a = fmaf(qs[t], __ldg(c + t), a);
                    v = __fmul_rn(a, scale);
                }
                if (s & 1) acc1 = __fadd_rn(acc1, v); else acc0 = __fadd_rn(acc0, v);
            }
        }
        const float key = __fadd_rn(base, __fadd_rn(acc0, acc1));
        bool keep = key <= bound;
        if (keep && bitset) keep = !bit_is_set(bitset, rows[pos]);
        key_out = keep ? key : INFINITY;
        return keep ? pack_kp(key, pos) : kEmpty;
    };
    if (k_trim > 0 && n <= 512u && n > (uint32_t)k_trim) {   // CTA-uniform
        // Trim to the k_trim best.  Nearly every logged survivor passes `key <= bound` (the filter's margin is small), and the
        // bound itself comes from a sample of the codes, so a row holds ~4x the k' entries finalize needs (C3: 146 on average,
        // rows above 256 went to the CTA-wide finalize).  All exact keys of the row are here: a 256-bin histogram over their
        // range gives a cut that keeps the k_trim smallest (+ the rest of the crossing bin); the kept entries are compacted
        // to the front of the row (every thread holds its entries in registers before the first write) and the row's count is
        // rewritten.  Shared memory stays at ~1.5 KB so that the L1 keeps serving the table gathers.
        uint64_t pk[4];
        float kf[4];
        float lo = INFINITY, hi = -INFINITY;
        s_h[threadIdx.x] = 0;
        s_h[threadIdx.x + 128] = 0;
        if (threadIdx.x == 0) s_ctl[1] = 0;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const uint32_t i = threadIdx.x + it * 128;
            pk[it] = kEmpty;
            kf[it] = INFINITY;
            if (i < n) pk[it] = eval(row[i], kf[it]);
            if (kf[it] < INFINITY) { lo = fminf(lo, kf[it]); hi = fmaxf(hi, kf[it]); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = lo; s_red[4 + (threadIdx.x >> 5)] = hi; }
        __syncthreads();   // all entries of the row are in registers from here on
        lo = fminf(fminf(s_red[0], s_red[1]), fminf(s_red[2], s_red[3]));
        hi = fmaxf(fmaxf(s_red[4], s_red[5]), fmaxf(s_red[6], s_red[7]));
        const float sc = (hi > lo) ? 256.f / (hi - lo) : 0.f;
        int bin[4];
#pragma unroll
        for (int it = 0; it < 4; it++) {
            bin[it] = 256;
            if (kf[it] < INFINITY) {
                bin[it] = min(255, (int)((kf[it] - lo) * sc));
                atomicAdd(&s_h[bin[it]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint32_t h[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { h[t] = s_h[lane * 8 + t]; sum += h[t]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t excl = incl - sum;
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0 && total < (uint32_t)k_trim) s_ctl[0] = 255u;   // fewer than k_trim valid entries: keep them all
            if (excl < (uint32_t)k_trim && incl >= (uint32_t)k_trim) {
                uint32_t run = excl;
                int b = lane * 8 + 7;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    run += h[t];
                    if (run >= (uint32_t)k_trim) { b = lane * 8 + t; break; }
                }
                s_ctl[0] = (uint32_t)b;
            }
        }
        __syncthreads();
        const int bstar = (int)s_ctl[0];
#pragma unroll
        for (int it = 0; it < 4; it++)
            if (bin[it] <= bstar) row[atomicAdd(&s_ctl[1], 1u)] = pk[it];
        __syncthreads();
        if (threadIdx.x == 0) cand_cnt[q] = s_ctl[1];
        return;
    }
    float lo = INFINITY, hi = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += 128) {
        float kf;
        row[i] = eval(row[i], kf);
        if (kf < INFINITY) { lo = fminf(lo, kf); hi = fmaxf(hi, kf); }
    }
    if (k_trim > 0 && k_trim <= 96 && n > 512u) {   // CTA-uniform
        // Long rows (queries in dense regions: up to `cap` logged survivors) are what the CTA-wide finalize spent its time on
        // (a 2048-entry bitonic sort for the 40 best).  Same cut as above, but the entries stay in global memory: every thread
        // re-reads the packed exact keys it has just written (its own stores), the kept ones are staged in shared memory
        // (at most 128, otherwise the row is left as it is) and written to the front of the row after a barrier.
        __shared__ uint64_t s_stage[128];
        s_h[threadIdx.x] = 0;
        s_h[threadIdx.x + 128] = 0;
        if (threadIdx.x == 0) { s_ctl[0] = 0xffffffffu; s_ctl[1] = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = lo; s_red[4 + (threadIdx.x >> 5)] = hi; }
        __syncthreads();
        lo = fminf(fminf(s_red[0], s_red[1]), fminf(s_red[2], s_red[3]));
        hi = fmaxf(fmaxf(s_red[4], s_red[5]), fmaxf(s_red[6], s_red[7]));
        const float sc = (hi > lo) ? 256.f / (hi - lo) : 0.f;
        for (uint32_t i = threadIdx.x; i < n; i += 128) {
            const uint64_t e = row[i];
            if (e != kEmpty) atomicAdd(&s_h[min(255, (int)((unpack_key(e) - lo) * sc))], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint32_t h[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { h[t] = s_h[lane * 8 + t]; sum += h[t]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t excl = incl - sum;
            if (excl < (uint32_t)k_trim && incl >= (uint32_t)k_trim) {
                uint32_t run = excl;
                int b = lane * 8 + 7;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    run += h[t];
                    if (run >= (uint32_t)k_trim) { b = lane * 8 + t; break; }
                }
                if (run <= 128u) s_ctl[0] = (uint32_t)b;   // entries in bins [0, b]: they fit the staging buffer
            }
        }
        __syncthreads();
        const uint32_t bstar = s_ctl[0];
        if (bstar == 0xffffffffu) return;   // fewer than k_trim valid entries, or a crowded crossing bin: leave the row alone
        for (uint32_t i = threadIdx.x; i < n; i += 128) {
            const uint64_t e = row[i];
            if (e != kEmpty && (uint32_t)min(255, (int)((unpack_key(e) - lo) * sc)) <= bstar) s_stage[atomicAdd(&s_ctl[1], 1u)] = e;
        }
        __syncthreads();   // every read of the row is done
        const uint32_t nk = s_ctl[1];
        for (uint32_t i = threadIdx.x; i < nk; i += 128) row[i] = s_stage[i];
        if (threadIdx.x == 0) cand_cnt[q] = nk;
    }
}

// queries whose nearest list is owned by this rank -> compact list: this rank runs their phase A (one CTA; order is irrelevant)
__global__ void __launch_bounds__(1024)
compact_resp_kernel(const int64_t* __restrict__ probe_ids, int nprobe, int64_t nq, const int32_t* __restrict__ list_owner, int rank,
                    int32_t* __restrict__ list, uint32_t* __restrict__ count) {
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int64_t q = threadIdx.x; q < nq; q += 1024) {
        const int64_t l = probe_ids[q * nprobe];
        if (l >= 0 && list_owner[l] == rank) list[atomicAdd(&s_n, 1u)] = (int32_t)q;
    }
    __syncthreads();
    if (threadIdx.x == 0) *count = s_n;
}
__global__ void
fill_f32_kernel(float* __restrict__ out, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
// max |x[i]| (one atomicMax on the bit pattern of a non-negative float)
__global__ void __launch_bounds__(256)
max_abs_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// flagged queries -> compact list (one CTA; order is irrelevant)
__global__ void __launch_bounds__(1024)
compact_flags_kernel(const uint32_t* __restrict__ qflag, int64_t nq, int32_t* __restrict__ list, uint32_t* __restrict__ count) {
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int64_t q = threadIdx.x; q < nq; q += 1024)
        if (qflag[q]) list[atomicAdd(&s_n, 1u)] = (int32_t)q;
    __syncthreads();
    if (threadIdx.x == 0) *count = s_n;
}

// Phase A: exact keys of the codes in the nearest probed lists of every query — lists are taken in probe order until
// `min_codes` codes were seen (at most `p0_max` lists; lists of other shards have length 0 and cost nothing) —
// the k_need-th smallest key (rounded up to a histogram bin edge) -> entry k_need-1 of the query's row = the admission
// bound of the filter pass.  When the lists did not
// hold 4*k_need codes the bound would be loose (a large share of every probed list would survive): the row is left
// without a bound and the LUT kernel redoes that query.
// The query's table is copied from `lut` into a skewed shared layout: code value j owns a row of 32 words,
// row[w] = LUT[w % 16][j]; lane i reads word (i % 16) + s at step s, i.e. sub-quantizer (pos + s) % 16 — the address
// is one byte-permute plus an immediate (PRMT + LDS + FADD per look-up, like the LUT kernel), lanes i and i+16 share
// a bank (2 wavefronts per gather).  The keys go to shared memory and the bound is read off a 1024-bin histogram
// (no top-k structure at all).  grid = nq, block = NT (128 or 256); dynamic smem = bound_smem(ROWW, bound_kmax(...)) (48 KB at C3).
#define KB2_BOUND_STEP(WORD, KB, S, ACC)                                                      \
    {                                                                                         \
        const uint32_t _x = __byte_perm((WORD), lane4, 0x6504u | ((KB) << 4));                \
        float _v;                                                                             \
        asm("ld.shared.f32 %0, [%1+%2];" : "=f"(_v) : "r"(_x >> (ROWW == 32 ? 1 : 0)), "n"(KB2_SMEM_BASE + 4 * (S))); \
        ACC += _v;                                                                            \
    }
// pqc [M][256][dsub] -> pqc_t [M/16][256][16][dsub] (the 16 sub-quantizers of a group adjacent: coalesced table builds)
__global__ void
transpose_codebook_kernel(const float* __restrict__ pqc, int M, int dsub, float* __restrict__ pqc_t) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)M * 256 * dsub;
    if (i >= n) return;
    const int x = (int)(i % dsub);
    const int j = (int)((i / dsub) % 256);
    const int m = (int)(i / ((int64_t)dsub * 256));
    pqc_t[((((int64_t)(m >> 4) * 256 + j) * 16) + (m & 15)) * dsub + x] = pqc[i];
}
constexpr int BOUND_KMAX = 6144;    // keys held per query (phase A looks at no more codes than this)
constexpr int BOUND_BINS = 1024;
// ROWW = words per code-value row of the skewed table: 32 (32 KB: lanes i and i+16 share a bank, 2 wavefronts per gather,
// 3 CTAs/SM) or 64 (64 KB: conflict-free like the LUT kernel, 2 CTAs/SM)
// keys actually held for a launch: the requested number of codes rounded up (C3: 3000 -> 3008 keys = 12 KB instead of 24 KB, i.e.
// 48 KB per CTA and four CTAs per SM instead of three)
__host__ __device__ constexpr int bound_kmax(int min_codes, int k_need) {
    const int want = ((min_codes > k_need ? min_codes : k_need) + 63) & ~63;
    return want < BOUND_KMAX ? want : BOUND_KMAX;
}
constexpr size_t bound_smem(int roww, int kmax = BOUND_KMAX) { return (size_t)roww * 1024 + (size_t)kmax * 4 + BOUND_BINS * 4 + 128; }
constexpr size_t BOUND_SMEM = bound_smem(32);

// G > 1 (m = 16 G sub-quantizers, e.g. m48 x dsub2): the groups are scanned one after the other through the same 32 KB
// table -- group g's table is built in the kernel from the query and the transposed codebook `pqc_t`
// ([g][code value][16 sub-quantizers][dsub], see transpose_codebook_kernel), the partial sums of the earlier groups wait in
// the shared key array.
// NT = threads per CTA: 128 (4 warps) or 256 (8 warps over the same tables: twice the gathers in flight per shared-memory byte)
template <int METRIC, int ROWW, int G = 1, int DSUB = 8, int NT = 128>
__global__ void __launch_bounds__(NT)
bound_kernel(const float* __restrict__ lut, const int32_t* __restrict__ qlist, const uint32_t* __restrict__ qcount, int64_t nq,
             const int64_t* __restrict__ probe_ids, const float* __restrict__ probe_dis,
             int probe_stride, int p0_max, int min_codes, int k_need, const int64_t* __restrict__ list_off,
             const int32_t* __restrict__ list_len, const uint4* __restrict__ codes, const float* __restrict__ t1,
             const uint8_t* __restrict__ bitset, const int32_t* __restrict__ rows, float* __restrict__ out,
             unsigned long long* __restrict__ counters, int64_t npad = 0, const float* __restrict__ queries = nullptr,
             const float* __restrict__ pqc_t = nullptr) {
    static_assert(G == 1 || ROWW == 32, "multi-group phase A uses the 32-word table rows");
    static_assert(NT == 128 || NT == 256, "bound_kernel block size");
    constexpr int NW = NT / 32;               // warps
    constexpr int BPT = BOUND_BINS / NT;      // histogram bins owned by a thread
    // work list: table i / query qlist[i] for i < *qcount (qlist == NULL: query i, i < nq); CTAs stride the list
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_lut = (float*)smem_raw;                              // [256][ROWW]
    const int key_cap = bound_kmax(min_codes, k_need);            // (the host sizes the shared memory with the same function)
    float* s_keys = (float*)(smem_raw + ROWW * 1024);             // [key_cap]
    uint32_t* s_hist = (uint32_t*)(s_keys + key_cap);             // [BOUND_BINS]
    float* s_red = (float*)(s_hist + BOUND_BINS);                 // [32]: min [0,8) max [8,16) warp sums [16,24)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((uint32_t)__cvta_generic_to_shared(smem_raw) != (uint32_t)KB2_SMEM_BASE) {
        if (threadIdx.x == 0 && counters) atomicExch(counters + 1, 0xBAD5ull);   // layout assumption violated: host raises an error
        return;
    }
    const int64_t n_work = qlist ? (int64_t)*qcount : nq;
    for (int64_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const int64_t q = qlist ? (int64_t)qlist[wi] : wi;
    int seen = 0, n_tot = 0;
    float kmin = INFINITY, kmax = -INFINITY;
#pragma unroll 1
    for (int g = 0; g < G; g++) {
    __syncthreads();   // the previous iteration's readers of the shared tables are done
    if constexpr (G > 1 || DSUB != 8) {
        // table of group g from the query: entry (j, mm) = scale * <q_m, c_pq[m][j]>, m = 16 g + mm
        const float scale = (METRIC == KB2_METRIC_L2) ? -2.f : -1.f;
        const int mm = threadIdx.x & 15, jsub = threadIdx.x >> 4;   // 16 sub-quantizers x 8 code values per pass
        float qv[DSUB];
#pragma unroll
        for (int x = 0; x < DSUB; x++) qv[x] = queries[q * (int64_t)(16 * G * DSUB) + (g * 16 + mm) * DSUB + x];
        for (int j = jsub; j < 256; j += NT / 16) {
            const float* cp = pqc_t + (((size_t)g * 256 + j) * 16 + mm) * DSUB;
            float a = 0.f;
#pragma unroll
            for (int x = 0; x < DSUB; x++) a = fmaf(qv[x], __ldg(cp + x), a);
            a *= scale;
            s_lut[j * ROWW + mm] = a;
            s_lut[j * ROWW + 16 + mm] = a;
        }
    } else {
        // lut[q][j*16 + m] -> s_lut[j*32 + m] and s_lut[j*32 + 16 + m]
        const float4* src = reinterpret_cast<const float4*>(lut + wi * 4096);
#pragma unroll
        for (int i = 0; i < 1024 / NT; i++) {
            const int idx = threadIdx.x + i * NT;        // float4 index: j = idx / 4, m4 = (idx % 4) * 4
            const float4 v = __ldg(src + idx);
            float4* dst = reinterpret_cast<float4*>(s_lut + (idx >> 2) * ROWW + (idx & 3) * 4);
            dst[0] = v;
            dst[4] = v;
            if (ROWW == 64) { dst[8] = v; dst[12] = v; }
        }
    }
    if (g == 0)
        for (int i = threadIdx.x; i < BOUND_BINS; i += NT) s_hist[i] = 0;
    __syncthreads();
    // PRMT builds (byte << 8) | (lane16 << 3) ; >> 1 = byte * 128 + lane16 * 4 (row pitch 128 B)
    // ROWW = 64: (byte << 8) | (lane << 2) is the address itself (row pitch 256 B, word lane + s)
    const uint32_t lane4 = (ROWW == 32) ? ((uint32_t)(lane & 15) << 3) : ((uint32_t)lane << 2);
    const uint4* gcodes = codes + (int64_t)g * npad;   // code plane of this group
    const bool first_g = (g == 0), last_g = (g == G - 1);
    seen = 0;
    n_tot = 0;
    const int code_cap = min(key_cap, max(min_codes, k_need));   // scan no more than the requested number of codes
    for (int j = 0; j < p0_max && seen < min_codes && n_tot < code_cap; j++) {
        const int64_t l = probe_ids[q * probe_stride + j];
        if (l < 0) continue;
        const int len_all = list_len[l];
        if (len_all == 0) continue;
        seen += len_all;
        const int len = min(len_all, code_cap - n_tot);   // any subset of the codes still yields a valid upper bound
        const int64_t off = list_off[l];
        const float dv = probe_dis[q * probe_stride + j];
        const float base = (METRIC == KB2_METRIC_L2) ? dv : -dv;
        // two chunks per iteration, software-pipelined: the code words / row terms of the NEXT iteration are in
        // flight while the gather chains of this one run (the kernel was latency bound on these loads)
        uint4 nA = make_uint4(0, 0, 0, 0), nB = nA;
        float ntA = 0.f, ntB = 0.f;
        auto load_iter = [&](int c0) {
            if (c0 < len) {
                nA = ldg_stream_u4(gcodes + off + c0 + lane);        // inside the padded position space even past len
                if (METRIC == KB2_METRIC_L2 && first_g) ntA = __ldg(t1 + off + c0 + lane);
                if (c0 + 32 < len) {
                    nB = ldg_stream_u4(gcodes + off + c0 + 32 + lane);
                    if (METRIC == KB2_METRIC_L2 && first_g) ntB = __ldg(t1 + off + c0 + 32 + lane);
                }
            }
        };
        load_iter(warp * 64);
        for (int c0 = warp * 64; c0 < len; c0 += NW * 64) {
            const int relA = c0 + lane, relB = c0 + 32 + lane;
            const bool okA = relA < len, okB = relB < len;
            const uint32_t posA = (uint32_t)(off + relA), posB = (uint32_t)(off + relB);
            const bool hasB = c0 + 32 < len;
            const uint4 wA = nA;
            const uint4 wB = hasB ? nB : nA;
            float a0 = ntA, a1 = 0.f, b0 = hasB ? ntB : 0.f, b1 = 0.f;
            load_iter(c0 + NW * 64);
            KB2_BOUND_STEP(wA.x, 0, 0, a0)  KB2_BOUND_STEP(wB.x, 0, 0, b0)  KB2_BOUND_STEP(wA.x, 1, 1, a1)  KB2_BOUND_STEP(wB.x, 1, 1, b1)
            KB2_BOUND_STEP(wA.x, 2, 2, a0)  KB2_BOUND_STEP(wB.x, 2, 2, b0)  KB2_BOUND_STEP(wA.x, 3, 3, a1)  KB2_BOUND_STEP(wB.x, 3, 3, b1)
            KB2_BOUND_STEP(wA.y, 0, 4, a0)  KB2_BOUND_STEP(wB.y, 0, 4, b0)  KB2_BOUND_STEP(wA.y, 1, 5, a1)  KB2_BOUND_STEP(wB.y, 1, 5, b1)
            KB2_BOUND_STEP(wA.y, 2, 6, a0)  KB2_BOUND_STEP(wB.y, 2, 6, b0)  KB2_BOUND_STEP(wA.y, 3, 7, a1)  KB2_BOUND_STEP(wB.y, 3, 7, b1)
            KB2_BOUND_STEP(wA.z, 0, 8, a0)  KB2_BOUND_STEP(wB.z, 0, 8, b0)  KB2_BOUND_STEP(wA.z, 1, 9, a1)  KB2_BOUND_STEP(wB.z, 1, 9, b1)
            KB2_BOUND_STEP(wA.z, 2, 10, a0) KB2_BOUND_STEP(wB.z, 2, 10, b0) KB2_BOUND_STEP(wA.z, 3, 11, a1) KB2_BOUND_STEP(wB.z, 3, 11, b1)
            KB2_BOUND_STEP(wA.w, 0, 12, a0) KB2_BOUND_STEP(wB.w, 0, 12, b0) KB2_BOUND_STEP(wA.w, 1, 13, a1) KB2_BOUND_STEP(wB.w, 1, 13, b1)
            KB2_BOUND_STEP(wA.w, 2, 14, a0) KB2_BOUND_STEP(wB.w, 2, 14, b0) KB2_BOUND_STEP(wA.w, 3, 15, a1) KB2_BOUND_STEP(wB.w, 3, 15, b1)
            float keyA = (first_g ? base : s_keys[n_tot + (okA ? relA : 0)]) + (a0 + a1);
            float keyB = (first_g ? base : s_keys[n_tot + ((hasB && okB) ? relB : 0)]) + (b0 + b1);
            if (okA) {
                if (last_g && bitset && bit_is_set(bitset, rows[posA])) keyA = INFINITY;
                s_keys[n_tot + relA] = keyA;
                if (last_g && keyA < INFINITY) { kmin = fminf(kmin, keyA); kmax = fmaxf(kmax, keyA); }
            }
            if (hasB && okB) {
                if (last_g && bitset && bit_is_set(bitset, rows[posB])) keyB = INFINITY;
                s_keys[n_tot + relB] = keyB;
                if (last_g && keyB < INFINITY) { kmin = fminf(kmin, keyB); kmax = fmaxf(kmax, keyB); }
            }
        }
        n_tot += len;
    }
    }   // groups
    // ---- K-th smallest key, rounded UP to the edge of one of 1024 linear bins over [min, max]: any value >= the
    //      k_need-th best is a valid admission bound, and a bin is far narrower than the filter's error margin
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = fminf(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = fmaxf(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) { s_red[warp] = kmin; s_red[8 + warp] = kmax; }
    __syncthreads();
    float lo = s_red[0], hi = s_red[8];
#pragma unroll
    for (int w = 1; w < NW; w++) { lo = fminf(lo, s_red[w]); hi = fmaxf(hi, s_red[8 + w]); }
    const float scale = (hi > lo) ? (float)BOUND_BINS / (hi - lo) : 0.f;
    for (int i = threadIdx.x; i < n_tot; i += NT) {
        const float kv = s_keys[i];
        if (kv < INFINITY) atomicAdd(&s_hist[min(BOUND_BINS - 1, (int)((kv - lo) * scale))], 1u);
    }
    __syncthreads();
    // thread t owns bins [BPT t, BPT (t+1)): exclusive prefix over threads, then the owner of the crossing writes the bound
    uint32_t mine[BPT], tsum = 0;
#pragma unroll
    for (int b = 0; b < BPT; b++) { mine[b] = s_hist[threadIdx.x * BPT + b]; tsum += mine[b]; }
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t* s_wsum = (uint32_t*)(s_red + 16);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t before = incl - tsum;
    for (int w = 0; w < warp; w++) before += s_wsum[w];
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) total += s_wsum[w];
    const uint32_t need = (uint32_t)k_need;
    if (total < need || seen < 4 * k_need) {
        if (threadIdx.x == 0) out[q] = INFINITY;   // no (or only a loose) bound: the LUT kernel redoes the query
    } else if (before < need && before + tsum >= need) {
        uint32_t cum = before;
        int b = 0;
#pragma unroll
        for (int bb = 0; bb < BPT; bb++) {
            if (cum < need) { cum += mine[bb]; b = bb; }
        }
        const float bound = (scale > 0.f) ? lo + ((float)(threadIdx.x * BPT + b) + 1.01f) / scale : hi;
        out[q] = fmaxf(bound, lo) + (G == 1 ? 4e-7f : 2e-6f) * fmaxf(fabsf(lo), fabsf(hi));
    }
    }   // work list
}
#undef KB2_BOUND_STEP

}  // namespace pqtc
}  // namespace kb2
