// kb2_gemm_tc.cuh — the dense query x base contraction on the 5th-generation tensor cores.
//
//   keys[q][j] = |q|^2 + |x_j|^2 - 2 <q, x_j>   (L2)        keys[q][j] = -<q, x_j>   (IP)
//
// Same contract and output as gemm_keys_kernel (kb2_flat.cuh), which stays as the bit-reproducible
// fp32 reference / fallback (d % 4 != 0).  Used by FLAT, BruteForce and the IVF coarse quantizer
// (reference: F/utils/distances.cpp:326-363,834-875; F/IndexIVF.cpp:336-342).
//
// sm_100a mapping
//   * operands: fp32 rows, K-major.  TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) brings 128 x 32-float
//     boxes (one 128-byte swizzle row per tensor row) of Q and X into a 3-stage shared-memory ring.
//   * fp32 fidelity on a tf32 pipe: 4 "converter" warps split every element into hi = top 19 bits and
//     lo = x - hi (both rounded to tf32), writing hi in place and lo to a twin tile; the MMA warp issues
//     D += hi*hi + hi*lo + lo*hi  (3 x tcgen05.mma.kind::tf32, M=128 N=128 K=8) — error ~2^-21 relative,
//     and the k+16 best candidates are re-ranked exactly afterwards anyway (finalize_kernel).
//   * accumulator: 128 lanes x 128 columns of TMEM (fp32); tcgen05.commit signals stage release and
//     accumulator completion through mbarriers; the same 4 warps then drain TMEM with tcgen05.ld
//     (32x32b.x32), apply the key epilogue (+norms, bitset) and store 128-bit rows.
//   * one 128x128 output tile per CTA (K = d is short: 4 k-blocks at d=128, so the kernel is bound by
//     operand/epilogue traffic, not by the tensor pipe — see DESIGN.md 4.1).
#pragma once
#include <cuda.h>

#include "kb2_common.cuh"

namespace kb2 {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;                      // ring depth of the long-K instantiation (1 CTA/SM)
constexpr int TILE_BYTES = 128 * BK * 4;       // 16 KB: a 128-row x 128-byte tile (A or B)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi | B_hi | A_lo | B_lo
constexpr int THREADS = 192;                   // warp 0: TMA   warp 1: MMA + TMEM alloc   warps 2-5: convert + epilogue
constexpr int CONV_THREADS = 128;
constexpr int TMEM_COLS = 128;
constexpr size_t smem_bytes(int nst) { return (size_t)nst * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/; }
constexpr size_t SMEM_BYTES = smem_bytes(STAGES);
// Short contractions (d <= 192: the IVF coarse quantizer, k-means assignment) run a single-stage instantiation with three
// CTAs per SM instead: a 128x128 tile is then ~8 us of strictly serial TMA -> split -> MMA -> store, and with one CTA per SM
// (ncu r2: 9 % warps active, 17 waves) nothing overlaps the 64 KB epilogue store; three resident CTAs overlap each other.

__device__ __forceinline__ uint32_t
smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void
mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void
mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void
mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void
mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "KB2_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra KB2_DONE;\n\t"
        "bra KB2_WAIT;\n\t"
        "KB2_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void
tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void
fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void
tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void
tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void
tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void
tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart, sm_100 version bit
__device__ __forceinline__ uint64_t
make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units        bits  0-13
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused here)   bits 16-29
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset = 1024 B         bits 32-45
    d |= (uint64_t)1 << 46;                         // descriptor version 1 (sm_100)       bits 46-47
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B            bits 61-63
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=128
__device__ __forceinline__ uint32_t
make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// round-to-nearest into the 19-bit tf32 container (low 13 mantissa bits cleared)
__device__ __forceinline__ float
tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

template <int METRIC, int STAGES = 3>
__global__ void __launch_bounds__(THREADS, STAGES == 1 ? 3 : 1)
gemm_keys_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                    const float* __restrict__ qn, const float* __restrict__ xn, int nq, int nb, int d,
                    float* __restrict__ keys, int64_t ldk, const uint8_t* __restrict__ bitset,
                    const int32_t* __restrict__ rows, int64_t row_base) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* base_ptr = smem_dyn + (base - raw);
    const uint32_t bars = base + STAGES * STAGE_BYTES;      // barrier block after the ring
    // barrier layout (8 bytes each): full_raw[S] | full_conv[S] | empty[S] | tmem_full | tmem slot(4B)
    auto bar_full_raw = [&](int s) { return bars + 8u * s; };
    auto bar_full_conv = [&](int s) { return bars + 8u * (STAGES + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t bar_tmem_full = bars + 8u * (3 * STAGES);
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 1);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.y * BM;
    const int j0 = blockIdx.x * BN;
    const int nkb = (d + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full_raw(s), 1);
            mbar_init(bar_full_conv(s), CONV_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int it = 0; it < nkb; it++) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(bar_empty(s), ph ^ 1u);
                const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
                mbar_expect_tx(bar_full_raw(s), 2 * TILE_BYTES);
                tma_load_2d(st, &tmap_q, it * BK, q0, bar_full_raw(s));                  // A_hi slot (raw fp32)
                tma_load_2d(st + TILE_BYTES, &tmap_x, it * BK, j0, bar_full_raw(s));     // B_hi slot (raw fp32)
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc();
        for (int it = 0; it < nkb; it++) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(bar_full_conv(s), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / 8; kk++) {
                    const uint32_t ko = (uint32_t)kk * 32u;   // 8 tf32 = 32 bytes inside the 128-byte swizzle row
                    const uint64_t a_hi = make_desc(st + ko);
                    const uint64_t b_hi = make_desc(st + TILE_BYTES + ko);
                    const uint64_t a_lo = make_desc(st + 2 * TILE_BYTES + ko);
                    const uint64_t b_lo = make_desc(st + 3 * TILE_BYTES + ko);
                    tc_mma_tf32(tmem_base, a_hi, b_hi, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                    tc_mma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
                    tc_mma_tf32(tmem_base, a_lo, b_hi, idesc, 1u);
                }
                tc_commit(bar_empty(s));                       // frees the smem stage when these MMAs retire
                if (it == nkb - 1) tc_commit(bar_tmem_full);   // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ================= converters (hi/lo split), then epilogue =================
        const int t = threadIdx.x - 64;  // 0..127
        for (int it = 0; it < nkb; it++) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(bar_full_raw(s), ph);
            float4* hi = reinterpret_cast<float4*>(base_ptr + (size_t)s * STAGE_BYTES);        // A_hi|B_hi contiguous
            float4* lo = reinterpret_cast<float4*>(base_ptr + (size_t)s * STAGE_BYTES + 2 * TILE_BYTES);
#pragma unroll 4
            for (int i = t; i < 2 * TILE_BYTES / 16; i += CONV_THREADS) {
                float4 v = hi[i];
                float4 h, l;
                // hi = x rounded to tf32 (nearest), lo = (x - hi) rounded to tf32: both exactly representable,
                // so the tensor core's own truncation of the low 13 bits changes nothing
                h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
                h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
                h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
                h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
                hi[i] = h;
                lo[i] = l;
            }
            fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bar_full_conv(s));
        }
        // ---- epilogue: TMEM -> registers -> keys
        mbar_wait(bar_tmem_full, 0);
        tc_fence_after();
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q0 + quarter * 32 + lane;
        const float qq = (METRIC == KB2_METRIC_L2 && row < nq) ? qn[row] : 0.f;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                  "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                  "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < nq) {
                float* out = keys + (int64_t)row * ldk + j0 + c;
#pragma unroll
                for (int v4 = 0; v4 < 8; v4++) {
                    float o[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int col = j0 + c + v4 * 4 + u;
                        const float acc = __uint_as_float(r[v4 * 4 + u]);
                        float key = INFINITY;
                        if (col < nb) {
                            key = (METRIC == KB2_METRIC_L2) ? (qq + xn[col] - 2.f * acc) : -acc;
                            if (bitset) {
                                const int64_t rr = rows ? (int64_t)rows[row_base + col] : (row_base + col);
                                if (bit_is_set(bitset, rr)) key = INFINITY;
                            }
                        }
                        o[u] = key;
                    }
                    const int col0 = j0 + c + v4 * 4;
                    if (col0 + 3 < ldk) {
                        *reinterpret_cast<float4*>(out + v4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
                        for (int u = 0; u < 4; u++)
                            if (col0 + u < ldk) out[v4 * 4 + u] = o[u];
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled
get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
        cudaGetLastError();
    });
    return fn;
}

// 2-D fp32 row-major [rows][d] -> box of 32 columns x 128 rows, 128-byte swizzle, zero fill out of bounds
inline bool
make_tmap(CUtensorMap* m, const float* ptr, int64_t rows, int d, int box_rows = 128) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (d & 3)) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)d * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace kb2
