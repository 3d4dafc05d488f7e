// kb2_index.cuh — host-side index objects behind the C ABI (FLAT, IVF_FLAT, IVF_PQ).
// They play the role of the reference's IndexNode implementations
//   FlatIndexNode  src/index/flat/flat.cc:33-427
//   IvfIndexNode   src/index/ivf/ivf.cc:68-1972   (IVF_FLAT + IVF_PQ branches)
// but hand the WHOLE query batch to the device in one call (like the in-tree GPU precedent,
// src/common/cuvs/integration/cuvs_knowhere_index.cuh:508-632) instead of nq thread-pool tasks.
#pragma once
#include <mutex>
#include <vector>

#include "kb2_build.cuh"
#include "kb2_comm.h"
#include "kb2_gemm_tc.cuh"
#include "kb2_ivf.cuh"
#include "kb2_ivfpq_tc.cuh"
#include "kb2_ivfflat_tc.cuh"
#include "kb2_json.h"

namespace kb2 {

constexpr int kMaxK = 1024;            // largest k' any selection kernel keeps
constexpr int kMaxSortEntries = 8192;  // finalize sorts at most this many candidates per query
constexpr int kMaxDynSmem = 227 * 1024;
#ifndef KB2_DEFAULT_SCAN_PREFETCH
#define KB2_DEFAULT_SCAN_PREFETCH 1
#endif
#ifndef KB2_DEFAULT_SCAN_NT
#define KB2_DEFAULT_SCAN_NT 256
#endif

inline void
init_kernel_attributes() {
    static PerDeviceOnce once;
    once.run([] {
        auto set = [](const void* f) {
            cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        };
        set((const void*)finalize_kernel);
        set((const void*)reduce_partials_kernel);
        set((const void*)select_keys_kernel);
        set((const void*)select_keys_hist_kernel);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_L2, false, 256>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_L2, true, 256>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_IP, false, 256>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_IP, true, 256>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_L2, false, 256>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_L2, true, 256>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_IP, false, 256>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_IP, true, 256>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_L2, false, 256>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_L2, true, 256>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_IP, false, 256>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_IP, true, 256>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_L2, false, 512>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_L2, true, 512>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_IP, false, 512>);
        set((const void*)ivfpq_scan_kernel<1, KB2_METRIC_IP, true, 512>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_L2, false, 512>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_L2, true, 512>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_IP, false, 512>);
        set((const void*)ivfpq_scan_kernel<2, KB2_METRIC_IP, true, 512>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_L2, false, 512>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_L2, true, 512>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_IP, false, 512>);
        set((const void*)ivfpq_scan_kernel<3, KB2_METRIC_IP, true, 512>);
        set((const void*)ivfpq_scan_generic_kernel<KB2_METRIC_L2>);
        set((const void*)ivfpq_scan_generic_kernel<KB2_METRIC_IP>);
        set((const void*)ivfflat_scan_kernel<KB2_METRIC_L2>);
        set((const void*)ivfflat_scan_kernel<KB2_METRIC_IP>);
        set((const void*)pqtc::ivfpq_tc_filter_kernel<KB2_METRIC_L2, 1, 8>);
        set((const void*)pqtc::ivfpq_tc_filter_kernel<KB2_METRIC_IP, 1, 8>);
        set((const void*)pqtc::ivfpq_tc_filter_kernel<KB2_METRIC_L2, 3, 2>);
        set((const void*)pqtc::ivfpq_tc_filter_kernel<KB2_METRIC_IP, 3, 2>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_L2, 32>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_IP, 32>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_L2, 64>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_IP, 64>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_L2, 32, 1, 8, 256>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_IP, 32, 1, 8, 256>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_L2, 32, 3, 2>);
        set((const void*)pqtc::bound_kernel<KB2_METRIC_IP, 32, 3, 2>);
        set((const void*)fltc::ivfflat_tc_kernel<KB2_METRIC_L2, 32>);
        set((const void*)fltc::ivfflat_tc_kernel<KB2_METRIC_IP, 32>);
        set((const void*)fltc::ivfflat_tc_kernel<KB2_METRIC_L2, 128>);
        set((const void*)fltc::ivfflat_tc_kernel<KB2_METRIC_IP, 128>);
        cudaGetLastError();
    });
}

struct Counters {
    int64_t launches = 0, codes = 0, code_bytes = 0, pairs = 0, h2d = 0, d2h = 0;
    int64_t survivors = 0, flagged = 0;   // tensor-core PQ engine: codes re-evaluated exactly / queries redone by the LUT kernel
};

// grow-by-doubling append of `count` elements (device->device or host->device)
template <typename T>
inline void
dev_append(DevBuf<T>& buf, size_t& used, const T* src, size_t count, cudaStream_t st) {
    if (used + count > buf.n) {
        size_t cap = std::max(used + count, buf.n * 2);
        DevBuf<T> nb;
        nb.ensure(cap);
        if (used) KB2_CUDA_CHECK(cudaMemcpyAsync(nb.p, buf.p, used * sizeof(T), cudaMemcpyDeviceToDevice, st));
        KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        buf = std::move(nb);
    }
    if (count) KB2_CUDA_CHECK(cudaMemcpyAsync(buf.p + used, src, count * sizeof(T), cudaMemcpyDefault, st));
    used += count;
}

// ============================================================================================
struct IndexBase {
    std::string type;
    int metric = KB2_METRIC_L2, dim = 0, device = 0;
    bool cosine = false;   // COSINE: metric == IP over vectors normalised on entry (see kb2_index_create)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int shard_rank = 0, shard_world = 1;
    std::mutex mu;
    Counters last;
    bool timing = false;
    float last_kernel_ms = 0.f;      // dominant kernel of the last search (IVF_PQ tensor-core engine: the filter kernel)
    float last_stage_ms = 0.f;       // whole list-scan stage of the last search (all engines)
    int last_engine = 0;             // 0: query-major scan kernels, 1: list-major tensor-core engine
    float last_comm_ms = 0.f;        // collectives (+ merge) of the last sharded search
    Comm* comm = nullptr;            // not owned (kb2_index_set_comm)
    virtual void set_comm(Comm* c) { comm = c; }
    bool distributed() const { return comm != nullptr && shard_world > 1; }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev_in = nullptr;
    cudaEvent_t ev_c0 = nullptr, ev_c1 = nullptr, ev_c2 = nullptr, ev_c3 = nullptr;   // collectives of a sharded search
    DevBuf<unsigned long long> d_counter;
    PinnedBuf h_counter;   // pinned landing zone of the per-search device counters (no pageable async copy)

    // per-search scratch (grow-only, reused across calls)
    DevBuf<float> s_q, s_keys, s_qn, s_out_dist, s_probe_dis;
    DevBuf<uint64_t> s_partial, s_partial2;
    DevBuf<int64_t> s_out_ids, s_probe_ids;
    DevBuf<uint8_t> s_bitset;
    DevBuf<float> s_cos_in, s_cos_out, s_typed_f32;
    // multi-GPU (kb2_index_set_comm): staging of the local top-k and of the gathered per-shard candidates
    DevBuf<int64_t> s_loc_ids, s_g_ids;
    DevBuf<float> s_loc_dist, s_g_dist;
    void
    ensure_gather_buffers(int64_t nq, int k) {
        s_loc_ids.ensure((size_t)nq * k);
        s_loc_dist.ensure((size_t)nq * k);
        s_g_ids.ensure((size_t)shard_world * nq * k);
        s_g_dist.ensure((size_t)shard_world * nq * k);
    }
    DevBuf<uint8_t> s_typed_raw;

    // L2-normalised device copy of n rows (COSINE)
    const float*
    normalized(const float* x, int64_t n) {
        if (n <= 0 || !x) return x;
        const float* dx = to_device(x, (size_t)n * dim, s_cos_in);
        s_cos_out.ensure((size_t)n * dim);
        normalize_rows_kernel<<<grid1d(n * 32, 256), 256, 0, stream>>>(dx, n, dim, s_cos_out.p);
        KB2_CUDA_CHECK(cudaGetLastError());
        return s_cos_out.p;
    }

    virtual ~IndexBase() {
        for (cudaEvent_t e : {ev0, ev1, ev2, ev3, ev_in, ev_c0, ev_c1, ev_c2, ev_c3})
            if (e) cudaEventDestroy(e);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
    // Stream contract: work runs on the handle's stream.  With the library-owned (non-blocking) stream, device buffers
    // handed in by the caller may still be in flight on the caller's side: order our stream after everything already
    // queued on the legacy default stream (which itself waits for all blocking streams, e.g. torch's default stream).
    // Callers that produce inputs on their own NON-blocking stream pass it through kb2_index_set_stream instead.
    void
    wait_caller_work() {
        if (!own_stream || !ev_in) return;
        KB2_CUDA_CHECK(cudaEventRecord(ev_in, cudaStreamLegacy));
        KB2_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_in, 0));
    }
    void
    init_common() {
        KB2_CUDA_CHECK(cudaSetDevice(device));
        init_kernel_attributes();
        KB2_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
        KB2_CUDA_CHECK(cudaEventCreate(&ev0));
        KB2_CUDA_CHECK(cudaEventCreate(&ev1));
        KB2_CUDA_CHECK(cudaEventCreate(&ev2));
        KB2_CUDA_CHECK(cudaEventCreate(&ev3));
        KB2_CUDA_CHECK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        for (cudaEvent_t* e : {&ev_c0, &ev_c1, &ev_c2, &ev_c3}) KB2_CUDA_CHECK(cudaEventCreate(e));
        d_counter.ensure(16);
        h_counter.ensure(128);
    }
    void
    set_stream(cudaStream_t s) {
        if (own_stream && stream) cudaStreamDestroy(stream);
        stream = s;
        own_stream = false;
    }
    void
    use_own_stream() {
        if (own_stream) return;
        KB2_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }

    // returns a device pointer to `count` floats of `src` (copying H2D on the stream if needed)
    const float*
    to_device(const float* src, size_t count, DevBuf<float>& buf, bool count_io = true) {
        if (is_device_ptr(src)) return src;
        buf.ensure(count);
        KB2_CUDA_CHECK(cudaMemcpyAsync(buf.p, src, count * sizeof(float), cudaMemcpyHostToDevice, stream));
        if (count_io) last.h2d += (int64_t)(count * sizeof(float));
        return buf.p;
    }
    const uint8_t*
    bitset_to_device(const uint8_t* bits, int64_t nbits) {
        if (!bits || nbits <= 0) return nullptr;
        // the kernels index the bitmap by any stored row: a shorter bitmap would be read out of bounds
        KB2_REQUIRE(nbits >= bitset_rows(), KB2_INVALID_ARGS, "bitset has fewer bits than the index has rows");
        if (is_device_ptr(bits)) return bits;
        const size_t nbytes = (size_t)((nbits + 7) / 8);
        s_bitset.ensure(nbytes);
        KB2_CUDA_CHECK(cudaMemcpyAsync(s_bitset.p, bits, nbytes, cudaMemcpyHostToDevice, stream));
        last.h2d += (int64_t)nbytes;
        return s_bitset.p;
    }
    // write [nq*k] results to the caller (device: results were produced in place)
    void
    results_out(int64_t nq, int k, int64_t* out_ids, float* out_dist, const int64_t* d_ids, const float* d_dist) {
        if (d_ids != out_ids) {
            KB2_CUDA_CHECK(cudaMemcpyAsync(out_ids, d_ids, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, stream));
            KB2_CUDA_CHECK(cudaMemcpyAsync(out_dist, d_dist, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
            last.d2h += nq * k * 12;
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    virtual void train(const float* x, int64_t n) = 0;
    virtual void add(const float* x, int64_t n, const int64_t* ids) = 0;
    virtual void search(const float* q, int64_t nq, int k, const JsonObj& cfg, const uint8_t* bitset, int64_t nbits,
                        int64_t* out_ids, float* out_dist) = 0;
    virtual int64_t count() const = 0;
    virtual int64_t bitset_rows() const { return count(); }   // rows a BitsetView must cover (whole index, also on a shard)
    virtual int64_t size_bytes() const = 0;
    virtual bool is_trained() const = 0;
    virtual bool has_raw() const = 0;
    virtual void get_vectors(const int64_t* ids, int64_t n, float* out) {
        throw Error(KB2_NOT_IMPLEMENTED, "GetVectorByIds not supported by this index");
    }
};

// ============================================================================================
// Dense candidate generation shared by FLAT and the IVF coarse quantizer:
//   partial[nq][S][Kout] <- per (query, base-slice) best Ksel approximate keys
// ============================================================================================
struct DensePlan {
    int Ksel = 32;
    int S = 1;        // slot capacity per query
    int used = 0;     // slots filled
    int64_t stride() const { return (int64_t)S * Ksel; }
};

inline DensePlan
dense_candidates(IndexBase& ix, const float* Q, int64_t nq, const float* X, const float* xn, int64_t n, int d,
                 int metric, int k_need, const uint8_t* bitset, const int32_t* rows, int64_t bit_offset = 0) {
    cudaStream_t st = ix.stream;
    DensePlan pl;
    pl.Ksel = next_pow2(std::max(32, k_need));
    KB2_REQUIRE(pl.Ksel <= kMaxK, KB2_INVALID_ARGS, "k too large for the GPU selection kernels (max 1008)");
    pl.S = std::max(2, kMaxSortEntries / pl.Ksel);
    ix.s_partial.ensure((size_t)nq * pl.stride());
    ix.s_qn.ensure(nq);
    if (metric == KB2_METRIC_L2) {
        row_norms_kernel<<<grid1d(nq * 32, 256), 256, 0, st>>>(Q, nq, d, ix.s_qn.p);
        ix.last.launches++;
    }
    const int64_t max_key_elems = 64ll << 20;  // 256 MB of keys
    int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1024, max_key_elems / std::max<int64_t>(nq, 1)));
    if (chunk < n) chunk = std::max<int64_t>(128, chunk / 128 * 128);
    const int64_t ldk = (chunk + 3) & ~(int64_t)3;   // 16-byte aligned key rows (vector stores in the epilogues)
    ix.s_keys.ensure((size_t)nq * ldk);
    int nsplit = (int)std::min<int64_t>(std::max<int64_t>(1, (2 * kNumSMs + nq - 1) / nq),
                                        std::max<int64_t>(1, chunk / 512));
    nsplit = std::min(nsplit, pl.S - 1);
    {
        // empty-entry fill of the slots this call can touch only.  (The whole [nq][S][Ksel] scratch used to be filled: 655 MB
        // per search at C3's coarse stage, where ONE 1 KB slot per query is used -- ~0.1 ms of a 2.4 ms step.)
        const int64_t n_chunks = (n + chunk - 1) / chunk;
        const int64_t slots = std::min<int64_t>(pl.S, n_chunks * nsplit);
        KB2_CUDA_CHECK(cudaMemset2DAsync(ix.s_partial.p, (size_t)pl.stride() * 8, 0xFF, (size_t)slots * pl.Ksel * 8, (size_t)nq, st));
    }
    const size_t sel_smem = (size_t)kScanWarps * 2 * pl.Ksel * 8;
    // chunk-minimum fast path of the wide select (KB2_SELECT_FAST=0: level-wise histogram only).  Measured at C3's coarse stage
    // (ncu, profiles/r2_summary.md): 0.218 ms -> 0.123 (128-bit loads) -> 0.087 (fast path)
    static const bool select_fast = [] { const char* e = getenv("KB2_SELECT_FAST"); return !(e && atoi(e) == 0); }();
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        const int64_t cols = std::min(chunk, n - c0);
        if (pl.used + nsplit > pl.S) {
            const int n_in = pl.used * pl.Ksel;
            const int n_sort = next_pow2(n_in);
            reduce_partials_kernel<<<(unsigned)nq, 256, (size_t)n_sort * 8, st>>>(ix.s_partial.p, (int)pl.stride(), n_in,
                                                                                 n_sort, pl.Ksel);
            ix.last.launches++;
            pl.used = 1;
        }
        launch_gemm_keys(st, gemm_mode(), metric, Q, X + c0 * d, ix.s_qn.p, xn + c0, (int)nq, (int)cols, d, ix.s_keys.p, ldk,
                         bitset, rows, c0 + bit_offset);
        const int per_slice = (int)(((cols + nsplit - 1) / nsplit + 31) / 32 * 32);
        const size_t hist_smem = (size_t)per_slice * 4 + 4160;
        if (pl.Ksel >= 64 && hist_smem <= (size_t)kMaxDynSmem) {
            select_keys_hist_kernel<<<dim3((unsigned)nq, nsplit), 256, hist_smem, st>>>(
                ix.s_keys.p, ldk, (int)cols, std::min(k_need, pl.Ksel), pl.Ksel, ix.s_partial.p, pl.S, pl.used, (uint32_t)c0,
                select_fast ? 1 : 0);
        } else {
            select_keys_kernel<<<dim3((unsigned)nq, nsplit), kScanThreads, sel_smem, st>>>(
                ix.s_keys.p, ldk, (int)cols, pl.Ksel, pl.Ksel, ix.s_partial.p, pl.S, pl.used, (uint32_t)c0);
        }
        ix.last.launches += 2;
        pl.used += nsplit;
    }
    KB2_CUDA_CHECK(cudaGetLastError());
    return pl;
}

inline void
launch_finalize(IndexBase& ix, FinalizeParams fp, int64_t nq) {
    fp.n_sort = next_pow2(std::max(fp.n_partial, 2));
    KB2_REQUIRE(fp.n_sort <= kMaxSortEntries, KB2_INTERNAL_ERROR, "finalize: too many partial candidates");
    KB2_REQUIRE(fp.k_sel <= kMaxK && fp.k_out <= fp.k_sel, KB2_INVALID_ARGS, "k too large");
    static const bool warp_path = [] { const char* e = getenv("KB2_FINALIZE"); return !(e && !strcmp(e, "cta")); }();
    if (warp_path && fp.k_sel <= 128 && fp.d <= 1024 && (fp.n_partial <= 256 || fp.counts)) {
        // one warp per query (see finalize_warp_kernel); variable-length rows longer than 256 entries fall through to the
        // CTA kernel below, which then skips the short ones (measured at C3: a 512-entry register sort for the tail costs
        // more than that second launch)
        const size_t smem_w = (size_t)kFinWarps * ((size_t)((fp.d + 3) & ~3) * 4 + 128 * 24);
        const unsigned g = (unsigned)((nq + kFinWarps - 1) / kFinWarps);
        if (fp.n_partial <= 128)
            finalize_warp_kernel<4><<<g, kFinWarps * 32, smem_w, ix.stream>>>(fp, nq);
        else
            finalize_warp_kernel<8><<<g, kFinWarps * 32, smem_w, ix.stream>>>(fp, nq);
        ix.last.launches++;
        KB2_CUDA_CHECK(cudaGetLastError());
        if (fp.n_partial <= 256) return;
        fp.split_small = 256;
    }
    const size_t smem = (size_t)fp.n_sort * 8 + (size_t)fp.k_sel * 16 + (size_t)fp.d * 4 + 16;
    unsigned grid = (unsigned)nq;
    if (fp.split_small > 0 && nq > 8 * kNumSMs) {   // tail pass: a few CTAs per SM walk the rows, most of which they skip
        grid = 8u * kNumSMs;
        fp.row_loop_nq = nq;
    }
    finalize_kernel<<<grid, 256, smem, ix.stream>>>(fp);
    ix.last.launches++;
    KB2_CUDA_CHECK(cudaGetLastError());
}

// ============================================================================================
// FLAT
// ============================================================================================
struct FlatIndex : IndexBase {
    DevBuf<float> base, norms;
    DevBuf<int64_t> labels;   // only when custom ids were given or the shard is offset
    size_t n_used = 0, norms_used = 0, labels_used = 0;
    bool custom_labels = false;
    int64_t n_global_added = 0;  // rows offered to add() over all calls (for sharding)
    int n_add_calls = 0;
    int64_t shard_lo = 0;        // first global row of this shard's slice (single add() call)
    int64_t bitset_rows() const override { return shard_world > 1 ? n_global_added : count(); }

    void train(const float*, int64_t) override {}
    bool is_trained() const override { return true; }
    bool has_raw() const override { return true; }
    int64_t count() const override { return (int64_t)(n_used / std::max(dim, 1)); }
    int64_t size_bytes() const override { return (int64_t)(n_used * 4 + norms_used * 4 + labels_used * 8); }

    void
    add(const float* x, int64_t n, const int64_t* ids) override {
        if (n <= 0) return;
        // sharding: this rank keeps the contiguous slice [lo, hi) of each add() call
        int64_t lo = 0, hi = n;
        if (shard_world > 1) {
            lo = n * shard_rank / shard_world;
            hi = n * (shard_rank + 1) / shard_world;
        }
        const int64_t m = hi - lo;
        const int64_t first_label = n_global_added + lo;
        if (n_add_calls++ == 0) shard_lo = lo;
        const bool need_labels = custom_labels || ids != nullptr || shard_world > 1;
        if (need_labels && !custom_labels) {
            // materialise identity labels for what is already stored
            const int64_t have = count();
            std::vector<int64_t> h(have);
            for (int64_t i = 0; i < have; i++) h[i] = i;
            labels_used = 0;
            if (have) dev_append(labels, labels_used, h.data(), (size_t)have, stream);
            KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
            custom_labels = true;
        }
        if (m > 0) {
            dev_append(base, n_used, x + lo * dim, (size_t)m * dim, stream);
            DevBuf<float> tmp;
            tmp.ensure(m);
            row_norms_kernel<<<grid1d(m * 32, 256), 256, 0, stream>>>(base.p + n_used - (size_t)m * dim, m, dim, tmp.p);
            dev_append(norms, norms_used, tmp.p, (size_t)m, stream);
            if (custom_labels) {
                std::vector<int64_t> h(m);
                if (ids) {
                    if (is_device_ptr(ids)) {
                        KB2_CUDA_CHECK(cudaMemcpyAsync(h.data(), ids + lo, m * 8, cudaMemcpyDeviceToHost, stream));
                        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
                    } else {
                        memcpy(h.data(), ids + lo, m * 8);
                    }
                } else {
                    for (int64_t i = 0; i < m; i++) h[i] = first_label + i;
                }
                dev_append(labels, labels_used, h.data(), (size_t)m, stream);
            }
            KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        n_global_added += n;
    }

    void
    search(const float* q, int64_t nq, int k, const JsonObj&, const uint8_t* bitset, int64_t nbits, int64_t* out_ids,
           float* out_dist) override {
        const int64_t n = count();
        KB2_REQUIRE(n > 0, KB2_EMPTY_INDEX, "index is empty");
        KB2_REQUIRE(k > 0 && k <= kMaxK - 16, KB2_INVALID_ARGS, "k out of range (1..1008)");
        const float* dq = to_device(q, (size_t)nq * dim, s_q);
        const uint8_t* dbits = bitset_to_device(bitset, nbits);
        const bool dev_out = is_device_ptr(out_ids);
        int64_t* d_ids = out_ids;
        float* d_dist = out_dist;
        if (!dev_out) {
            s_out_ids.ensure((size_t)nq * k);
            s_out_dist.ensure((size_t)nq * k);
            d_ids = s_out_ids.p;
            d_dist = s_out_dist.p;
        }
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev0, stream));
        // bitset indexes internal rows == labels when labels are the identity (like BitsetView over segment offsets);
        // a shard holds the contiguous slice [shard_lo, shard_lo + n) of ONE add() call, so bit = shard_lo + local row
        KB2_REQUIRE(!(dbits && shard_world > 1 && n_add_calls > 1), KB2_NOT_IMPLEMENTED,
                    "FLAT shard: bitset after several add() calls");
        DensePlan pl = dense_candidates(*this, dq, nq, base.p, norms.p, n, dim, metric, k + 16, dbits, nullptr,
                                        shard_world > 1 ? shard_lo : 0);
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev1, stream));
        FinalizeParams fp{};
        fp.partial = s_partial.p;
        fp.partial_stride = pl.stride();
        fp.n_partial = pl.used * pl.Ksel;
        fp.k_sel = std::min(pl.Ksel, k + 16);
        fp.k_out = k;
        fp.rows = nullptr;
        fp.labels = custom_labels ? labels.p : nullptr;
        fp.rerank = 1;
        fp.raw = base.p;
        fp.raw_by_pos = 1;
        fp.queries = dq;
        fp.d = dim;
        fp.metric = metric;
        fp.out_ids = d_ids;
        fp.out_dist = d_dist;
        fp.out_pos = nullptr;
        launch_finalize(*this, fp, nq);
        last.codes = nq * n;
        last.code_bytes = n * (int64_t)dim * 4;  // list-major contraction reads the base once per batch
        last.pairs = nq;
        results_out(nq, k, out_ids, out_dist, d_ids, d_dist);
        last_engine = 0;
        if (timing) {
            KB2_CUDA_CHECK(cudaEventElapsedTime(&last_stage_ms, ev0, ev1));
            last_kernel_ms = last_stage_ms;
        }
    }

    void
    get_vectors(const int64_t* ids, int64_t n, float* out) override {
        KB2_REQUIRE(!custom_labels, KB2_NOT_IMPLEMENTED, "GetVectorByIds with custom ids");
        std::vector<int64_t> h(n);
        if (is_device_ptr(ids)) {
            KB2_CUDA_CHECK(cudaMemcpy(h.data(), ids, n * 8, cudaMemcpyDeviceToHost));
        } else {
            memcpy(h.data(), ids, n * 8);
        }
        for (int64_t i = 0; i < n; i++) {
            KB2_REQUIRE(h[i] >= 0 && h[i] < count(), KB2_INVALID_ARGS, "id out of range");
            KB2_CUDA_CHECK(cudaMemcpyAsync(out + i * dim, base.p + h[i] * dim, (size_t)dim * 4, cudaMemcpyDefault, stream));
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
};

// ============================================================================================
// IVF_FLAT / IVF_PQ
// ============================================================================================
struct IvfIndex : IndexBase {
    bool is_pq = false;
    int64_t nlist = 128;
    int M = 0, nbits = 8, dsub = 0;
    bool refine = false;
    int refine_kind = 0;       // refine store element type: 0 fp32 ("flat"), 1 fp16, 2 bf16 (ivf_config.h:97-128)
    bool trained = false;
    // trained state
    DevBuf<float> centroids, cnorms, pqc;
    // flat (insertion-order) staging, valid while !sealed
    DevBuf<int32_t> f_assign;
    DevBuf<uint8_t> f_codes;
    DevBuf<float> f_vecs;
    DevBuf<int64_t> f_labels;
    size_t f_assign_used = 0, f_codes_used = 0, f_vecs_used = 0, f_labels_used = 0;
    bool custom_labels = false;
    int64_t n_total = 0;
    // sealed (list-order) layout
    bool sealed = false;
    int64_t npad = 0;
    int G = 0;                 // 16-sub-quantizer groups when the skewed kernel applies, else 0
    std::vector<int64_t> h_list_off;
    std::vector<int32_t> h_list_len, h_list_cnt_all, h_list_owner;
    DevBuf<int32_t> list_owner;   // [nlist] rank that holds each list (size-balanced packing, identical on every rank)
    DevBuf<int64_t> list_off;
    DevBuf<int32_t> list_len, rows, pos_of_row;
    DevBuf<uint8_t> codes;     // [G][npad][16] or [npad][M]
    DevBuf<uint16_t> vecs16;          // refine store when refine_kind != 0 (vecs is released after seal)
    DevBuf<float> t1, vecs, vnorm2;   // vnorm2[pos] = |x|^2 (IVF_FLAT: row term of the list-major tensor-core engine)
    DevBuf<int64_t> labels;    // row -> label (sealed copy of f_labels)
    DevBuf<int32_t> s_qkey, s_qkey2, s_qidx, s_qperm;
    DevBuf<uint8_t> s_sort_tmp;

    bool keeps_vecs() const { return !is_pq || refine; }
    bool is_trained() const override { return trained; }
    bool has_raw() const override { return keeps_vecs() && refine_kind == 0; }
    // fp32 view of the list-order vector store (decoded into tmp when it is kept in 16 bits)
    const float*
    vecs_f32(DevBuf<float>& tmp) {
        if (!refine_kind || !is_pq) return vecs.p;
        tmp.ensure((size_t)npad * dim);
        widen16_kernel<<<grid1d(npad * dim, 256), 256, 0, stream>>>(vecs16.p, npad * dim, refine_kind, tmp.p);
        return tmp.p;
    }
    int64_t count() const override { return n_total; }
    int64_t
    size_bytes() const override {
        return (int64_t)(centroids.bytes() + pqc.bytes() + codes.bytes() + t1.bytes() + vecs.bytes() + vecs16.bytes() + rows.bytes() +
                         pos_of_row.bytes() + f_codes.bytes() + f_vecs.bytes() + f_assign.bytes());
    }

    void
    set_centroids_common() {
        cnorms.ensure(nlist);
        row_norms_kernel<<<grid1d(nlist * 32, 256), 256, 0, stream>>>(centroids.p, nlist, dim, cnorms.p);
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    // ---------------------------------------------------------------- Train (ivf.cc:545-807)
    void
    train(const float* x, int64_t n) override {
        KB2_REQUIRE(!trained, KB2_INDEX_ALREADY_TRAINED, "index already trained");
        KB2_REQUIRE(n > 0, KB2_INVALID_ARGS, "empty training set");
        // MatchNlist (ivf.cc:479-489)
        if (nlist * 39 > n) nlist = std::max<int64_t>(1, n / 39);
        DevBuf<float> xbuf;
        const float* dx = to_device(x, (size_t)n * dim, xbuf, false);
        centroids.alloc_exact((size_t)nlist * dim);
        kmeans_train(dx, n, dim, (int)nlist, metric, 25, 1234, centroids.p, stream);
        set_centroids_common();
        if (is_pq) {
            KB2_REQUIRE(nbits == 8, KB2_NOT_IMPLEMENTED, "IVF_PQ: only nbits=8 is implemented on the GPU path");
            KB2_REQUIRE(M > 0 && dim % M == 0, KB2_INVALID_ARGS, "IVF_PQ: dim must be a multiple of m");
            KB2_REQUIRE(n >= 256, KB2_INVALID_ARGS, "IVF_PQ: need at least 256 training rows for nbits=8");
            dsub = dim / M;
            // residuals of (a subsample of) the training set: F/IndexIVF.cpp:1307-1329, IndexIVFPQ.cpp:76-95
            const int64_t nt = std::min<int64_t>(n, 256 * 256);
            DevBuf<float> sample;
            const float* xt = dx;
            if (nt < n) {
                std::mt19937_64 rng(1234 + 7);
                std::vector<int32_t> perm(n);
                for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
                for (int64_t i = 0; i < nt; i++) std::swap(perm[i], perm[i + (int64_t)(rng() % (uint64_t)(n - i))]);
                DevBuf<int32_t> didx;
                didx.ensure(nt);
                KB2_CUDA_CHECK(cudaMemcpyAsync(didx.p, perm.data(), nt * 4, cudaMemcpyHostToDevice, stream));
                sample.ensure((size_t)nt * dim);
                gather_rows_kernel<<<grid1d(nt * 32, 256), 256, 0, stream>>>(dx, didx.p, nt, dim, dim, sample.p);
                KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
                xt = sample.p;
            }
            DevBuf<int32_t> asg;
            asg.ensure(nt);
            AssignScratch sc;
            assign_nearest(xt, nt, dim, centroids.p, (int)nlist, metric, asg.p, nullptr, sc, stream);
            pqc.alloc_exact((size_t)M * 256 * dsub);
            tc_ready = false;
            DevBuf<float> sub;
            sub.ensure((size_t)nt * dsub);
            for (int m = 0; m < M; m++) {
                slice_residual_kernel<<<grid1d(nt * dsub, 256), 256, 0, stream>>>(xt, centroids.p, asg.p, nt, dim, m, dsub,
                                                                                sub.p);
                // every sub-quantizer is seeded identically, like the reference (one ClusteringParameters, seed 1234, for all
                // M Clustering objects: F/impl/ProductQuantizer.cpp:130-180) => the M codebooks start from the same 256 rows
                kmeans_train(sub.p, nt, dsub, 256, KB2_METRIC_L2, 25, 1234, pqc.p + (size_t)m * 256 * dsub, stream);
            }
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
        trained = true;
    }

    // ---------------------------------------------------------------- Add (ivf.cc:809-844; F/IndexIVF.cpp:212-287)
    void
    add(const float* x, int64_t n, const int64_t* ids) override {
        KB2_REQUIRE(trained, KB2_INDEX_NOT_TRAINED, "index not trained");
        if (n <= 0) return;
        if (sealed) unseal();
        DevBuf<float> xbuf;
        const float* dx = to_device(x, (size_t)n * dim, xbuf, false);
        DevBuf<int32_t> asg;
        asg.ensure(n);
        AssignScratch sc;
        assign_nearest(dx, n, dim, centroids.p, (int)nlist, metric, asg.p, nullptr, sc, stream);
        dev_append(f_assign, f_assign_used, asg.p, (size_t)n, stream);
        if (is_pq) {
            DevBuf<uint8_t> cb;
            cb.ensure((size_t)n * M);
            pq_encode_kernel<<<grid1d(n * 32, 256), 256, 0, stream>>>(dx, centroids.p, asg.p, pqc.p, n, dim, M, dsub, cb.p);
            dev_append(f_codes, f_codes_used, cb.p, (size_t)n * M, stream);
        }
        if (keeps_vecs()) dev_append(f_vecs, f_vecs_used, dx, (size_t)n * dim, stream);
        append_labels(ids, n);
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
        KB2_CUDA_CHECK(cudaGetLastError());
        n_total += n;
    }

    void
    append_labels(const int64_t* ids, int64_t n) {
        if (ids && !custom_labels) {
            std::vector<int64_t> h(n_total);
            for (int64_t i = 0; i < n_total; i++) h[i] = i;
            f_labels_used = 0;
            if (n_total) dev_append(f_labels, f_labels_used, h.data(), (size_t)n_total, stream);
            KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
            custom_labels = true;
        }
        if (custom_labels) {
            if (ids) {
                dev_append(f_labels, f_labels_used, ids, (size_t)n, stream);
            } else {
                std::vector<int64_t> h(n);
                for (int64_t i = 0; i < n; i++) h[i] = n_total + i;
                dev_append(f_labels, f_labels_used, h.data(), (size_t)n, stream);
            }
            KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
    }

    // ---------------------------------------------------------------- list-order layout
    void
    seal() {
        if (sealed) return;
        const int64_t n = n_total;
        cudaStream_t st = stream;
        // list sizes
        DevBuf<int32_t> dcnt;
        dcnt.ensure(nlist);
        KB2_CUDA_CHECK(cudaMemsetAsync(dcnt.p, 0, nlist * 4, st));
        if (n) histogram_kernel<<<grid1d(n, 256), 256, 0, st>>>(f_assign.p, n, dcnt.p);
        h_list_cnt_all.assign(nlist, 0);
        KB2_CUDA_CHECK(cudaMemcpyAsync(h_list_cnt_all.data(), dcnt.p, nlist * 4, cudaMemcpyDeviceToHost, st));
        KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        h_list_off.assign(nlist, 0);
        h_list_len.assign(nlist, 0);
        std::vector<int64_t> first_rank(nlist, 0);
        int64_t cur = 0, rank = 0;
        // list -> shard: greedy size-balanced packing (longest list first onto the lightest shard), computed from
        // the global list sizes, which every rank holds, so all ranks derive the same table.  KB2_SHARD_POLICY=mod: l % world.
        h_list_owner.assign(nlist, 0);
        if (shard_world > 1) {
            const char* pol = getenv("KB2_SHARD_POLICY");
            if (pol && !strcmp(pol, "mod")) {
                for (int64_t l = 0; l < nlist; l++) h_list_owner[l] = (int32_t)(l % shard_world);
            } else {
                std::vector<int64_t> order(nlist), load(shard_world, 0);
                for (int64_t l = 0; l < nlist; l++) order[l] = l;
                std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return h_list_cnt_all[a] > h_list_cnt_all[b]; });
                for (int64_t l : order) {
                    int best = 0;
                    for (int r = 1; r < shard_world; r++)
                        if (load[r] < load[best]) best = r;
                    h_list_owner[l] = best;
                    load[best] += h_list_cnt_all[l];
                }
            }
        }
        list_owner.alloc_exact(nlist);
        KB2_CUDA_CHECK(cudaMemcpyAsync(list_owner.p, h_list_owner.data(), nlist * 4, cudaMemcpyHostToDevice, st));
        for (int64_t l = 0; l < nlist; l++) {
            first_rank[l] = rank;
            rank += h_list_cnt_all[l];
            const bool owned = h_list_owner[l] == shard_rank;
            h_list_len[l] = owned ? h_list_cnt_all[l] : 0;
            h_list_off[l] = cur;
            cur += round_up(h_list_len[l], 32);
        }
        npad = cur + 32;
        KB2_REQUIRE(npad < (int64_t)0xfffffff0ll, KB2_INVALID_ARGS, "index too large for 32-bit positions");
        list_off.alloc_exact(nlist);
        list_len.alloc_exact(nlist);
        DevBuf<int64_t> d_first;
        d_first.ensure(nlist);
        KB2_CUDA_CHECK(cudaMemcpyAsync(list_off.p, h_list_off.data(), nlist * 8, cudaMemcpyHostToDevice, st));
        KB2_CUDA_CHECK(cudaMemcpyAsync(list_len.p, h_list_len.data(), nlist * 4, cudaMemcpyHostToDevice, st));
        KB2_CUDA_CHECK(cudaMemcpyAsync(d_first.p, first_rank.data(), nlist * 8, cudaMemcpyHostToDevice, st));
        // stable sort rows by list id
        rows.alloc_exact(npad);
        pos_of_row.alloc_exact(std::max<int64_t>(n, 1));
        fill_i32_kernel<<<grid1d(npad, 256), 256, 0, st>>>(rows.p, npad, -1);
        if (n) {
            DevBuf<int32_t> idx_in, idx_out, key_out;
            idx_in.ensure(n);
            idx_out.ensure(n);
            key_out.ensure(n);
            iota_kernel<<<grid1d(n, 256), 256, 0, st>>>(idx_in.p, n);
            size_t tmp_bytes = 0;
            int end_bit = 1;
            while ((1ll << end_bit) < nlist) end_bit++;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, f_assign.p, key_out.p, idx_in.p, idx_out.p, (int)n, 0,
                                            end_bit, st);
            DevBuf<uint8_t> tmp;
            tmp.ensure(tmp_bytes);
            cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, f_assign.p, key_out.p, idx_in.p, idx_out.p, (int)n, 0,
                                            end_bit, st);
            place_rows_kernel<<<grid1d(n, 256), 256, 0, st>>>(key_out.p, idx_out.p, n, d_first.p, list_off.p, list_len.p,
                                                            rows.p, pos_of_row.p);
            KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        }
        // payload in list order
        if (is_pq) {
            G = (M % 16 == 0 && M / 16 <= 3) ? M / 16 : 0;
            DevBuf<float> t1_flat;
            if (metric == KB2_METRIC_L2) {
                t1_flat.ensure(std::max<int64_t>(n, 1));
                if (n) pq_t1_kernel<<<grid1d(n * 32, 256), 256, 0, st>>>(f_codes.p, centroids.p, f_assign.p, pqc.p, n, dim, M,
                                                                         dsub, t1_flat.p);
                t1.alloc_exact(npad);
                gather_f32_kernel<<<grid1d(npad, 256), 256, 0, st>>>(t1_flat.p, rows.p, npad, t1.p, 0.f);
            }
            if (G > 0) {
                codes.alloc_exact((size_t)G * npad * 16);
                layout_codes_kernel<<<grid1d((int64_t)G * npad * 16, 256), 256, 0, st>>>(f_codes.p, rows.p, npad, M, G, codes.p);
            } else {
                codes.alloc_exact((size_t)npad * M);
                layout_codes_plain_kernel<<<grid1d(npad * M, 256), 256, 0, st>>>(f_codes.p, rows.p, npad, M, codes.p);
            }
        }
        if (keeps_vecs()) {
            vecs.alloc_exact((size_t)npad * dim);
            gather_rows_kernel<<<grid1d(npad * 32, 256), 256, 0, st>>>(f_vecs.p, rows.p, npad, dim, dim, vecs.p);
            if (!is_pq) {
                vnorm2.alloc_exact(npad);
                row_norms_kernel<<<grid1d(npad * 32, 256), 256, 0, st>>>(vecs.p, npad, dim, vnorm2.p);
            } else if (refine_kind) {
                vecs16.alloc_exact((size_t)npad * dim);
                narrow_kernel<<<grid1d(npad * dim, 256), 256, 0, st>>>(vecs.p, npad * dim, refine_kind, vecs16.p);
                KB2_CUDA_CHECK(cudaStreamSynchronize(st));
                vecs.release();
            }
        }
        if (custom_labels) {
            labels.alloc_exact(std::max<int64_t>(n, 1));
            KB2_CUDA_CHECK(cudaMemcpyAsync(labels.p, f_labels.p, n * 8, cudaMemcpyDeviceToDevice, st));
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        KB2_CUDA_CHECK(cudaGetLastError());
        // the insertion-order payload is no longer needed (assign/labels stay: small)
        f_codes.release();
        f_vecs.release();
        sealed = true;
    }

    // rebuild the insertion-order payload from the list-order one so that add() can append
    void
    unseal() {
        KB2_REQUIRE(shard_world == 1, KB2_NOT_IMPLEMENTED, "add() after search on a sharded index");
        const int64_t n = n_total;
        cudaStream_t st = stream;
        if (is_pq) {
            // gather the codes back into insertion order: one thread per (row, m)
            f_codes.alloc_exact((size_t)std::max<int64_t>(n, 1) * M);
            if (n) unlayout_codes_kernel<<<grid1d(n * M, 256), 256, 0, st>>>(codes.p, pos_of_row.p, n, npad, M, G, f_codes.p);
            f_codes_used = (size_t)n * M;
        }
        if (keeps_vecs()) {
            DevBuf<float> dec;
            const float* v32 = vecs_f32(dec);
            f_vecs.alloc_exact((size_t)std::max<int64_t>(n, 1) * dim);
            gather_rows_kernel<<<grid1d(n * 32, 256), 256, 0, st>>>(v32, pos_of_row.p, n, dim, dim, f_vecs.p);
            f_vecs_used = (size_t)n * dim;
            KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        sealed = false;
    }

    // ---------------------------------------------------------------- query-major scan launch (all IVF kinds)
    void
    launch_scan(IvfScanParams sp, unsigned grid, int Ksel, int np_max, bool has_bits) {
        cudaStream_t st = stream;
        const uint8_t* dbits = has_bits ? sp.bitset : nullptr;
        const size_t common_smem = (size_t)kScanWarps * 2 * Ksel * 8 + (size_t)(np_max + 1) * 4 + (size_t)np_max * 12 +
                                   (size_t)dim * 4 + 64 + 8 * (2 * kScanWarps + 4) + (size_t)4 * Ksel * 8;   // + CTA bound block + merge buffer
        if (is_pq) {
            if (G > 0) {
                const int scan_nt_env = [] { const char* e = getenv("KB2_SCAN_NT"); return e ? atoi(e) : 0; }();
                int scan_nt = (scan_nt_env == 256 || scan_nt_env == 512) ? scan_nt_env : KB2_DEFAULT_SCAN_NT;
                size_t smem = (size_t)G * 65536 + common_smem;
                if (scan_nt == 512) {
                    const size_t smem512 = smem + (size_t)kScanWarps * 2 * Ksel * 8 + 8 * kScanWarps;  // 16 warp buffers
                    if (smem512 <= (size_t)kMaxDynSmem) smem = smem512; else scan_nt = 256;
                }
                KB2_REQUIRE(smem <= (size_t)kMaxDynSmem, KB2_INVALID_ARGS, "IVF_PQ: k too large for shared memory");
#define KB2_LAUNCH_PQ_NT(GG, NTT)                                                                              \
    if (metric == KB2_METRIC_L2) {                                                                             \
        if (dbits) ivfpq_scan_kernel<GG, KB2_METRIC_L2, true, NTT><<<grid, NTT, smem, st>>>(sp);               \
        else ivfpq_scan_kernel<GG, KB2_METRIC_L2, false, NTT><<<grid, NTT, smem, st>>>(sp);                    \
    } else {                                                                                                   \
        if (dbits) ivfpq_scan_kernel<GG, KB2_METRIC_IP, true, NTT><<<grid, NTT, smem, st>>>(sp);               \
        else ivfpq_scan_kernel<GG, KB2_METRIC_IP, false, NTT><<<grid, NTT, smem, st>>>(sp);                    \
    }
#define KB2_LAUNCH_PQ(GG)                                     \
    if (scan_nt == 512) { KB2_LAUNCH_PQ_NT(GG, 512) } else { KB2_LAUNCH_PQ_NT(GG, 256) }
                const char* e_pf = getenv("KB2_SCAN_PREFETCH");
                sp.flags = (e_pf ? atoi(e_pf) : KB2_DEFAULT_SCAN_PREFETCH) ? 2 : 0;
                if (const char* e_fm = getenv("KB2_SCAN_FULLMERGE")) sp.flags |= atoi(e_fm) ? 4 : 0;
                if (G == 1) { KB2_LAUNCH_PQ(1) } else if (G == 2) { KB2_LAUNCH_PQ(2) } else { KB2_LAUNCH_PQ(3) }
#undef KB2_LAUNCH_PQ
#undef KB2_LAUNCH_PQ_NT
            } else {
                const size_t smem = (size_t)M * 1024 + common_smem;
                KB2_REQUIRE(smem <= (size_t)kMaxDynSmem, KB2_NOT_IMPLEMENTED, "IVF_PQ: m too large for the generic kernel");
                if (metric == KB2_METRIC_L2)
                    ivfpq_scan_generic_kernel<KB2_METRIC_L2><<<grid, kScanThreads, smem, st>>>(sp, codes.p);
                else
                    ivfpq_scan_generic_kernel<KB2_METRIC_IP><<<grid, kScanThreads, smem, st>>>(sp, codes.p);
            }
        } else {
            KB2_REQUIRE(dim % 4 == 0, KB2_NOT_IMPLEMENTED, "IVF_FLAT: dim must be a multiple of 4 on the GPU path");
            if (metric == KB2_METRIC_L2)
                ivfflat_scan_kernel<KB2_METRIC_L2><<<grid, kScanThreads, common_smem, st>>>(sp);
            else
                ivfflat_scan_kernel<KB2_METRIC_IP><<<grid, kScanThreads, common_smem, st>>>(sp);
        }
        last.launches++;
        KB2_CUDA_CHECK(cudaGetLastError());
    }

    // ---------------------------------------------------------------- list-major tensor-core engine (kb2_ivfpq_tc.cuh)
    static constexpr int kTcCandCap = 2048;     // survivor slots per query (overflow -> LUT kernel redoes the query)
    DevBuf<uint16_t> tc_pqc16, s_qb16;
    DevBuf<float> tc_pqc_t;   // codebook transposed for the in-kernel tables of bound_kernel<..., 3, 2>
    DevBuf<float> tc_maxn2, s_qnorm, s_pair_base, s_lut, s_bound;
    DevBuf<int32_t> s_lcount, s_lstart, s_items, s_pair_q, s_plan_out, s_flaglist, s_resp;
    DevBuf<uint64_t> s_cand;
    DevBuf<uint32_t> s_cand_cnt, s_logcnt;
    DevBuf<uint4> s_log;
    float tc_rmax = 0.f, tc_rowmax = 0.f;
    bool tc_ready = false;

    DevBuf<uint8_t> tc_codes_plain;   // un-rotated code bytes for the geometries whose decode assembles 16-byte chunks from several sub-quantizers
    // engine instances: <G=1, dsub=8> (m16 d128: C3) and <G=3, dsub=2> (m48 d96: C5)
    bool tc_geom_18() const { return G == 1 && M == 16 && dsub == 8; }
    bool tc_geom_32() const { return G == 3 && M == 48 && dsub == 2; }
    DevBuf<int32_t> s_items2, s_bal_idx, s_bal_idx2;
    DevBuf<uint32_t> s_bal_key, s_bal_key2;
    // Side stream of the list-major engine: the plan (pairs grouped by list, item table, cost sort: seven small, latency-bound
    // launches that depend on the coarse result only) runs beside phase A (which fills the SMs with 3 x 128 threads each) and
    // joins before the filter kernel.  KB2_TC_OVERLAP=0 keeps everything on the handle's stream.
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false;   // forked, not yet joined (only ever observed true after an exception)
    bool
    plan_overlap() {
        static const bool on = [] { const char* e = getenv("KB2_TC_OVERLAP"); return !(e && atoi(e) == 0); }();
        if (!on || getenv("KB2_TC_VERBOSE")) return false;
        if (!side_stream) {
            KB2_CUDA_CHECK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
            KB2_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            KB2_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        return true;
    }
    ~IvfIndex() override {
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side_stream) cudaStreamDestroy(side_stream);
    }
    // sort the items by descending cost and deal them to the G persistent CTAs in snake order; returns the new item arrays
    int32_t*
    balance_items(int32_t* items, int64_t max_items, int G_ctas, int tile_cost, int col_cost, cudaStream_t st = nullptr) {
        if (!st) st = stream;
        const char* e = getenv("KB2_TC_BALANCE");
        if (e && atoi(e) == 0) return items;
        s_items2.ensure((size_t)3 * max_items);
        s_bal_key.ensure((size_t)max_items); s_bal_key2.ensure((size_t)max_items);
        s_bal_idx.ensure((size_t)max_items); s_bal_idx2.ensure((size_t)max_items);
        pqtc::item_cost_kernel<<<grid1d(max_items, 256), 256, 0, st>>>(s_plan_out.p, items, items + 2 * max_items, list_len.p, max_items,
                                                                     tile_cost, col_cost, s_bal_key.p, s_bal_idx.p);
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s_bal_key.p, s_bal_key2.p, s_bal_idx.p, s_bal_idx2.p, (int)max_items, 0, 16, st);
        s_sort_tmp.ensure(tmp_bytes);
        cub::DeviceRadixSort::SortPairs(s_sort_tmp.p, tmp_bytes, s_bal_key.p, s_bal_key2.p, s_bal_idx.p, s_bal_idx2.p, (int)max_items, 0, 16, st);
        pqtc::deal_items_kernel<<<grid1d(max_items, 256), 256, 0, st>>>(s_plan_out.p, s_bal_idx2.p, G_ctas, items, items + max_items,
                                                                      items + 2 * max_items, s_items2.p, s_items2.p + max_items,
                                                                      s_items2.p + 2 * max_items);
        last.launches += 4;
        return s_items2.p;
    }

    static bool
    tc_dynamic_sched() {
        static const bool dyn = [] { const char* e = getenv("KB2_TC_SCHED"); return !(e && !strcmp(e, "static")); }();
        return dyn;
    }
    bool
    use_tc_engine(int64_t nq, int nprobe, int Ksel) const {
        if (!is_pq || !(tc_geom_18() || tc_geom_32())) return false;
        const char* e = getenv("KB2_PQ_ENGINE");
        if (e && !strcmp(e, "lut")) return false;
        if (nprobe < 8 || Ksel > kTcCandCap / 2) return false;
        if (e && !strcmp(e, "tc")) return true;
        // the decode of a list is amortised over the queries that probe it.  Measured at C5 (100M x 96, nlist 65536, 19.5
        // queries per list on average): the query-major LUT engine needs 34.5 ms per 10000-query batch on two GPUs, so the
        // list-major engine is taken from 8 queries per list on (KB2_TC_MIN_QPL overrides)
        static const double min_qpl = [] { const char* t = getenv("KB2_TC_MIN_QPL"); return t ? atof(t) : 8.0; }();
        return (double)nq * nprobe >= min_qpl * (double)nlist;
    }

    void
    search_tc(IvfScanParams sp, int64_t nq, int nprobe, int Ksel, int k_base, bool has_bits) {
        cudaStream_t st = stream;
        const bool dist = distributed();
        if (!tc_ready) {
            tc_pqc16.alloc_exact((size_t)M * 256 * dsub);
            tc_maxn2.alloc_exact(M + 4);
            KB2_CUDA_CHECK(cudaMemsetAsync(tc_maxn2.p + M, 0, 16, st));
            pqtc::prepare_tables_kernel<<<M, 256, 0, st>>>(pqc.p, dsub, (__nv_bfloat16*)tc_pqc16.p, tc_maxn2.p);
            if (tc_geom_32()) {
                tc_pqc_t.alloc_exact((size_t)M * 256 * dsub);
                pqtc::transpose_codebook_kernel<<<grid1d((int64_t)M * 256 * dsub, 256), 256, 0, st>>>(pqc.p, M, dsub, tc_pqc_t.p);
            }
            if (metric == KB2_METRIC_L2 && npad > 0)
                pqtc::max_abs_kernel<<<kNumSMs * 2, 256, 0, st>>>(t1.p, npad, (uint32_t*)(tc_maxn2.p + M));
            if (dsub < 8) {
                tc_codes_plain.alloc_exact((size_t)G * npad * 16);
                pqtc::unrotate_codes_kernel<<<grid1d((int64_t)G * npad * 16, 256), 256, 0, st>>>(codes.p, (int64_t)G * npad, npad,
                                                                                            tc_codes_plain.p);
            }
            std::vector<float> h(M + 4);
            KB2_CUDA_CHECK(cudaMemcpyAsync(h.data(), tc_maxn2.p, (M + 4) * 4, cudaMemcpyDeviceToHost, st));
            KB2_CUDA_CHECK(cudaStreamSynchronize(st));
            double acc = 0;
            for (int i = 0; i < M; i++) acc += h[i];
            tc_rmax = (float)std::sqrt(acc) * 1.0001f;
            tc_rowmax = 0.5f * h[M] * 1.0001f;   // max |t1| / 2: the largest row term of the admission test
            tc_ready = true;
        }
        const int64_t npairs = nq * nprobe;
        // development aid: KB2_TC_VERBOSE=1 prints the device time of every stage of this engine
        const bool verbose = getenv("KB2_TC_VERBOSE") != nullptr;
        std::vector<std::pair<const char*, cudaEvent_t>> marks;
        auto mark = [&](const char* name) {
            if (!verbose) return;
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            marks.emplace_back(name, e);
        };
        mark("start");
        // the plan below depends on the coarse result only: with the side stream it runs beside phase A
        const bool ov = plan_overlap();
        cudaStream_t ps = ov ? side_stream : st;
        if (ov) {
            if (side_pending) {
                // a previous call left between fork and join (an error was thrown): drain its side-stream work before the
                // scratch buffers are reused
                KB2_CUDA_CHECK(cudaStreamSynchronize(side_stream));
                side_pending = false;
            }
            KB2_CUDA_CHECK(cudaEventRecord(ev_fork, st));
            KB2_CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
            side_pending = true;
        }
        // ---- phase A: exact scan of each query's nearest lists -> upper bound of its k_base-th best key.  A bound taken from ANY
        //      subset of the codes is valid on every rank, so with a communicator the query is handled by the rank that owns
        //      its nearest list (1/world of the batch each; tables only for those) and the bounds are min-reduced.
        const char* e_p0 = getenv("KB2_TC_P0");
        const char* e_ac = getenv("KB2_TC_A_CODES");
        const int p0 = std::max(1, std::min((e_p0 ? atoi(e_p0) : 8) * std::max(1, shard_world), nprobe));   // at most this many lists
        const int a_codes = e_ac ? atoi(e_ac) : 3000;                                                        // ... until this many codes
        s_bound.ensure((size_t)nq);
        static const bool generic_a = [] { const char* e = getenv("KB2_TC_PHASE_A"); return e && !strcmp(e, "scan"); }();
        if (tc_geom_18() || (tc_geom_32() && !generic_a)) {
        if (tc_geom_18()) s_lut.ensure((size_t)nq * 4096);
        const int32_t* qlist = nullptr;
        const uint32_t* qcount = nullptr;
        unsigned bound_grid = (unsigned)nq;
        if (dist) {
            s_resp.ensure((size_t)nq + 1);
            pqtc::compact_resp_kernel<<<1, 1024, 0, st>>>(sp.probe_ids, nprobe, nq, list_owner.p, shard_rank, s_resp.p + 1,
                                                          (uint32_t*)s_resp.p);
            pqtc::fill_f32_kernel<<<grid1d(nq, 256), 256, 0, st>>>(s_bound.p, nq, INFINITY);
            qlist = s_resp.p + 1;
            qcount = (const uint32_t*)s_resp.p;
            bound_grid = (unsigned)std::min<int64_t>(nq, std::max<int64_t>(4 * kNumSMs, 2 * nq / shard_world));
            last.launches += 2;
        }
        {
            const char* e_rw = getenv("KB2_BOUND_ROWW");
            const int roww = (e_rw && atoi(e_rw) == 64) ? 64 : 32;   // measured at C3: 0.37 ms (32, 3 CTAs/SM) vs 0.52 ms (64, 2 CTAs/SM)
            const size_t smem = pqtc::bound_smem(roww, pqtc::bound_kmax(a_codes, k_base));
            // measured at C3 (profiles/r2_summary.md): 0.274 ms with 128 threads, 0.234 ms with 256
            static const int bound_nt = [] { const char* e = getenv("KB2_BOUND_NT"); return e ? atoi(e) : 256; }();
#define KB2_BOUND_LAUNCH(MM, RW)                                                                                                    \
    pqtc::lut_build_kernel<MM><<<kNumSMs, 256, 0, st>>>(sp.queries, nq, qlist, qcount, pqc.p, s_lut.p);                             \
    mark("lut");                                                                                                                    \
    pqtc::bound_kernel<MM, RW><<<bound_grid, 128, smem, st>>>(s_lut.p, qlist, qcount, nq, sp.probe_ids, sp.probe_dis, nprobe, p0,   \
                                                             a_codes, k_base, list_off.p, list_len.p, (const uint4*)codes.p, t1.p, \
                                                             sp.bitset, rows.p, s_bound.p, d_counter.p + 4);
            if (tc_geom_32()) {
                // m48 x dsub2: three groups through one in-kernel table each (no [nq][m][256] table in global memory)
#define KB2_BOUND_LAUNCH3(MM)                                                                                                       \
    pqtc::bound_kernel<MM, 32, 3, 2><<<bound_grid, 128, pqtc::bound_smem(32, pqtc::bound_kmax(a_codes, k_base)), st>>>(             \
        nullptr, qlist, qcount, nq, sp.probe_ids, sp.probe_dis, nprobe, p0, a_codes, k_base, list_off.p, list_len.p,                \
        (const uint4*)codes.p, t1.p, sp.bitset, rows.p, s_bound.p, d_counter.p + 4, npad, sp.queries, tc_pqc_t.p);
                if (metric == KB2_METRIC_L2) { KB2_BOUND_LAUNCH3(KB2_METRIC_L2) } else { KB2_BOUND_LAUNCH3(KB2_METRIC_IP) }
#undef KB2_BOUND_LAUNCH3
            } else if (bound_nt == 256 && roww == 32) {
                // 8 warps per CTA over the same tables (default; KB2_BOUND_NT=128 for the 4-warp instance)
#define KB2_BOUND_LAUNCH256(MM)                                                                                                     \
    pqtc::lut_build_kernel<MM><<<kNumSMs, 256, 0, st>>>(sp.queries, nq, qlist, qcount, pqc.p, s_lut.p);                             \
    mark("lut");                                                                                                                    \
    pqtc::bound_kernel<MM, 32, 1, 8, 256><<<bound_grid, 256, smem, st>>>(s_lut.p, qlist, qcount, nq, sp.probe_ids, sp.probe_dis,     \
                                                                         nprobe, p0, a_codes, k_base, list_off.p, list_len.p,       \
                                                                         (const uint4*)codes.p, t1.p, sp.bitset, rows.p, s_bound.p, \
                                                                         d_counter.p + 4);
                if (metric == KB2_METRIC_L2) { KB2_BOUND_LAUNCH256(KB2_METRIC_L2) } else { KB2_BOUND_LAUNCH256(KB2_METRIC_IP) }
#undef KB2_BOUND_LAUNCH256
            } else if (metric == KB2_METRIC_L2) {
                if (roww == 32) { KB2_BOUND_LAUNCH(KB2_METRIC_L2, 32) } else { KB2_BOUND_LAUNCH(KB2_METRIC_L2, 64) }
            } else {
                if (roww == 32) { KB2_BOUND_LAUNCH(KB2_METRIC_IP, 32) } else { KB2_BOUND_LAUNCH(KB2_METRIC_IP, 64) }
            }
#undef KB2_BOUND_LAUNCH
            KB2_CUDA_CHECK(cudaGetLastError());
            last.launches += 2;
        }
        } else {
            // other geometries: the query-major LUT kernel over the first probes gives the exact k_base-th best key of those
            // lists (any subset of the codes yields a valid bound)
            int64_t avg_len = std::max<int64_t>(1, n_total / std::max<int64_t>(1, nlist));
            const int pA = (int)std::min<int64_t>(nprobe, std::max<int64_t>(2, (a_codes + avg_len - 1) / avg_len + 1) * std::max(1, shard_world));
            IvfScanParams a = sp;
            a.nprobe = pA;
            a.probe_stride = nprobe;
            a.nsplit = 1;
            a.partial = s_partial2.p;
            a.partial_stride = 0;
            a.counters = d_counter.p + 4;
            a.qperm = nullptr;
            launch_scan(a, (unsigned)nq, Ksel, pA, has_bits);
            fltc::extract_bound_kernel<<<grid1d(nq, 256), 256, 0, st>>>(s_partial2.p, Ksel, k_base, nq, s_bound.p);
            last.launches += 1;
        }
        if (dist) comm->all_reduce_min_f32(s_bound.p, s_bound.p, (size_t)nq, st);
        mark("phaseA");
        // ---- plan: pairs grouped by list, work items
        const int64_t max_items = nlist + npairs / pqtc::NQT + 2;
        s_lcount.ensure((size_t)2 * nlist);
        s_lstart.ensure((size_t)nlist);
        s_items.ensure((size_t)3 * max_items);
        s_plan_out.ensure(8);
        s_pair_q.ensure((size_t)npairs);
        s_pair_base.ensure((size_t)npairs);
        s_qb16.ensure((size_t)nq * dim);
        s_qnorm.ensure((size_t)nq);
        s_cand.ensure((size_t)nq * kTcCandCap);
        s_cand_cnt.ensure((size_t)2 * nq);
        KB2_CUDA_CHECK(cudaMemsetAsync(s_lcount.p, 0, (size_t)2 * nlist * 4, ps));
        KB2_CUDA_CHECK(cudaMemsetAsync(s_cand_cnt.p, 0, (size_t)2 * nq * 4, ps));
        pqtc::count_pairs_kernel<<<grid1d(npairs, 256), 256, 0, ps>>>(sp.probe_ids, npairs, list_len.p, s_lcount.p);
        int32_t* item_list = s_items.p;
        int32_t* item_q0 = s_items.p + max_items;
        int32_t* item_nq = s_items.p + 2 * max_items;
        pqtc::plan_kernel<<<1, 1024, 0, ps>>>(s_lcount.p, (int)nlist, s_lstart.p, item_list, item_q0, item_nq, s_plan_out.p);
        {
            // per tile: decode ~ constant, contraction ~ columns (+ the test K-step)
            // dynamic draw (default): items in descending cost order, CTAs take the next one when free; KB2_TC_SCHED=static
            // keeps the fixed round-robin assignment with the snake deal
            int32_t* bal = balance_items(s_items.p, max_items, tc_dynamic_sched() ? 0 : kNumSMs, 600, 5, ps);
            item_list = bal;
            item_q0 = bal + max_items;
            item_nq = bal + 2 * max_items;
        }
        pqtc::fill_pairs_kernel<<<grid1d(npairs, 256), 256, 0, ps>>>(sp.probe_ids, sp.probe_dis, npairs, nprobe, metric, list_len.p,
                                                                     s_lstart.p, s_lcount.p + nlist, s_pair_q.p, s_pair_base.p);
        pqtc::prepare_queries_kernel<<<grid1d(nq * 32, 256), 256, 0, ps>>>(sp.queries, nq, dim, (__nv_bfloat16*)s_qb16.p, s_qnorm.p);
        if (ov) {
            KB2_CUDA_CHECK(cudaGetLastError());
            KB2_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
            KB2_CUDA_CHECK(cudaStreamWaitEvent(st, ev_join, 0));
            side_pending = false;
        }
        mark("plan");
        // ---- tensor-core filter + exact re-evaluation of the survivors
        pqtc::Params tp{};
        tp.metric = metric;
        tp.nq = (int)nq;
        tp.nprobe = nprobe;
        tp.queries = sp.queries;
        tp.qb16 = (const __nv_bfloat16*)s_qb16.p;
        tp.qnorm = s_qnorm.p;
        tp.n_items = s_plan_out.p;
        if (tc_dynamic_sched()) {
            tp.ticket = s_plan_out.p + 4;
            KB2_CUDA_CHECK(cudaMemsetAsync(tp.ticket, 0, 4, st));
        }
        tp.item_list = item_list;
        tp.item_q0 = item_q0;
        tp.item_nq = item_nq;
        tp.pair_q = s_pair_q.p;
        tp.pair_base = s_pair_base.p;
        tp.bound = s_bound.p;
        tp.k_need = k_base;
        tp.margin_coef = (metric == KB2_METRIC_L2 ? 2.f : 1.f) * pqtc::kErrCoef * tc_rmax;
        tp.rmax = (metric == KB2_METRIC_L2) ? tc_rowmax : 0.f;
        tp.list_off = list_off.p;
        tp.list_len = list_len.p;
        tp.codes = (const uint4*)codes.p;
        tp.codes_plain = (const uint4*)tc_codes_plain.p;
        tp.npad = npad;
        tp.t1 = t1.p;
        tp.pqc = pqc.p;
        tp.pqc16 = (const uint4*)tc_pqc16.p;
        tp.bitset = sp.bitset;
        tp.rows = rows.p;
        const int n_logs = 2 * kNumSMs;   // one per epilogue group (+ 1 legacy slot that stays empty)
        const uint32_t log_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(nq * 1024 / n_logs, 16384), 1 << 19);
        s_log.ensure((size_t)(n_logs + 1) * log_cap);
        s_logcnt.ensure(n_logs + 8);
        KB2_CUDA_CHECK(cudaMemsetAsync(s_logcnt.p, 0, (n_logs + 8) * 4, st));
        tp.log = s_log.p;
        tp.log_cnt = s_logcnt.p;
        tp.log_cap = log_cap;
        tp.shared_cap = log_cap;
        tp.qflag = s_cand_cnt.p + nq;
        tp.counters = d_counter.p;
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev2, st));
#define KB2_TC_LAUNCH(GG, DD)                                                                                                         \
    if (metric == KB2_METRIC_L2)                                                                                                     \
        pqtc::ivfpq_tc_filter_kernel<KB2_METRIC_L2, GG, DD><<<kNumSMs, pqtc::THREADS, pqtc::TcCfg<GG, DD>::SMEM_BYTES, st>>>(tp);    \
    else                                                                                                                             \
        pqtc::ivfpq_tc_filter_kernel<KB2_METRIC_IP, GG, DD><<<kNumSMs, pqtc::THREADS, pqtc::TcCfg<GG, DD>::SMEM_BYTES, st>>>(tp);
        if (tc_geom_18()) { KB2_TC_LAUNCH(1, 8) } else { KB2_TC_LAUNCH(3, 2) }
#undef KB2_TC_LAUNCH
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev3, st));
        KB2_CUDA_CHECK(cudaGetLastError());
        mark("tc_filter");
        // ---- survivors: group by query, exact fp32 keys (bit-identical to the LUT engine's)
        pqtc::scatter_survivors_kernel<<<dim3(16, n_logs + 1), 256, 0, st>>>(s_log.p, s_logcnt.p, log_cap, log_cap, s_cand.p, s_cand_cnt.p,
                                                                         kTcCandCap, tp.qflag, d_counter.p);
#define KB2_TC_EVAL(MM, GG, DD)                                                                                                      \
    pqtc::exact_eval_kernel<MM, GG, DD><<<(unsigned)nq, 128, 0, st>>>(sp.queries, pqc.p, eval_lut, s_bound.p, (const uint4*)codes.p, npad, t1.p,  \
                                                                      sp.bitset, rows.p, s_cand.p, s_cand_cnt.p, kTcCandCap, tp.qflag,  \
                                                                      s_logcnt.p + n_logs + 1, eval_trim ? k_base : 0);
        // trim every survivor row to its k' best inside exact_eval (KB2_EVAL_TRIM=0: off).  Measured at C3: step 1.977 -> 1.957 ms
        static const bool eval_trim = [] { const char* e = getenv("KB2_EVAL_TRIM"); return !(e && atoi(e) == 0); }();
        const float* eval_lut = (tc_geom_18() && !dist) ? s_lut.p : nullptr;   // tables of the whole batch exist only without a communicator
        if (tc_geom_18()) {
            if (metric == KB2_METRIC_L2) { KB2_TC_EVAL(KB2_METRIC_L2, 1, 8) } else { KB2_TC_EVAL(KB2_METRIC_IP, 1, 8) }
        } else {
            if (metric == KB2_METRIC_L2) { KB2_TC_EVAL(KB2_METRIC_L2, 3, 2) } else { KB2_TC_EVAL(KB2_METRIC_IP, 3, 2) }
        }
#undef KB2_TC_EVAL
        KB2_CUDA_CHECK(cudaGetLastError());
        last.launches += 7;
        mark("scatter+eval");
        // ---- flagged queries (no bound / buffer overflow): complete LUT scan into their candidate rows
        {
            IvfScanParams f = sp;
            f.nsplit = std::max(1, std::min(kTcCandCap / Ksel, nprobe));   // probe slices per flagged query: their lists fill the row
            f.partial = s_cand.p;
            f.partial_stride = kTcCandCap;
            f.clear_to = kTcCandCap - (f.nsplit - 1) * Ksel;   // == Ksel (nothing to clear) when the slices fill the row
            s_flaglist.ensure((size_t)nq + 1);
            pqtc::compact_flags_kernel<<<1, 1024, 0, st>>>(tp.qflag, nq, s_flaglist.p + 1, (uint32_t*)s_flaglist.p);
            last.launches++;
            f.only_flagged = tp.qflag;
            f.flag_list = s_flaglist.p + 1;
            f.flag_count = (const uint32_t*)s_flaglist.p;
            f.qperm = nullptr;
            f.lut_global = nullptr;   // the redo pass builds its own tables (a handful of queries)
            f.counters = d_counter.p + 4;
            launch_scan(f, (unsigned)std::min<int64_t>(nq * f.nsplit, 3 * kNumSMs), Ksel, (nprobe + f.nsplit - 1) / f.nsplit, has_bits);
        }
        mark("fallback");
        if (verbose) {
            KB2_CUDA_CHECK(cudaStreamSynchronize(st));
            fprintf(stderr, "[kb2 tc]");
            for (size_t i = 1; i < marks.size(); i++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
                fprintf(stderr, " %s %.3f ms |", marks[i].first, ms);
            }
            uint32_t hn = 0;
            cudaMemcpy(&hn, s_plan_out.p, 4, cudaMemcpyDeviceToHost);
            unsigned long long hc2[8];
            cudaMemcpy(hc2, d_counter.p, 64, cudaMemcpyDeviceToHost);
            fprintf(stderr, " items %u survivors %llu flagged %llu (row overflows %llu, no bound %llu)\n", hn, hc2[2], hc2[7],
                    hc2[6] & 0xffffffffull, hc2[6] >> 32);
            for (auto& m : marks) cudaEventDestroy(m.second);
        }
    }

    // ---------------------------------------------------------------- IVF_FLAT list-major tensor-core engine (kb2_ivfflat_tc.cuh)
    DevBuf<float> s_qhi, s_qlo;
    bool
    use_flat_tc_engine(int64_t nq, int nprobe, int k) const {
        if (is_pq || dim % fltc::BK != 0 || dim < fltc::BK || npad >= (1ll << 31)) return false;
        const char* e = getenv("KB2_FLAT_ENGINE");
        if (e && !strcmp(e, "scan")) return false;
        if (k + 16 > kTcCandCap / 2) return false;
        if (e && !strcmp(e, "tc")) return true;
        // a list tile is amortised over the queries probing it
        return (double)nq * nprobe >= 8.0 * (double)nlist && nq >= 64;
    }

    // returns false when some query overflowed its candidate row (caller falls back to the query-major scan)
    bool
    search_flat_tc(IvfScanParams sp, int64_t nq, int nprobe, int Ksel, int k, bool has_bits) {
        cudaStream_t st = stream;
        const int64_t npairs = nq * nprobe;
        const int64_t npairs_pad = npairs + fltc::NQ_ITEM;
        // items of <= 32 queries (small B tiles, 5 stages in flight) while a list is probed by few queries, else <= 128
        const int item_cap = ((double)npairs / (double)std::max<int64_t>(1, nlist) <= 40.0) ? 32 : 128;
        // ---- phase A: exact k-th best key over the query's nearest probed lists (query-major kernel) = admission bound
        const int pA = std::min(nprobe, std::max(1, 2 * shard_world));
        s_bound.ensure((size_t)nq);
        {
            IvfScanParams a = sp;
            a.nprobe = pA;
            a.probe_stride = nprobe;
            a.nsplit = 1;
            a.partial = s_partial2.p;
            a.partial_stride = 0;
            a.counters = nullptr;
            a.qperm = nullptr;
            launch_scan(a, (unsigned)nq, Ksel, pA, has_bits);
            fltc::extract_bound_kernel<<<grid1d(nq, 256), 256, 0, st>>>(s_partial2.p, Ksel, k, nq, s_bound.p);
            if (distributed()) comm->all_reduce_min_f32(s_bound.p, s_bound.p, (size_t)nq, st);
            last.launches += 1;
        }
        // ---- plan: pairs grouped by list, items of <= 128 queries
        const int64_t max_items = nlist + npairs / item_cap + 2;
        s_lcount.ensure((size_t)2 * nlist);
        s_lstart.ensure((size_t)nlist);
        s_items.ensure((size_t)3 * max_items);
        s_plan_out.ensure(8);
        s_pair_q.ensure((size_t)npairs);
        s_pair_base.ensure((size_t)npairs);
        s_qnorm.ensure((size_t)nq);
        s_cand.ensure((size_t)nq * kTcCandCap);
        s_cand_cnt.ensure((size_t)2 * nq + 4);
        s_qhi.ensure((size_t)npairs_pad * dim);
        s_qlo.ensure((size_t)npairs_pad * dim);
        KB2_CUDA_CHECK(cudaMemsetAsync(s_lcount.p, 0, (size_t)2 * nlist * 4, st));
        KB2_CUDA_CHECK(cudaMemsetAsync(s_cand_cnt.p, 0, ((size_t)2 * nq + 4) * 4, st));
        pqtc::count_pairs_kernel<<<grid1d(npairs, 256), 256, 0, st>>>(sp.probe_ids, npairs, list_len.p, s_lcount.p);
        int32_t* item_list = s_items.p;
        int32_t* item_q0 = s_items.p + max_items;
        int32_t* item_nq = s_items.p + 2 * max_items;
        fltc::plan_kernel<<<1, 1024, 0, st>>>(s_lcount.p, (int)nlist, item_cap, s_lstart.p, item_list, item_q0, item_nq, s_plan_out.p);
        {
            int32_t* bal = balance_items(s_items.p, max_items, tc_dynamic_sched() ? 0 : kNumSMs, 1000, 2);   // a tile is bound by its HBM stream
            item_list = bal;
            item_q0 = bal + max_items;
            item_nq = bal + 2 * max_items;
        }
        pqtc::fill_pairs_kernel<<<grid1d(npairs, 256), 256, 0, st>>>(sp.probe_ids, sp.probe_dis, npairs, nprobe, metric, list_len.p,
                                                                     s_lstart.p, s_lcount.p + nlist, s_pair_q.p, s_pair_base.p);
        // pairs of lists owned by other shards leave holes at the end of the pair array: point them at no query
        fltc::gather_split_queries_kernel<<<grid1d(npairs_pad * 32, 256), 256, 0, st>>>(sp.queries, s_pair_q.p, npairs, npairs_pad, dim,
                                                                                     s_qhi.p, s_qlo.p);
        row_norms_kernel<<<grid1d(nq * 32, 256), 256, 0, st>>>(sp.queries, nq, dim, s_qnorm.p);
        CUtensorMap tx, thi, tlo;
        KB2_REQUIRE(tc::make_tmap(&tx, vecs.p, npad, dim) && tc::make_tmap(&thi, s_qhi.p, npairs_pad, dim, item_cap) &&
                        tc::make_tmap(&tlo, s_qlo.p, npairs_pad, dim, item_cap),
                    KB2_INTERNAL_ERROR, "IVF_FLAT tensor-core engine: tensor map encoding failed");
        fltc::Params fpar{};
        fpar.metric = metric;
        fpar.d = dim;
        fpar.n_items = s_plan_out.p;
        if (tc_dynamic_sched()) {
            fpar.ticket = s_plan_out.p + 4;
            KB2_CUDA_CHECK(cudaMemsetAsync(fpar.ticket, 0, 4, st));
        }
        fpar.item_list = item_list;
        fpar.item_q0 = item_q0;
        fpar.item_nq = item_nq;
        fpar.pair_q = s_pair_q.p;
        fpar.qnorm2 = s_qnorm.p;
        fpar.bound = s_bound.p;
        fpar.list_off = list_off.p;
        fpar.list_len = list_len.p;
        fpar.xnorm2 = vnorm2.p;
        fpar.bitset = sp.bitset;
        fpar.rows = rows.p;
        const uint32_t log_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(nq * 1024 / kNumSMs, 32768), 1 << 20);
        s_log.ensure((size_t)kNumSMs * log_cap);
        s_logcnt.ensure(kNumSMs + 8);
        KB2_CUDA_CHECK(cudaMemsetAsync(s_logcnt.p, 0, (kNumSMs + 8) * 4, st));
        fpar.log = s_log.p;
        fpar.log_cnt = s_logcnt.p;
        fpar.log_cap = log_cap;
        fpar.counters = d_counter.p;
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev2, st));
#define KB2_FL_LAUNCH(MM, BR) \
    fltc::ivfflat_tc_kernel<MM, BR><<<kNumSMs, fltc::THREADS, fltc::FlCfg<BR>::SMEM_BYTES, st>>>(tx, thi, tlo, fpar);
        if (metric == KB2_METRIC_L2) {
            if (item_cap == 32) { KB2_FL_LAUNCH(KB2_METRIC_L2, 32) } else { KB2_FL_LAUNCH(KB2_METRIC_L2, 128) }
        } else {
            if (item_cap == 32) { KB2_FL_LAUNCH(KB2_METRIC_IP, 32) } else { KB2_FL_LAUNCH(KB2_METRIC_IP, 128) }
        }
#undef KB2_FL_LAUNCH
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev3, st));
        KB2_CUDA_CHECK(cudaGetLastError());
        uint32_t* qflag = s_cand_cnt.p + nq;
        fltc::scatter_kernel<<<dim3(16, kNumSMs), 256, 0, st>>>(s_log.p, s_logcnt.p, log_cap, s_cand.p, s_cand_cnt.p, kTcCandCap, qflag);
        fltc::count_flags_kernel<<<grid1d(nq, 256), 256, 0, st>>>(qflag, nq, s_logcnt.p + kNumSMs, s_cand_cnt.p + 2 * nq);
        last.launches += 9;
        uint32_t* hflag = (uint32_t*)h_counter.p + 12;
        KB2_CUDA_CHECK(cudaMemcpyAsync(hflag, s_cand_cnt.p + 2 * nq, 4, cudaMemcpyDeviceToHost, st));
        KB2_CUDA_CHECK(cudaStreamSynchronize(st));
        last.flagged = hflag[0];
        return hflag[0] == 0;
    }

    // coarse quantizer for queries [q_lo, q_hi): top-nprobe centroids with exact dis0 (F/IndexIVF.cpp:336-342) into rows
    // [q_lo, q_hi) of s_probe_ids / s_probe_dis
    // Coarse stage on the list-major tensor-core kernel: the centroid table is ONE pseudo-list scanned by every query, i.e. an
    // IVF_FLAT search with k = nprobe + 16.  The admission bound comes from a sample (the first nlist/8 centroids through the
    // dense path: its k'-th best key, k' = 2x the expected share + 8), so the kernel logs ~2(nprobe+16) candidates per query
    // instead of writing the [nq][nlist] key matrix that a separate selection kernel had to read back.  The bound is a
    // heuristic, so it is CHECKED: a query with fewer than nprobe + 16 candidates raises a counter and the caller repeats
    // the search with the dense path (counter slot 8; never observed on the benchmark shapes).
    bool coarse_tc_disabled = false;
    DevBuf<float> s_cbound;
    DevBuf<int64_t> s_coff;
    DevBuf<int32_t> s_clen;
    bool
    use_coarse_tc(int64_t m, int nprobe) const {
        const char* e = getenv("KB2_COARSE");
        if (coarse_tc_disabled || (e && !strcmp(e, "dense"))) return false;
        if (dim % fltc::BK != 0 || nlist < 1024 || nprobe + 16 > 256 || nprobe + 16 > nlist / 8) return false;
        // opt-in (KB2_COARSE=tc): measured at C3 (10000 x 4096 centroids) the list-major kernel + its sample bound cost
        // 0.75 ms against 0.43 ms of the dense GEMM + select it would replace (profiles/r2_summary.md)
        return e && !strcmp(e, "tc") && m >= 296;
    }
    void
    coarse_probes_tc(const float* Q, int64_t m, int nprobe, int64_t* out_ids, float* out_dis) {
        cudaStream_t st = stream;
        const int need = nprobe + 16;
        // ---- sample bound
        const int64_t ns = std::max<int64_t>(512, (nlist / 8 + 127) / 128 * 128);
        const int k_sample = (int)std::min<int64_t>(ns, 2 * ((int64_t)need * ns + nlist - 1) / nlist + 8);
        DensePlan pl = dense_candidates(*this, Q, m, centroids.p, cnorms.p, ns, dim, metric, std::max(k_sample, 17) - 16 + 16, nullptr, nullptr);
        KB2_REQUIRE(k_sample <= pl.Ksel, KB2_INTERNAL_ERROR, "coarse sample selection too small");
        s_cbound.ensure((size_t)m);
        fltc::extract_bound_kernel<<<grid1d(m, 256), 256, 0, st>>>(s_partial.p, pl.stride(), k_sample, m, s_cbound.p);
        // ---- items: chunks of consecutive queries over the single pseudo-list [0, nlist)
        const int item_cap = (m / 128 < 2 * kNumSMs) ? 32 : 128;
        const int64_t n_it = (m + item_cap - 1) / item_cap;
        const int64_t npairs_pad = m + fltc::NQ_ITEM;
        s_items.ensure((size_t)3 * n_it);
        s_plan_out.ensure(8);
        s_pair_q.ensure((size_t)m);
        s_qnorm.ensure((size_t)m);
        s_cand.ensure((size_t)m * kTcCandCap);
        s_cand_cnt.ensure((size_t)2 * m + 4);
        s_qhi.ensure((size_t)npairs_pad * dim);
        s_qlo.ensure((size_t)npairs_pad * dim);
        if (s_coff.n < 1) {
            s_coff.ensure(1);
            s_clen.ensure(1);
        }
        const int64_t h_off = 0;
        const int32_t h_len = (int32_t)nlist;
        KB2_CUDA_CHECK(cudaMemcpyAsync(s_coff.p, &h_off, 8, cudaMemcpyHostToDevice, st));
        KB2_CUDA_CHECK(cudaMemcpyAsync(s_clen.p, &h_len, 4, cudaMemcpyHostToDevice, st));
        KB2_CUDA_CHECK(cudaMemsetAsync(s_cand_cnt.p, 0, ((size_t)2 * m + 4) * 4, st));
        int32_t* item_list = s_items.p;
        int32_t* item_q0 = s_items.p + n_it;
        int32_t* item_nq = s_items.p + 2 * n_it;
        fltc::uniform_items_kernel<<<grid1d(std::max<int64_t>(m, n_it), 256), 256, 0, st>>>(m, item_cap, item_list, item_q0, item_nq,
                                                                                         s_plan_out.p, s_pair_q.p);
        fltc::gather_split_queries_kernel<<<grid1d(npairs_pad * 32, 256), 256, 0, st>>>(Q, s_pair_q.p, m, npairs_pad, dim, s_qhi.p, s_qlo.p);
        row_norms_kernel<<<grid1d(m * 32, 256), 256, 0, st>>>(Q, m, dim, s_qnorm.p);
        CUtensorMap tx, thi, tlo;
        KB2_REQUIRE(tc::make_tmap(&tx, centroids.p, nlist, dim) && tc::make_tmap(&thi, s_qhi.p, npairs_pad, dim, item_cap) &&
                        tc::make_tmap(&tlo, s_qlo.p, npairs_pad, dim, item_cap),
                    KB2_INTERNAL_ERROR, "coarse tensor-core stage: tensor map encoding failed");
        fltc::Params fpar{};
        fpar.metric = metric;
        fpar.d = dim;
        fpar.n_items = s_plan_out.p;
        if (tc_dynamic_sched()) {
            fpar.ticket = s_plan_out.p + 4;
            KB2_CUDA_CHECK(cudaMemsetAsync(fpar.ticket, 0, 4, st));
        }
        fpar.item_list = item_list;
        fpar.item_q0 = item_q0;
        fpar.item_nq = item_nq;
        fpar.pair_q = s_pair_q.p;
        fpar.qnorm2 = s_qnorm.p;
        fpar.bound = s_cbound.p;
        fpar.list_off = s_coff.p;
        fpar.list_len = s_clen.p;
        fpar.xnorm2 = cnorms.p;
        fpar.bitset = nullptr;
        fpar.rows = nullptr;
        const int grid = (int)std::min<int64_t>(kNumSMs, n_it);
        const uint32_t log_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(m * 1024 / grid, 32768), 1 << 20);
        s_log.ensure((size_t)kNumSMs * log_cap);
        s_logcnt.ensure(kNumSMs + 8);
        KB2_CUDA_CHECK(cudaMemsetAsync(s_logcnt.p, 0, (kNumSMs + 8) * 4, st));
        fpar.log = s_log.p;
        fpar.log_cnt = s_logcnt.p;
        fpar.log_cap = log_cap;
        fpar.counters = nullptr;
#define KB2_FL_LAUNCH(MM, BR) \
    fltc::ivfflat_tc_kernel<MM, BR><<<grid, fltc::THREADS, fltc::FlCfg<BR>::SMEM_BYTES, st>>>(tx, thi, tlo, fpar);
        if (metric == KB2_METRIC_L2) {
            if (item_cap == 32) { KB2_FL_LAUNCH(KB2_METRIC_L2, 32) } else { KB2_FL_LAUNCH(KB2_METRIC_L2, 128) }
        } else {
            if (item_cap == 32) { KB2_FL_LAUNCH(KB2_METRIC_IP, 32) } else { KB2_FL_LAUNCH(KB2_METRIC_IP, 128) }
        }
#undef KB2_FL_LAUNCH
        KB2_CUDA_CHECK(cudaGetLastError());
        uint32_t* qflag = s_cand_cnt.p + m;
        fltc::scatter_kernel<<<dim3(16, grid), 256, 0, st>>>(s_log.p, s_logcnt.p, log_cap, s_cand.p, s_cand_cnt.p, kTcCandCap, qflag);
        fltc::check_counts_kernel<<<grid1d(m, 256), 256, 0, st>>>(s_cand_cnt.p, qflag, m, (uint32_t)std::min<int64_t>(need, nlist),
                                                                s_logcnt.p + grid, d_counter.p + 8);
        last.launches += 8;
        FinalizeParams fp{};
        fp.partial = s_cand.p;
        fp.partial_stride = kTcCandCap;
        fp.n_partial = kTcCandCap;
        fp.counts = s_cand_cnt.p;
        fp.k_sel = (int)std::min<int64_t>(need, nlist);
        fp.k_out = nprobe;
        fp.rerank = 1;
        fp.raw = centroids.p;
        fp.raw_by_pos = 1;
        fp.queries = Q;
        fp.d = dim;
        fp.metric = metric;
        fp.out_ids = out_ids;
        fp.out_dist = out_dis;
        launch_finalize(*this, fp, m);
    }

    void
    coarse_probes(const float* dq, int64_t q_lo, int64_t q_hi, int nprobe) {
        const int64_t m = q_hi - q_lo;
        if (m <= 0) return;
        if (!distributed() && use_coarse_tc(m, nprobe)) {
            coarse_probes_tc(dq + q_lo * dim, m, nprobe, s_probe_ids.p + q_lo * nprobe, s_probe_dis.p + q_lo * nprobe);
            return;
        }
        DensePlan pl = dense_candidates(*this, dq + q_lo * dim, m, centroids.p, cnorms.p, nlist, dim, metric, nprobe + 16, nullptr,
                                        nullptr);
        FinalizeParams fp{};
        fp.partial = s_partial.p;
        fp.partial_stride = pl.stride();
        fp.n_partial = pl.used * pl.Ksel;
        fp.k_sel = (int)std::min<int64_t>(std::min(pl.Ksel, nprobe + 16), nlist);
        fp.k_out = nprobe;
        fp.rerank = 1;
        fp.raw = centroids.p;
        fp.raw_by_pos = 1;
        fp.queries = dq + q_lo * dim;
        fp.d = dim;
        fp.metric = metric;
        fp.out_ids = s_probe_ids.p + q_lo * nprobe;
        fp.out_dist = s_probe_dis.p + q_lo * nprobe;
        launch_finalize(*this, fp, m);
    }

    // ---------------------------------------------------------------- Search (ivf.cc:887-1168)
    void
    search(const float* q, int64_t nq, int k, const JsonObj& cfg, const uint8_t* bitset, int64_t nbits, int64_t* out_ids,
           float* out_dist) override {
        KB2_REQUIRE(trained, KB2_INDEX_NOT_TRAINED, "index not trained");
        KB2_REQUIRE(n_total > 0, KB2_EMPTY_INDEX, "index is empty");
        seal();
        int nprobe = (int)cfg.get_int("nprobe", 8);
        nprobe = (int)std::min<int64_t>(std::max(nprobe, 1), nlist);
        KB2_REQUIRE(nprobe <= kMaxK - 16, KB2_OUT_OF_RANGE_IN_JSON, "nprobe too large for the GPU path (max 1008)");
        const bool use_refine = is_pq && refine;
        const double refine_k = cfg.get_num("refine_k", 1.0);
        KB2_REQUIRE(refine_k >= 1.0, KB2_OUT_OF_RANGE_IN_JSON, "refine_k must be >= 1");
        const int k_base = use_refine ? (int)((double)k * refine_k) : k;  // K/IndexRefine.cpp:80-83
        KB2_REQUIRE(k > 0 && k_base <= kMaxK, KB2_INVALID_ARGS, "k (x refine_k) out of range (max 1024)");
        const bool dist = distributed();
        KB2_REQUIRE(!dist || (int64_t)shard_world * k <= kMaxSortEntries, KB2_INVALID_ARGS, "world * k too large for the merge");

        cudaStream_t st = stream;
        // (Uploading host queries in four pieces on a side stream with the coarse stage started per piece was measured at C3:
        //  e2e 2.10 / 2.03 ms without vs 2.12 / 2.09 ms with -- four quarter-size coarse passes cost what the copy hides.)
        const float* dq = to_device(q, (size_t)nq * dim, s_q);
        const uint8_t* dbits = bitset_to_device(bitset, nbits);
        const bool dev_out = is_device_ptr(out_ids);
        int64_t* d_ids = out_ids;
        float* d_dist = out_dist;
        if (!dev_out) {
            s_out_ids.ensure((size_t)nq * k);
            s_out_dist.ensure((size_t)nq * k);
            d_ids = s_out_ids.p;
            d_dist = s_out_dist.p;
        }

        KB2_CUDA_CHECK(cudaMemsetAsync(d_counter.p + 8, 0, 8, st));
        // ---- coarse quantizer.  With a communicator every rank ranks the centroids for its slice of the batch only and the
        //      probe lists are all-gathered (in place; slices padded to the same length).
        const int64_t per = dist ? (nq + shard_world - 1) / shard_world : nq;
        const int64_t nq_pad = dist ? per * shard_world : nq;
        s_probe_ids.ensure((size_t)nq_pad * nprobe);
        s_probe_dis.ensure((size_t)nq_pad * nprobe);
        if (dist) {
            const int64_t q_lo = std::min<int64_t>(nq, per * shard_rank), q_hi = std::min<int64_t>(nq, q_lo + per);
            coarse_probes(dq, q_lo, q_hi, nprobe);
            if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev_c0, st));
            comm->all_gather2(s_probe_ids.p + per * shard_rank * nprobe, s_probe_ids.p, (size_t)per * nprobe * 8,
                              s_probe_dis.p + per * shard_rank * nprobe, s_probe_dis.p, (size_t)per * nprobe * 4, st);
            if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev_c1, st));
        } else {
            coarse_probes(dq, 0, nq, nprobe);
        }

        // ---- visiting order of the queries: sorted by nearest list, so that CTAs resident at the same
        //      time probe the same lists (L2 reuse of codes; results are order-independent)
        const int32_t* qperm = nullptr;
        const bool tc_engine = use_tc_engine(nq, nprobe, next_pow2(std::max(32, k_base)));
        if (nq >= 2 * kNumSMs && !tc_engine) {
            s_qkey.ensure(nq); s_qkey2.ensure(nq); s_qidx.ensure(nq); s_qperm.ensure(nq);
            first_probe_kernel<<<grid1d(nq, 256), 256, 0, st>>>(s_probe_ids.p, nprobe, nq, s_qkey.p, s_qidx.p);
            size_t tmp_bytes = 0;
            int end_bit = 1;
            while ((1ll << end_bit) < nlist) end_bit++;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s_qkey.p, s_qkey2.p, s_qidx.p, s_qperm.p, (int)nq, 0, end_bit, st);
            s_sort_tmp.ensure(tmp_bytes);
            cub::DeviceRadixSort::SortPairs(s_sort_tmp.p, tmp_bytes, s_qkey.p, s_qkey2.p, s_qidx.p, s_qperm.p, (int)nq, 0,
                                            end_bit, st);
            qperm = s_qperm.p;
            last.launches += 3;
        }

        // ---- list scan
        int nsplit = 1;
        if (nq < 2 * kNumSMs) nsplit = (int)std::min<int64_t>(nprobe, (2 * kNumSMs + nq - 1) / nq);
        const int Ksel = next_pow2(std::max(32, k_base));
        while ((int64_t)nsplit * Ksel > kMaxSortEntries) nsplit--;
        const int np_max = (nprobe + nsplit - 1) / nsplit;
        s_partial2.ensure((size_t)nq * nsplit * Ksel);
        KB2_CUDA_CHECK(cudaMemsetAsync(d_counter.p, 0, 64, st));
        IvfScanParams sp{};
        sp.queries = dq;
        sp.nq = (int)nq;
        sp.d = dim;
        sp.metric = metric;
        sp.probe_ids = s_probe_ids.p;
        sp.probe_dis = s_probe_dis.p;
        sp.nprobe = nprobe;
        sp.list_off = list_off.p;
        sp.list_len = list_len.p;
        sp.nsplit = nsplit;
        sp.K = Ksel;
        sp.kout = Ksel;
        sp.partial = s_partial2.p;
        sp.bitset = dbits;
        sp.rows = rows.p;
        sp.vecs = vecs.p;
        sp.pq_centroids = pqc.p;
        sp.M = M;
        sp.dsub = dsub;
        sp.codes = (const uint4*)codes.p;
        sp.npad = npad;
        sp.t1 = t1.p;
        sp.counters = d_counter.p;
        sp.qperm = qperm;
        const unsigned grid = (unsigned)(nq * nsplit);
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev0, st));
        const uint64_t* fin_partial = s_partial2.p;
        int64_t fin_stride = (int64_t)nsplit * Ksel;
        int fin_n = nsplit * Ksel;
        const uint32_t *fin_counts = nullptr, *fin_flags = nullptr;
        bool flat_tc = false;
        if (tc_engine) {
            search_tc(sp, nq, nprobe, Ksel, k_base, dbits != nullptr);
            fin_partial = s_cand.p;
            fin_stride = kTcCandCap;
            fin_n = kTcCandCap;
            fin_counts = s_cand_cnt.p;
            fin_flags = s_cand_cnt.p + nq;
        } else if (use_flat_tc_engine(nq, nprobe, k_base) && search_flat_tc(sp, nq, nprobe, Ksel, k_base, dbits != nullptr)) {
            flat_tc = true;
            fin_partial = s_cand.p;
            fin_stride = kTcCandCap;
            fin_n = kTcCandCap;
            fin_counts = s_cand_cnt.p;
        } else {
            launch_scan(sp, grid, Ksel, np_max, dbits != nullptr);
        }
        if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev1, st));

        // ---- finalize: merge CTA lists, optional exact refine, labels.  With a communicator the local top-k goes to a
        //      staging buffer, ONE fused all-gather ships ids + distances of every shard, and the merge kernel writes the result.
        if (dist) ensure_gather_buffers(nq, k);
        {
            FinalizeParams fp{};
            fp.partial = fin_partial;
            fp.partial_stride = fin_stride;
            fp.n_partial = fin_n;
            fp.counts = fin_counts;
            fp.count_flags = fin_flags;
            fp.k_sel = flat_tc ? k_base + 16 : k_base;   // tensor-core IVF_FLAT: 3xTF32 keys, exact re-rank of the k+16 best
            fp.k_out = k;
            fp.rows = rows.p;
            fp.labels = custom_labels ? labels.p : nullptr;
            fp.rerank = (use_refine || flat_tc) ? 1 : 0;
            fp.raw = vecs.p;
            fp.raw16 = (is_pq && refine_kind) ? vecs16.p : nullptr;
            fp.raw16_kind = refine_kind;
            fp.raw_by_pos = 1;
            fp.queries = dq;
            fp.d = dim;
            fp.metric = metric;
            fp.out_ids = dist ? s_loc_ids.p : d_ids;
            fp.out_dist = dist ? s_loc_dist.p : d_dist;
            launch_finalize(*this, fp, nq);
        }
        if (dist) {
            if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev_c2, st));
            comm->all_gather2(s_loc_ids.p, s_g_ids.p, (size_t)nq * k * 8, s_loc_dist.p, s_g_dist.p, (size_t)nq * k * 4, st);
            launch_merge_topk(metric, shard_world, nq, k, s_g_ids.p, s_g_dist.p, d_ids, d_dist, st);
            if (timing) KB2_CUDA_CHECK(cudaEventRecord(ev_c3, st));
            last.launches += 3;
        }
        unsigned long long* hc = (unsigned long long*)h_counter.p;
        KB2_CUDA_CHECK(cudaMemcpyAsync(hc, d_counter.p, 72, cudaMemcpyDeviceToHost, st));
        results_out(nq, k, out_ids, out_dist, d_ids, d_dist);
        KB2_REQUIRE(hc[1] == 0 && hc[5] == 0, KB2_INTERNAL_ERROR, "ivfpq_scan_kernel: unexpected shared-memory window base");
        if (hc[8] != 0 && !coarse_tc_disabled) {
            // the sampled admission bound of the tensor-core coarse stage left some query short of candidates: repeat the
            // whole search with the dense coarse path (results of this pass are discarded)
            coarse_tc_disabled = true;
            try {
                search(q, nq, k, cfg, bitset, nbits, out_ids, out_dist);
            } catch (...) {
                coarse_tc_disabled = false;
                throw;
            }
            coarse_tc_disabled = false;
            last.flagged += (int64_t)hc[8];
            return;
        }
        last.survivors = (int64_t)hc[2];
        last.flagged = (int64_t)hc[7];
        const unsigned long long scanned = hc[0];
        last.codes = (int64_t)scanned;
        last.code_bytes = (int64_t)scanned * (is_pq ? (int64_t)M : (int64_t)dim * 4);
        last.pairs = nq * nprobe;
        last_engine = (tc_engine || flat_tc) ? 1 : 0;
        if (timing) {
            KB2_CUDA_CHECK(cudaEventElapsedTime(&last_stage_ms, ev0, ev1));
            last_kernel_ms = last_stage_ms;
            if (tc_engine || flat_tc) KB2_CUDA_CHECK(cudaEventElapsedTime(&last_kernel_ms, ev2, ev3));
            last_comm_ms = 0.f;
            if (dist) {
                float a = 0.f, b = 0.f;
                KB2_CUDA_CHECK(cudaEventElapsedTime(&a, ev_c0, ev_c1));
                KB2_CUDA_CHECK(cudaEventElapsedTime(&b, ev_c2, ev_c3));
                last_comm_ms = a + b;
            }
        }
    }

    void
    get_vectors(const int64_t* ids, int64_t n, float* out) override {
        KB2_REQUIRE(keeps_vecs(), KB2_NOT_IMPLEMENTED, "index holds no raw data");
        KB2_REQUIRE(!custom_labels && shard_world == 1, KB2_NOT_IMPLEMENTED, "GetVectorByIds with custom ids / shards");
        seal();
        std::vector<int64_t> h(n);
        if (is_device_ptr(ids)) {
            KB2_CUDA_CHECK(cudaMemcpy(h.data(), ids, n * 8, cudaMemcpyDeviceToHost));
        } else {
            memcpy(h.data(), ids, n * 8);
        }
        std::vector<int32_t> hpos(n_total);
        KB2_CUDA_CHECK(cudaMemcpy(hpos.data(), pos_of_row.p, n_total * 4, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; i++) {
            KB2_REQUIRE(h[i] >= 0 && h[i] < n_total, KB2_INVALID_ARGS, "id out of range");
            KB2_CUDA_CHECK(cudaMemcpyAsync(out + i * dim, vecs.p + (int64_t)hpos[h[i]] * dim, (size_t)dim * 4,
                                           cudaMemcpyDefault, stream));
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    // ---------------------------------------------------------------- import of an externally built index
    std::vector<int32_t> imp_assign;
    std::vector<int64_t> imp_labels;
    std::vector<uint8_t> imp_codes;
    void
    import_begin(int64_t nl, const float* cent, const float* pq_cent) {
        KB2_REQUIRE(n_total == 0, KB2_INVALID_ARGS, "import into a non-empty index");
        nlist = nl;
        centroids.alloc_exact((size_t)nlist * dim);
        KB2_CUDA_CHECK(cudaMemcpyAsync(centroids.p, cent, (size_t)nlist * dim * 4, cudaMemcpyDefault, stream));
        set_centroids_common();
        if (is_pq) {
            KB2_REQUIRE(pq_cent != nullptr, KB2_INVALID_ARGS, "IVF_PQ import needs pq centroids");
            KB2_REQUIRE(M > 0 && dim % M == 0 && nbits == 8, KB2_INVALID_ARGS, "IVF_PQ import: bad m / nbits");
            dsub = dim / M;
            pqc.alloc_exact((size_t)M * 256 * dsub);
            tc_ready = false;
            KB2_CUDA_CHECK(cudaMemcpyAsync(pqc.p, pq_cent, (size_t)M * 256 * dsub * 4, cudaMemcpyDefault, stream));
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
        trained = true;
        imp_assign.clear();
        imp_labels.clear();
        imp_codes.clear();
    }
    void
    import_list(int64_t l, int64_t sz, const int64_t* ids, const uint8_t* cds) {
        KB2_REQUIRE(l >= 0 && l < nlist, KB2_INVALID_ARGS, "list number out of range");
        const size_t cs = is_pq ? (size_t)M : (size_t)dim * 4;
        imp_assign.insert(imp_assign.end(), (size_t)sz, (int32_t)l);
        imp_labels.insert(imp_labels.end(), ids, ids + sz);
        imp_codes.insert(imp_codes.end(), cds, cds + (size_t)sz * cs);
    }
    // raw (refine=true): [n_raw][dim] rows addressed by label, or — raw_in_import_order — by the order of the import_list calls
    void
    import_finish(const float* raw, int64_t n_raw, bool raw_in_import_order = false) {
        const int64_t n = (int64_t)imp_assign.size();
        std::vector<int32_t> orig_of_row;   // row -> position in the import stream (only when raw_in_import_order)
        // Rows were numbered in import (list) order.  The reference's ids are segment offsets 0..n-1 and a BitsetView /
        // GetVectorByIds address vectors by that id (bitsetview.h:131-175), so when the labels are a permutation of
        // 0..n-1 renumber the rows so that row == label: bitset tests, GetVectorByIds and add() then behave exactly as
        // on an index built here.  Within a list the scan order becomes ascending id (the reference's insertion order).
        {
            bool perm = n > 0;
            std::vector<uint8_t> seen((size_t)n, 0);
            for (int64_t i = 0; i < n && perm; i++) {
                const int64_t l = imp_labels[i];
                if (l < 0 || l >= n || seen[l]) perm = false; else seen[l] = 1;
            }
            if (perm) {
                const size_t cs = is_pq ? (size_t)M : (size_t)dim * 4;
                std::vector<int32_t> a2((size_t)n);
                std::vector<uint8_t> c2((size_t)n * cs);
                if (raw_in_import_order) orig_of_row.resize((size_t)n);
                for (int64_t i = 0; i < n; i++) {
                    const int64_t l = imp_labels[i];
                    if (raw_in_import_order) orig_of_row[l] = (int32_t)i;
                    a2[l] = imp_assign[i];
                    memcpy(c2.data() + (size_t)l * cs, imp_codes.data() + (size_t)i * cs, cs);
                }
                imp_assign.swap(a2);
                imp_codes.swap(c2);
                for (int64_t i = 0; i < n; i++) imp_labels[i] = i;
            }
            custom_labels = !perm;
        }
        f_assign_used = 0;
        dev_append(f_assign, f_assign_used, imp_assign.data(), (size_t)n, stream);
        f_labels_used = 0;
        if (custom_labels) dev_append(f_labels, f_labels_used, imp_labels.data(), (size_t)n, stream);
        if (is_pq) {
            f_codes_used = 0;
            dev_append(f_codes, f_codes_used, imp_codes.data(), imp_codes.size(), stream);
            if (refine) {
                KB2_REQUIRE(raw != nullptr, KB2_INVALID_ARGS, "refine=true import needs the raw vectors");
                // gather the raw vector of every row: raw[label[row]], or raw[import position of row]
                std::vector<int32_t> lab32(n);
                for (int64_t i = 0; i < n; i++) {
                    const int64_t src = raw_in_import_order ? (orig_of_row.empty() ? i : (int64_t)orig_of_row[i]) : imp_labels[i];
                    KB2_REQUIRE(src >= 0 && src < n_raw, KB2_INVALID_ARGS, "label outside raw data");
                    lab32[i] = (int32_t)src;
                }
                DevBuf<float> rbuf;
                const float* draw = to_device(raw, (size_t)n_raw * dim, rbuf, false);
                DevBuf<int32_t> dl;
                dl.ensure(n);
                KB2_CUDA_CHECK(cudaMemcpyAsync(dl.p, lab32.data(), n * 4, cudaMemcpyHostToDevice, stream));
                f_vecs.alloc_exact((size_t)n * dim);
                gather_rows_kernel<<<grid1d(n * 32, 256), 256, 0, stream>>>(draw, dl.p, n, dim, dim, f_vecs.p);
                f_vecs_used = (size_t)n * dim;
                KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
            }
        } else {
            f_vecs_used = 0;
            dev_append(f_vecs, f_vecs_used, (const float*)imp_codes.data(), (size_t)n * dim, stream);
        }
        KB2_CUDA_CHECK(cudaStreamSynchronize(stream));
        n_total = n;
        imp_assign.clear(); imp_assign.shrink_to_fit();
        imp_labels.clear(); imp_labels.shrink_to_fit();
        imp_codes.clear(); imp_codes.shrink_to_fit();
        sealed = false;
    }
    // export one list in scan order (host buffers)
    void
    export_list(int64_t l, int64_t* ids, uint8_t* cds) {
        seal();
        KB2_REQUIRE(l >= 0 && l < nlist, KB2_INVALID_ARGS, "list number out of range");
        const int64_t off = h_list_off[l], len = h_list_len[l];
        if (len == 0) return;
        std::vector<int32_t> hrows(len);
        KB2_CUDA_CHECK(cudaMemcpy(hrows.data(), rows.p + off, len * 4, cudaMemcpyDeviceToHost));
        std::vector<int64_t> hl;
        if (custom_labels) {
            hl.resize(n_total);
            KB2_CUDA_CHECK(cudaMemcpy(hl.data(), labels.p, n_total * 8, cudaMemcpyDeviceToHost));
        }
        for (int64_t i = 0; i < len; i++) ids[i] = custom_labels ? hl[hrows[i]] : hrows[i];
        if (!is_pq) {
            KB2_CUDA_CHECK(cudaMemcpy(cds, vecs.p + off * dim, (size_t)len * dim * 4, cudaMemcpyDeviceToHost));
        } else if (G > 0) {
            std::vector<uint8_t> tmp((size_t)len * 16);
            for (int g = 0; g < G; g++) {
                KB2_CUDA_CHECK(cudaMemcpy(tmp.data(), codes.p + ((size_t)g * npad + off) * 16, (size_t)len * 16,
                                          cudaMemcpyDeviceToHost));
                for (int64_t i = 0; i < len; i++)
                    for (int s = 0; s < 16; s++)
                        cds[i * M + g * 16 + ((s + (off + i)) & 15)] = tmp[i * 16 + s];
            }
        } else {
            KB2_CUDA_CHECK(cudaMemcpy(cds, codes.p + (size_t)off * M, (size_t)len * M, cudaMemcpyDeviceToHost));
        }
    }
};

}  // namespace kb2
