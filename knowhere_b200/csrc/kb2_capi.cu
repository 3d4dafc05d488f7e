// kb2_capi.cu — the extern "C" boundary declared in include/knowhere_b200.h.
// Everything behind it is CUDA; there is no CPU fallback: without a usable sm_100 device every
// entry point fails with KB2_CUDA_RUNTIME_ERROR.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstring>
#include <memory>

#include "kb2_fourcc.h"
#include "kb2_hnsw.cuh"
#include "kb2_index.cuh"
#include "kb2_range.cuh"

using namespace kb2;

namespace {
thread_local std::string g_last_error;

template <typename F>
int
guarded(F&& f) {
    try {
        f();
        return KB2_SUCCESS;
    } catch (const Error& e) {
        g_last_error = e.what();
        cudaGetLastError();
        return e.status;
    } catch (const std::bad_alloc&) {
        g_last_error = "host allocation failed";
        return KB2_MALLOC_ERROR;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return KB2_INTERNAL_ERROR;
    } catch (...) {
        g_last_error = "unknown error";
        return KB2_INTERNAL_ERROR;
    }
}

int
usable_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

void
require_device(int device) {
    // validated devices are cached: cudaGetDeviceProperties costs milliseconds and this sits on per-call paths
    static std::atomic<uint64_t> ok_mask{0};
    if (device >= 0 && device < 64 && (ok_mask.load(std::memory_order_relaxed) >> device) & 1) {
        KB2_CUDA_CHECK(cudaSetDevice(device));
        return;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        throw Error(KB2_CUDA_RUNTIME_ERROR, "no CUDA device available (this library has no CPU fallback)");
    }
    KB2_REQUIRE(device >= 0 && device < n, KB2_INVALID_ARGS, "bad device ordinal");
    cudaDeviceProp p;
    KB2_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    KB2_REQUIRE(p.major == 10, KB2_CUDA_RUNTIME_ERROR,
                "device is not sm_100 (this library ships sm_100a SASS only)");
    KB2_CUDA_CHECK(cudaSetDevice(device));
    if (device < 64) ok_mask.fetch_or(1ull << device, std::memory_order_relaxed);
}

struct Handle {
    std::unique_ptr<IndexBase> ix;
};
inline IndexBase*
ix_of(kb2_index_t h) {
    KB2_REQUIRE(h != nullptr, KB2_INVALID_ARGS, "null index handle");
    return reinterpret_cast<Handle*>(h)->ix.get();
}
template <typename T>
inline T*
ix_as(kb2_index_t h, const char* what) {
    T* p = dynamic_cast<T*>(ix_of(h));
    KB2_REQUIRE(p != nullptr, KB2_INVALID_ARGS, std::string("handle is not an ") + what + " index");
    return p;
}

int
parse_metric(int metric, const JsonObj& cfg) {
    if (cfg.has("metric_type")) {
        const std::string m = cfg.get_str("metric_type", "L2");
        if (m == "L2") return KB2_METRIC_L2;
        if (m == "IP") return KB2_METRIC_IP;
        if (m == "COSINE") return KB2_METRIC_COSINE;
        throw Error(KB2_INVALID_METRIC_TYPE, "unsupported metric_type " + m);
    }
    return metric;
}
}  // namespace

extern "C" {

const char*
kb2_version(void) {
    return "knowhere_b200 0.1 (sm_100a)";
}
const char*
kb2_last_error(void) {
    return g_last_error.c_str();
}
int
kb2_device_count(void) {
    return usable_devices();
}

int
kb2_index_create(const char* index_type, int metric, int dim, const char* json_cfg, int device, kb2_index_t* out) {
    return guarded([&] {
        KB2_REQUIRE(out != nullptr && index_type != nullptr, KB2_INVALID_ARGS, "null argument");
        *out = nullptr;
        JsonObj cfg = JsonObj::parse(json_cfg);
        KB2_REQUIRE(cfg.ok, KB2_INVALID_PARAM_IN_JSON, "malformed json");
        metric = parse_metric(metric, cfg);
        KB2_REQUIRE(metric == KB2_METRIC_L2 || metric == KB2_METRIC_IP || metric == KB2_METRIC_COSINE,
                    KB2_INVALID_METRIC_TYPE, "metric must be L2, IP or COSINE");
        // COSINE = inner product of L2-normalised vectors: data is normalised when it enters the index and
        // queries when they are searched (what the reference does for IVF_PQ, ivf.cc:557-565,1067-1071; for FLAT
        // the reference keeps inverse norms instead, flat.cc:57-62 — same similarities)
        const bool cosine = (metric == KB2_METRIC_COSINE);
        if (cosine) metric = KB2_METRIC_IP;
        if (dim <= 0) dim = (int)cfg.get_int("dim", 0);
        KB2_REQUIRE(dim > 0, KB2_INVALID_ARGS, "dim must be positive");
        require_device(device);
        const std::string t = index_type;
        std::unique_ptr<IndexBase> ix;
        if (t == "FLAT") {
            ix.reset(new FlatIndex());
        } else if (t == "IVF_FLAT" || t == "IVF_PQ") {
            auto* iv = new IvfIndex();
            ix.reset(iv);
            iv->is_pq = (t == "IVF_PQ");
            iv->nlist = cfg.get_int("nlist", 128);
            KB2_REQUIRE(iv->nlist >= 1 && iv->nlist <= 65536 * 16, KB2_OUT_OF_RANGE_IN_JSON, "nlist out of range");
            if (iv->is_pq) {
                KB2_REQUIRE(cfg.has("m"), KB2_INVALID_PARAM_IN_JSON, "IVF_PQ requires m");
                iv->M = (int)cfg.get_int("m", 0);
                iv->nbits = (int)cfg.get_int("nbits", 8);
                KB2_REQUIRE(iv->M >= 1 && dim % iv->M == 0, KB2_INVALID_ARGS, "dim must be a multiple of m");
                KB2_REQUIRE(iv->nbits >= 1 && iv->nbits <= 24, KB2_OUT_OF_RANGE_IN_JSON, "nbits out of range");
                iv->refine = cfg.get_bool("refine", false);
                if (iv->refine) {
                    std::string rt = cfg.get_str("refine_type", "flat");
                    for (auto& ch : rt) ch = (char)tolower((unsigned char)ch);
                    if (rt == "flat" || rt == "fp32" || rt == "float32" || rt == "data_view") iv->refine_kind = 0;
                    else if (rt == "fp16" || rt == "float16") iv->refine_kind = 1;
                    else if (rt == "bf16" || rt == "bfloat16") iv->refine_kind = 2;
                    else throw Error(KB2_NOT_IMPLEMENTED, "refine_type " + rt + " is not implemented (flat / fp16 / bf16 are)");
                }
            }
        } else if (t == "HNSW") {
            auto* hn = new HnswIndex();
            ix.reset(hn);
            hn->M = (int)cfg.get_int("M", 30);
            hn->efConstruction = (int)cfg.get_int("efConstruction", 360);
            KB2_REQUIRE(hn->M >= 2 && hn->M <= 2048, KB2_OUT_OF_RANGE_IN_JSON, "M out of range");
        } else {
            throw Error(KB2_INVALID_ARGS, "unknown index type " + t);
        }
        ix->type = t;
        ix->metric = metric;
        ix->cosine = cosine;
        ix->dim = dim;
        ix->device = device;
        ix->init_common();
        auto* h = new Handle();
        h->ix = std::move(ix);
        *out = reinterpret_cast<kb2_index_t>(h);
    });
}

void
kb2_index_destroy(kb2_index_t h) {
    if (!h) return;
    Handle* hh = reinterpret_cast<Handle*>(h);
    if (hh->ix) cudaSetDevice(hh->ix->device);
    delete hh;
}

int
kb2_index_set_stream(kb2_index_t h, void* cuda_stream) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        ix->set_stream((cudaStream_t)cuda_stream);
    });
}

int
kb2_index_set_shard(kb2_index_t h, int rank, int world) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        KB2_REQUIRE(world >= 1 && rank >= 0 && rank < world, KB2_INVALID_ARGS, "bad shard rank/world");
        KB2_REQUIRE(ix->count() == 0, KB2_INVALID_ARGS, "set_shard must precede add/import");
        ix->shard_rank = rank;
        ix->shard_world = world;
    });
}

namespace {
// element types of the reference's data-type registrations: fp16 / bf16 / int8 indexes are "mock" wrappers that convert the
// whole dataset and every query batch to fp32 (include/knowhere/index/index_factory.h:95-103,
// src/index/index_node_data_mock_wrapper.cc:24-60); here the widening runs on the device
__global__ void
widen_kernel(const void* __restrict__ src, int dtype, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v;
    if (dtype == KB2_DTYPE_F16) v = __half2float(((const __half*)src)[i]);
    else if (dtype == KB2_DTYPE_BF16) v = __bfloat162float(((const __nv_bfloat16*)src)[i]);
    else v = (float)((const int8_t*)src)[i];
    out[i] = v;
}
// device fp32 view of `count` elements of type dtype (host or device source); valid until the next typed call on the handle
const float*
widen_to_f32(IndexBase* ix, const void* x, int dtype, int64_t count) {
    if (dtype == KB2_DTYPE_F32) return (const float*)x;
    KB2_REQUIRE(dtype == KB2_DTYPE_F16 || dtype == KB2_DTYPE_BF16 || dtype == KB2_DTYPE_INT8, KB2_INVALID_ARGS, "unknown data type");
    if (count <= 0 || !x) return nullptr;
    const size_t esz = dtype == KB2_DTYPE_INT8 ? 1 : 2;
    const void* dsrc = x;
    if (!is_device_ptr(x)) {
        ix->s_typed_raw.ensure((size_t)count * esz);
        KB2_CUDA_CHECK(cudaMemcpyAsync(ix->s_typed_raw.p, x, (size_t)count * esz, cudaMemcpyHostToDevice, ix->stream));
        ix->last.h2d += (int64_t)count * (int64_t)esz;
        dsrc = ix->s_typed_raw.p;
    }
    ix->s_typed_f32.ensure((size_t)count);
    widen_kernel<<<grid1d(count, 256), 256, 0, ix->stream>>>(dsrc, dtype, count, ix->s_typed_f32.p);
    KB2_CUDA_CHECK(cudaGetLastError());
    return ix->s_typed_f32.p;
}
}  // namespace

int
kb2_index_train_typed(kb2_index_t h, const void* x, int dtype, int64_t n) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        KB2_REQUIRE(x != nullptr || n == 0, KB2_INVALID_ARGS, "null training data");
        ix->wait_caller_work();
        const float* xf = widen_to_f32(ix, x, dtype, n * ix->dim);
        ix->train(ix->cosine ? ix->normalized(xf, n) : xf, n);
    });
}
int
kb2_index_train(kb2_index_t h, const float* x, int64_t n) {
    return kb2_index_train_typed(h, x, KB2_DTYPE_F32, n);
}

int
kb2_index_add_typed(kb2_index_t h, const void* x, int dtype, int64_t n, const int64_t* ids) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        KB2_REQUIRE(x != nullptr || n == 0, KB2_INVALID_ARGS, "null data");
        ix->wait_caller_work();
        const float* xf = widen_to_f32(ix, x, dtype, n * ix->dim);
        ix->add(ix->cosine ? ix->normalized(xf, n) : xf, n, ids);
    });
}
int
kb2_index_add(kb2_index_t h, const float* x, int64_t n, const int64_t* ids) {
    return kb2_index_add_typed(h, x, KB2_DTYPE_F32, n, ids);
}

int
kb2_index_search_typed(kb2_index_t h, const void* queries, int dtype, int64_t nq, int k, const char* json, const uint8_t* bitset,
                       int64_t bitset_nbits, int64_t* out_ids, float* out_dist) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        KB2_REQUIRE(nq >= 0 && k > 0, KB2_INVALID_ARGS, "bad nq / k");
        if (nq == 0) return;
        KB2_REQUIRE(queries && out_ids && out_dist, KB2_INVALID_ARGS, "null buffer");
        KB2_REQUIRE(is_device_ptr(out_ids) == is_device_ptr(out_dist), KB2_INVALID_ARGS,
                    "out_ids and out_dist must both be host or both be device buffers");
        JsonObj cfg = JsonObj::parse(json);
        KB2_REQUIRE(cfg.ok, KB2_INVALID_PARAM_IN_JSON, "malformed json");
        ix->last = Counters{};
        ix->wait_caller_work();
        const float* qf = widen_to_f32(ix, queries, dtype, nq * ix->dim);
        ix->search(ix->cosine ? ix->normalized(qf, nq) : qf, nq, k, cfg, bitset, bitset_nbits, out_ids, out_dist);
    });
}
int
kb2_index_search(kb2_index_t h, const float* queries, int64_t nq, int k, const char* json, const uint8_t* bitset,
                 int64_t bitset_nbits, int64_t* out_ids, float* out_dist) {
    return kb2_index_search_typed(h, queries, KB2_DTYPE_F32, nq, k, json, bitset, bitset_nbits, out_ids, out_dist);
}

int
kb2_index_range_search(kb2_index_t h, const float* queries, int64_t nq, float radius, float range_filter,
                       int has_range_filter, const char* json, const uint8_t* bitset, int64_t bitset_nbits,
                       int64_t** out_lims, int64_t** out_ids, float** out_dist) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        KB2_REQUIRE(out_lims && out_ids && out_dist, KB2_INVALID_ARGS, "null output");
        JsonObj cfg = JsonObj::parse(json);
        KB2_REQUIRE(cfg.ok, KB2_INVALID_PARAM_IN_JSON, "malformed json");
        ix->last = Counters{};
        ix->wait_caller_work();
        range_search_index(*ix, ix->cosine ? ix->normalized(queries, nq) : queries, nq, radius, range_filter,
                           has_range_filter != 0, cfg, bitset, bitset_nbits, out_lims, out_ids, out_dist);
    });
}

void
kb2_free(void* p) {
    free(p);
}

int64_t
kb2_index_count(kb2_index_t h) {
    return h ? reinterpret_cast<Handle*>(h)->ix->count() : 0;
}
int
kb2_index_dim(kb2_index_t h) {
    return h ? reinterpret_cast<Handle*>(h)->ix->dim : 0;
}
int64_t
kb2_index_size_bytes(kb2_index_t h) {
    return h ? reinterpret_cast<Handle*>(h)->ix->size_bytes() : 0;
}
int
kb2_index_is_trained(kb2_index_t h) {
    return h ? (int)reinterpret_cast<Handle*>(h)->ix->is_trained() : 0;
}
int
kb2_index_has_raw_data(kb2_index_t h) {
    // COSINE stores the normalised vectors, not the caller's raw data
    return h ? (int)(reinterpret_cast<Handle*>(h)->ix->has_raw() && !reinterpret_cast<Handle*>(h)->ix->cosine) : 0;
}
int
kb2_index_get_vector_by_ids(kb2_index_t h, const int64_t* ids, int64_t n, float* out) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        KB2_REQUIRE(!ix->cosine, KB2_NOT_IMPLEMENTED, "GetVectorByIds: a COSINE index keeps normalised vectors only");
        ix->get_vectors(ids, n, out);
    });
}

// ---------------------------------------------------------------- IVF import / export
int
kb2_ivf_import_begin(kb2_index_t h, int64_t nlist, const float* centroids, const float* pq_centroids) {
    return guarded([&] {
        auto* iv = ix_as<IvfIndex>(h, "IVF");
        std::lock_guard<std::mutex> lk(iv->mu);
        KB2_CUDA_CHECK(cudaSetDevice(iv->device));
        iv->import_begin(nlist, centroids, pq_centroids);
    });
}
int
kb2_ivf_import_list(kb2_index_t h, int64_t list_no, int64_t list_size, const int64_t* ids, const uint8_t* codes) {
    return guarded([&] {
        auto* iv = ix_as<IvfIndex>(h, "IVF");
        std::lock_guard<std::mutex> lk(iv->mu);
        if (list_size > 0) iv->import_list(list_no, list_size, ids, codes);
    });
}
int
kb2_ivf_import_finish(kb2_index_t h, const float* raw, int64_t n_raw) {
    return guarded([&] {
        auto* iv = ix_as<IvfIndex>(h, "IVF");
        std::lock_guard<std::mutex> lk(iv->mu);
        KB2_CUDA_CHECK(cudaSetDevice(iv->device));
        iv->import_finish(raw, n_raw);
    });
}
int64_t
kb2_ivf_nlist(kb2_index_t h) {
    auto* iv = dynamic_cast<IvfIndex*>(reinterpret_cast<Handle*>(h)->ix.get());
    return iv ? iv->nlist : -1;
}
int64_t
kb2_ivf_list_size(kb2_index_t h, int64_t list_no) {
    int64_t r = -1;
    guarded([&] {
        auto* iv = ix_as<IvfIndex>(h, "IVF");
        std::lock_guard<std::mutex> lk(iv->mu);
        KB2_CUDA_CHECK(cudaSetDevice(iv->device));
        iv->seal();
        KB2_REQUIRE(list_no >= 0 && list_no < iv->nlist, KB2_INVALID_ARGS, "list number out of range");
        r = iv->h_list_len[list_no];
    });
    return r;
}
int
kb2_ivf_export_centroids(kb2_index_t h, float* centroids, float* pq_centroids) {
    return guarded([&] {
        auto* iv = ix_as<IvfIndex>(h, "IVF");
        std::lock_guard<std::mutex> lk(iv->mu);
        KB2_CUDA_CHECK(cudaSetDevice(iv->device));
        KB2_REQUIRE(iv->trained, KB2_INDEX_NOT_TRAINED, "index not trained");
        if (centroids)
            KB2_CUDA_CHECK(cudaMemcpy(centroids, iv->centroids.p, (size_t)iv->nlist * iv->dim * 4, cudaMemcpyDefault));
        if (pq_centroids && iv->is_pq)
            KB2_CUDA_CHECK(cudaMemcpy(pq_centroids, iv->pqc.p, (size_t)iv->M * 256 * iv->dsub * 4, cudaMemcpyDefault));
    });
}
int
kb2_ivf_export_list(kb2_index_t h, int64_t list_no, int64_t* ids, uint8_t* codes) {
    return guarded([&] {
        auto* iv = ix_as<IvfIndex>(h, "IVF");
        std::lock_guard<std::mutex> lk(iv->mu);
        KB2_CUDA_CHECK(cudaSetDevice(iv->device));
        iv->export_list(list_no, ids, codes);
    });
}

// ---------------------------------------------------------------- HNSW import / export
int
kb2_hnsw_import(kb2_index_t h, int64_t n, const float* vectors, const int32_t* levels, const int64_t* offsets,
                const int32_t* neighbors, const int32_t* cum_nneighbor, int n_cum, int32_t entry_point,
                int32_t max_level) {
    return guarded([&] {
        auto* hn = ix_as<HnswIndex>(h, "HNSW");
        std::lock_guard<std::mutex> lk(hn->mu);
        KB2_CUDA_CHECK(cudaSetDevice(hn->device));
        hn->import_graph(n, vectors, levels, offsets, neighbors, cum_nneighbor, n_cum, entry_point, max_level);
    });
}
int
kb2_hnsw_export_meta(kb2_index_t h, int64_t* out5) {
    return guarded([&] {
        auto* hn = ix_as<HnswIndex>(h, "HNSW");
        out5[0] = hn->n;
        out5[1] = hn->entry_point;
        out5[2] = hn->max_level;
        out5[3] = (int64_t)hn->h_neighbors.size();
        out5[4] = (int64_t)hn->h_cum.size();
    });
}
int
kb2_hnsw_export(kb2_index_t h, int32_t* levels, int64_t* offsets, int32_t* neighbors, int32_t* cum) {
    return guarded([&] {
        auto* hn = ix_as<HnswIndex>(h, "HNSW");
        memcpy(levels, hn->h_levels.data(), hn->h_levels.size() * 4);
        memcpy(offsets, hn->h_offsets.data(), hn->h_offsets.size() * 8);
        memcpy(neighbors, hn->h_neighbors.data(), hn->h_neighbors.size() * 4);
        memcpy(cum, hn->h_cum.data(), hn->h_cum.size() * 4);
    });
}
int
kb2_hnsw_last_stats(kb2_index_t h, int64_t* out2) {
    return guarded([&] {
        auto* hn = ix_as<HnswIndex>(h, "HNSW");
        out2[0] = hn->last_ndis;
        out2[1] = hn->last_nhops;
    });
}

// ---------------------------------------------------------------- serialisation ("KB2I" container)
int
kb2_index_serialize(kb2_index_t h, uint8_t** out, size_t* out_size) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        std::vector<uint8_t> blob;
        serialize_index(*ix, blob);
        *out = (uint8_t*)malloc(blob.size() ? blob.size() : 1);
        KB2_REQUIRE(*out != nullptr, KB2_MALLOC_ERROR, "malloc failed");
        memcpy(*out, blob.data(), blob.size());
        *out_size = blob.size();
    });
}
int
kb2_index_deserialize(const uint8_t* blob, size_t size, int device, kb2_index_t* out) {
    return guarded([&] {
        KB2_REQUIRE(blob && out, KB2_INVALID_ARGS, "null argument");
        require_device(device);
        std::unique_ptr<IndexBase> ix = deserialize_index(blob, size, device);
        auto* hh = new Handle();
        hh->ix = std::move(ix);
        *out = reinterpret_cast<kb2_index_t>(hh);
    });
}

// ---------------------------------------------------------------- faiss fourcc streams (the reference's BinarySet payload)
namespace {
std::unique_ptr<IndexBase>
index_from_faiss(const FaissIndexData& o, int device) {
    std::unique_ptr<IndexBase> ix;
    if (o.kind == "FLAT") {
        auto* fi = new FlatIndex();
        ix.reset(fi);
    } else if (o.kind == "IVF_FLAT" || o.kind == "IVF_PQ") {
        KB2_REQUIRE(!o.cosine, KB2_NOT_IMPLEMENTED, "faiss stream: cosine IVF indexes are not supported");
        auto* iv = new IvfIndex();
        ix.reset(iv);
        iv->is_pq = (o.kind == "IVF_PQ");
        iv->nlist = o.nlist;
        iv->M = o.M;
        iv->nbits = 8;
        iv->refine = o.has_refine;
    } else if (o.kind == "HNSW") {
        auto* hn = new HnswIndex();
        ix.reset(hn);
        hn->M = o.cum.size() >= 2 ? o.cum[1] / 2 : 16;
        hn->efConstruction = o.efConstruction;
    } else {
        throw Error(KB2_NOT_IMPLEMENTED, "faiss stream: unsupported index kind");
    }
    ix->type = o.kind;
    ix->metric = o.metric;
    ix->cosine = o.cosine;
    if (o.cosine) ix->metric = KB2_METRIC_IP;
    ix->dim = o.d;
    ix->device = device;
    ix->init_common();
    if (auto* fi = dynamic_cast<FlatIndex*>(ix.get())) {
        if (o.ntotal) fi->add(o.cosine ? fi->normalized(o.xb.data(), o.ntotal) : o.xb.data(), o.ntotal, nullptr);
    } else if (auto* iv = dynamic_cast<IvfIndex*>(ix.get())) {
        iv->import_begin(o.nlist, o.centroids.data(), iv->is_pq ? o.pq_centroids.data() : nullptr);
        for (int64_t l = 0; l < o.nlist; l++)
            if (!o.list_ids[l].empty()) iv->import_list(l, (int64_t)o.list_ids[l].size(), o.list_ids[l].data(), o.list_codes[l].data());
        iv->import_finish(o.has_refine ? o.refine_xb.data() : nullptr, o.has_refine ? o.ntotal : 0);
    } else if (auto* hn = dynamic_cast<HnswIndex*>(ix.get())) {
        std::vector<int64_t> off(o.offsets.begin(), o.offsets.end());
        std::vector<float> normed;
        const float* xb = o.xb.data();
        if (o.cosine) {   // the reference keeps raw rows + norms (IHN9); this core keeps unit rows
            normed = o.xb;
            for (int64_t i = 0; i < o.ntotal; i++) {
                double s2 = 0;
                for (int j = 0; j < o.d; j++) s2 += (double)normed[i * o.d + j] * normed[i * o.d + j];
                const float inv = s2 > 0 ? (float)(1.0 / std::sqrt(s2)) : 1.f;
                for (int j = 0; j < o.d; j++) normed[i * o.d + j] *= inv;
            }
            xb = normed.data();
        }
        hn->import_graph(o.ntotal, xb, o.levels.data(), off.data(), o.neighbors.data(), o.cum.data(), (int)o.cum.size(), o.entry_point,
                         o.max_level);
    }
    return ix;
}

void
index_to_faiss(IndexBase& ix, FaissIndexData& o) {
    o.kind = ix.type;
    o.d = ix.dim;
    o.metric = ix.metric;
    KB2_REQUIRE(!ix.cosine, KB2_NOT_IMPLEMENTED, "faiss stream: a COSINE index keeps normalised vectors only (the reference stores raw rows + norms)");
    KB2_REQUIRE(ix.shard_world == 1, KB2_NOT_IMPLEMENTED, "serialising a shard");
    o.ntotal = ix.count();
    if (auto* fi = dynamic_cast<FlatIndex*>(&ix)) {
        KB2_REQUIRE(!fi->custom_labels, KB2_NOT_IMPLEMENTED, "faiss stream: FLAT with custom ids");
        o.xb.resize((size_t)o.ntotal * o.d);
        if (o.ntotal) KB2_CUDA_CHECK(cudaMemcpy(o.xb.data(), fi->base.p, o.xb.size() * 4, cudaMemcpyDeviceToHost));
    } else if (auto* iv = dynamic_cast<IvfIndex*>(&ix)) {
        KB2_REQUIRE(iv->trained, KB2_INDEX_NOT_TRAINED, "index not trained");
        iv->seal();
        o.nlist = iv->nlist;
        o.nprobe = 1;
        o.centroids.resize((size_t)iv->nlist * o.d);
        KB2_CUDA_CHECK(cudaMemcpy(o.centroids.data(), iv->centroids.p, o.centroids.size() * 4, cudaMemcpyDeviceToHost));
        o.M = iv->M;
        o.code_size = iv->is_pq ? (uint64_t)iv->M : (uint64_t)o.d * 4;
        if (iv->is_pq) {
            KB2_REQUIRE(iv->nbits == 8, KB2_NOT_IMPLEMENTED, "faiss stream: nbits != 8");
            o.pq_centroids.resize((size_t)256 * o.d);
            KB2_CUDA_CHECK(cudaMemcpy(o.pq_centroids.data(), iv->pqc.p, o.pq_centroids.size() * 4, cudaMemcpyDeviceToHost));
        }
        o.list_ids.assign(iv->nlist, {});
        o.list_codes.assign(iv->nlist, {});
        for (int64_t l = 0; l < iv->nlist; l++) {
            const int64_t len = iv->h_list_len[l];
            if (!len) continue;
            o.list_ids[l].resize(len);
            o.list_codes[l].resize((size_t)len * o.code_size);
            iv->export_list(l, o.list_ids[l].data(), o.list_codes[l].data());
        }
        o.has_refine = iv->is_pq && iv->refine;
        if (o.has_refine) {
            KB2_REQUIRE(!iv->custom_labels, KB2_NOT_IMPLEMENTED, "faiss stream: refine store with custom ids");
            KB2_REQUIRE(iv->refine_kind == 0, KB2_NOT_IMPLEMENTED, "faiss stream: only a flat fp32 refine store is written");
            // refine store in id order: row r lives at position pos_of_row[r]
            DevBuf<float> byrow;
            byrow.ensure((size_t)std::max<int64_t>(o.ntotal, 1) * o.d);
            gather_rows_kernel<<<grid1d(o.ntotal * 32, 256), 256, 0, iv->stream>>>(iv->vecs.p, iv->pos_of_row.p, o.ntotal, o.d, o.d, byrow.p);
            o.refine_xb.resize((size_t)o.ntotal * o.d);
            KB2_CUDA_CHECK(cudaMemcpyAsync(o.refine_xb.data(), byrow.p, o.refine_xb.size() * 4, cudaMemcpyDeviceToHost, iv->stream));
            KB2_CUDA_CHECK(cudaStreamSynchronize(iv->stream));
            o.k_factor = 1.f;
        }
    } else if (auto* hn = dynamic_cast<HnswIndex*>(&ix)) {
        KB2_REQUIRE(!hn->custom_labels, KB2_NOT_IMPLEMENTED, "faiss stream: HNSW with custom ids");
        o.xb = hn->h_vecs;
        o.levels = hn->h_levels;
        o.neighbors = hn->h_neighbors;
        o.offsets.assign(hn->h_offsets.begin(), hn->h_offsets.end());
        o.entry_point = hn->entry_point;
        o.max_level = hn->max_level;
        o.efConstruction = hn->efConstruction;
        // K/impl/HNSW.cpp:78-89 set_default_probas(M, 1 / ln M): the full level table, of which h_cum is a prefix
        const double mult = 1.0 / std::log((double)hn->M);
        int nn = 0;
        o.cum.assign(1, 0);
        for (int level = 0;; level++) {
            const double proba = std::exp(-level / mult) * (1 - std::exp(-1 / mult));
            if (proba < 1e-9) break;
            o.assign_probas.push_back(proba);
            nn += level == 0 ? hn->M * 2 : hn->M;
            o.cum.push_back(nn);
        }
        KB2_REQUIRE(o.cum.size() >= hn->h_cum.size(), KB2_INTERNAL_ERROR, "HNSW level table shorter than the graph's");
        for (size_t i = 0; i < hn->h_cum.size(); i++)
            KB2_REQUIRE(o.cum[i] == hn->h_cum[i], KB2_NOT_IMPLEMENTED, "faiss stream: non-default HNSW link counts");
    } else {
        throw Error(KB2_NOT_IMPLEMENTED, "faiss stream: unknown index class");
    }
}
}  // namespace

int
kb2_faiss_describe(const uint8_t* blob, size_t size, int with_norm, char* json_out, size_t cap) {
    return guarded([&] {
        KB2_REQUIRE(blob && json_out && cap > 0, KB2_INVALID_ARGS, "null argument");
        FaissReader rd{BlobReader{blob, size}, with_norm != 0};
        const FaissIndexData o = rd.read_index();
        const std::string s = faiss_describe(o);
        KB2_REQUIRE(s.size() + 1 <= cap, KB2_INVALID_ARGS, "output buffer too small");
        memcpy(json_out, s.c_str(), s.size() + 1);
    });
}
// parse + re-emit on the host (no device): the writer's output for exactly what the reader understood
int
kb2_faiss_rewrite(const uint8_t* blob, size_t size, int with_norm, uint8_t** out, size_t* out_size) {
    return guarded([&] {
        KB2_REQUIRE(blob && out && out_size, KB2_INVALID_ARGS, "null argument");
        FaissReader rd{BlobReader{blob, size}, with_norm != 0};
        const FaissIndexData o = rd.read_index();
        std::vector<uint8_t> b;
        FaissWriter wr{BlobWriter{b}};
        wr.write_index(o);
        *out = (uint8_t*)malloc(b.size() ? b.size() : 1);
        KB2_REQUIRE(*out != nullptr, KB2_MALLOC_ERROR, "malloc failed");
        memcpy(*out, b.data(), b.size());
        *out_size = b.size();
    });
}
int
kb2_index_deserialize_faiss(const uint8_t* blob, size_t size, int with_norm, int device, kb2_index_t* out) {
    return guarded([&] {
        KB2_REQUIRE(blob && out, KB2_INVALID_ARGS, "null argument");
        *out = nullptr;
        FaissReader rd{BlobReader{blob, size}, with_norm != 0};
        const FaissIndexData o = rd.read_index();   // parse (and reject) before touching the device
        require_device(device);
        std::unique_ptr<IndexBase> ix = index_from_faiss(o, device);
        auto* hh = new Handle();
        hh->ix = std::move(ix);
        *out = reinterpret_cast<kb2_index_t>(hh);
    });
}
int
kb2_index_serialize_faiss(kb2_index_t h, uint8_t** out, size_t* out_size) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        KB2_REQUIRE(out && out_size, KB2_INVALID_ARGS, "null argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        KB2_CUDA_CHECK(cudaSetDevice(ix->device));
        FaissIndexData o;
        index_to_faiss(*ix, o);
        std::vector<uint8_t> blob;
        FaissWriter wr{BlobWriter{blob}};
        wr.write_index(o);
        *out = (uint8_t*)malloc(blob.size() ? blob.size() : 1);
        KB2_REQUIRE(*out != nullptr, KB2_MALLOC_ERROR, "malloc failed");
        memcpy(*out, blob.data(), blob.size());
        *out_size = blob.size();
    });
}
// IndexNode::DeserializeFromFile (index_node.h:329-395): a file holding either container
int
kb2_index_deserialize_from_file(const char* path, int device, kb2_index_t* out) {
    std::vector<uint8_t> buf;
    int st = guarded([&] {
        KB2_REQUIRE(path && out, KB2_INVALID_ARGS, "null argument");
        FILE* f = fopen(path, "rb");
        KB2_REQUIRE(f != nullptr, KB2_INVALID_ARGS, std::string("cannot open ") + path);
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(n > 0 ? (size_t)n : 0);
        const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
        fclose(f);
        KB2_REQUIRE(got == buf.size() && got >= 4, KB2_INVALID_BINARY_SET, "short read");
    });
    if (st != KB2_SUCCESS) return st;
    uint32_t magic;
    memcpy(&magic, buf.data(), 4);
    if (magic == 0x4932424b) return kb2_index_deserialize(buf.data(), buf.size(), device, out);
    return kb2_index_deserialize_faiss(buf.data(), buf.size(), 0, device, out);
}
// IndexNode::GetIndexMeta (index_node.h:329-395): JSON description of the loaded index
int
kb2_index_get_meta(kb2_index_t h, char* json_out, size_t cap) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        KB2_REQUIRE(json_out && cap > 0, KB2_INVALID_ARGS, "null argument");
        std::string s = "{\"type\": \"" + ix->type + "\", \"dim\": " + std::to_string(ix->dim) + ", \"rows\": " + std::to_string(ix->count()) +
                        ", \"metric_type\": \"" + (ix->cosine ? "COSINE" : (ix->metric == KB2_METRIC_IP ? "IP" : "L2")) + "\"" +
                        ", \"size_bytes\": " + std::to_string(ix->size_bytes()) + ", \"device\": " + std::to_string(ix->device) +
                        ", \"shard_rank\": " + std::to_string(ix->shard_rank) + ", \"shard_world\": " + std::to_string(ix->shard_world);
        if (auto* iv = dynamic_cast<IvfIndex*>(ix)) {
            s += ", \"nlist\": " + std::to_string(iv->nlist);
            if (iv->is_pq) s += ", \"m\": " + std::to_string(iv->M) + ", \"nbits\": " + std::to_string(iv->nbits) + ", \"refine\": " + (iv->refine ? "true" : "false");
        } else if (auto* hn = dynamic_cast<HnswIndex*>(ix)) {
            s += ", \"M\": " + std::to_string(hn->M) + ", \"efConstruction\": " + std::to_string(hn->efConstruction) +
                 ", \"max_level\": " + std::to_string(hn->max_level) + ", \"entry_point\": " + std::to_string(hn->entry_point);
        }
        s += "}";
        KB2_REQUIRE(s.size() + 1 <= cap, KB2_INVALID_ARGS, "output buffer too small");
        memcpy(json_out, s.c_str(), s.size() + 1);
    });
}

// ---------------------------------------------------------------- BruteForce
// One scratch FLAT object per device, reused across calls (stream, events, selection scratch); a device-resident base is
// viewed in place (no copy), a host base is staged into the scratch buffer.
namespace {
struct BfSlot {
    std::mutex mu;
    std::unique_ptr<FlatIndex> fi;
};
BfSlot g_bf[64];

FlatIndex&
bf_prepare(BfSlot& slot, int device, int metric, int dim, void* cuda_stream, const float* base, int64_t nb) {
    if (!slot.fi) {
        slot.fi.reset(new FlatIndex());
        slot.fi->type = "FLAT";
        slot.fi->device = device;
        slot.fi->init_common();
    }
    FlatIndex& fi = *slot.fi;
    fi.cosine = (metric == KB2_METRIC_COSINE);
    fi.metric = fi.cosine ? KB2_METRIC_IP : metric;
    fi.dim = dim;
    if (cuda_stream) fi.set_stream((cudaStream_t)cuda_stream); else fi.use_own_stream();
    fi.wait_caller_work();
    fi.last = Counters{};
    const size_t cnt = (size_t)nb * dim;
    if (fi.cosine) {
        const float* dx = fi.to_device(base, cnt, fi.s_cos_in);
        fi.base.ensure(cnt);
        normalize_rows_kernel<<<grid1d(nb * 32, 256), 256, 0, fi.stream>>>(dx, nb, dim, fi.base.p);
    } else if (is_device_ptr(base)) {
        fi.base.borrow(base, cnt);
    } else {
        fi.base.ensure(cnt);
        KB2_CUDA_CHECK(cudaMemcpyAsync(fi.base.p, base, cnt * 4, cudaMemcpyHostToDevice, fi.stream));
        fi.last.h2d += (int64_t)cnt * 4;
    }
    fi.n_used = cnt;
    fi.norms.ensure(nb);
    row_norms_kernel<<<grid1d(nb * 32, 256), 256, 0, fi.stream>>>(fi.base.p, nb, dim, fi.norms.p);
    fi.norms_used = (size_t)nb;
    fi.custom_labels = false;
    fi.n_global_added = nb;
    KB2_CUDA_CHECK(cudaGetLastError());
    return fi;
}
}  // namespace

int
kb2_bruteforce_search(const float* base, int64_t nb, int dim, int metric, const float* queries, int64_t nq, int k,
                      const uint8_t* bitset, int64_t bitset_nbits, int64_t* out_ids, float* out_dist, int device,
                      void* cuda_stream) {
    return guarded([&] {
        KB2_REQUIRE(base && queries && out_ids && out_dist, KB2_INVALID_ARGS, "null buffer");
        KB2_REQUIRE(nb > 0 && dim > 0 && nq >= 0 && k > 0, KB2_INVALID_ARGS, "bad sizes");
        KB2_REQUIRE(metric == KB2_METRIC_L2 || metric == KB2_METRIC_IP || metric == KB2_METRIC_COSINE,
                    KB2_INVALID_METRIC_TYPE, "metric must be L2, IP or COSINE");
        require_device(device);
        KB2_REQUIRE(device < 64, KB2_INVALID_ARGS, "bad device ordinal");
        if (nq == 0) return;
        BfSlot& slot = g_bf[device];
        std::lock_guard<std::mutex> lk(slot.mu);
        FlatIndex& fi = bf_prepare(slot, device, metric, dim, cuda_stream, base, nb);
        JsonObj cfg;
        fi.search(fi.cosine ? fi.normalized(queries, nq) : queries, nq, k, cfg, bitset, bitset_nbits, out_ids, out_dist);
        fi.base.release();   // never keep a view of caller memory (or a stale copy) between calls
        fi.n_used = 0;
    });
}
int
kb2_bruteforce_range_search(const float* base, int64_t nb, int dim, int metric, const float* queries, int64_t nq,
                            float radius, float range_filter, int has_range_filter, const uint8_t* bitset,
                            int64_t bitset_nbits, int64_t** out_lims, int64_t** out_ids, float** out_dist, int device,
                            void* cuda_stream) {
    return guarded([&] {
        KB2_REQUIRE(base && queries, KB2_INVALID_ARGS, "null buffer");
        KB2_REQUIRE(nb > 0 && dim > 0 && nq >= 0, KB2_INVALID_ARGS, "bad sizes");
        KB2_REQUIRE(metric == KB2_METRIC_L2 || metric == KB2_METRIC_IP || metric == KB2_METRIC_COSINE,
                    KB2_INVALID_METRIC_TYPE, "metric must be L2, IP or COSINE");
        require_device(device);
        KB2_REQUIRE(device < 64, KB2_INVALID_ARGS, "bad device ordinal");
        BfSlot& slot = g_bf[device];
        std::lock_guard<std::mutex> lk(slot.mu);
        FlatIndex& fi = bf_prepare(slot, device, metric, dim, cuda_stream, base, nb);
        JsonObj cfg;
        range_search_index(fi, fi.cosine ? fi.normalized(queries, nq) : queries, nq, radius, range_filter,
                           has_range_filter != 0, cfg, bitset, bitset_nbits, out_lims, out_ids, out_dist);
        fi.base.release();
        fi.n_used = 0;
    });
}

// ---------------------------------------------------------------- multi-GPU: NCCL communicator behind the ABI
int
kb2_comm_unique_id(uint8_t* out128) {
    return guarded([&] {
        KB2_REQUIRE(out128 != nullptr, KB2_INVALID_ARGS, "null buffer");
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
        ncclUniqueId id;
        KB2_NCCL_CHECK(Comm::api().GetUniqueId(&id));
        memcpy(out128, &id, 128);
    });
}
int
kb2_comm_create(const uint8_t* id128, int rank, int world, int device, kb2_comm_t* out) {
    return guarded([&] {
        KB2_REQUIRE(id128 && out, KB2_INVALID_ARGS, "null argument");
        KB2_REQUIRE(world >= 1 && rank >= 0 && rank < world, KB2_INVALID_ARGS, "bad rank/world");
        *out = nullptr;
        require_device(device);
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        std::unique_ptr<Comm> c(new Comm());
        c->rank = rank;
        c->world = world;
        c->device = device;
        KB2_NCCL_CHECK(Comm::api().CommInitRank(&c->comm, world, id, rank));
        *out = reinterpret_cast<kb2_comm_t>(c.release());
    });
}
void
kb2_comm_destroy(kb2_comm_t c) {
    delete reinterpret_cast<Comm*>(c);
}
int
kb2_comm_all_gather(kb2_comm_t c, const void* send, void* recv, size_t bytes, void* cuda_stream) {
    return guarded([&] {
        KB2_REQUIRE(c && send && recv, KB2_INVALID_ARGS, "null argument");
        Comm* cc = reinterpret_cast<Comm*>(c);
        KB2_CUDA_CHECK(cudaSetDevice(cc->device));
        cc->all_gather(send, recv, bytes, (cudaStream_t)cuda_stream);
    });
}
int
kb2_index_set_comm(kb2_index_t h, kb2_comm_t c) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        std::lock_guard<std::mutex> lk(ix->mu);
        Comm* cc = reinterpret_cast<Comm*>(c);
        if (cc) {
            KB2_REQUIRE(cc->rank == ix->shard_rank && cc->world == ix->shard_world, KB2_INVALID_ARGS,
                        "communicator rank/world differ from kb2_index_set_shard");
            KB2_REQUIRE(cc->device == ix->device, KB2_INVALID_ARGS, "communicator lives on another device");
        }
        ix->set_comm(cc);
    });
}

// ---------------------------------------------------------------- multi-GPU merge
int
kb2_merge_topk(int metric, int world, int64_t nq, int k, const int64_t* in_ids, const float* in_dist, int64_t* out_ids,
               float* out_dist, int device, void* cuda_stream) {
    return guarded([&] {
        KB2_REQUIRE(in_ids && in_dist && out_ids && out_dist, KB2_INVALID_ARGS, "null buffer");
        KB2_REQUIRE(world >= 1 && k >= 1 && (int64_t)world * k <= kMaxSortEntries, KB2_INVALID_ARGS, "world*k too large");
        require_device(device);
        merge_topk_device(metric, world, nq, k, in_ids, in_dist, out_ids, out_dist, (cudaStream_t)cuda_stream);
    });
}

// ---------------------------------------------------------------- validation hook for the two contractions
int
kb2_debug_gemm_keys(const float* q, int64_t nq, const float* x, int64_t nb, int dim, int metric, int use_tc, float* out_keys,
                    int device) {
    return guarded([&] {
        require_device(device);
        KB2_REQUIRE(is_device_ptr(q) && is_device_ptr(x) && is_device_ptr(out_keys), KB2_INVALID_ARGS,
                    "debug_gemm_keys takes device pointers");
        const int64_t ldk = (nb + 3) & ~(int64_t)3;
        DevBuf<float> qn, xn;
        qn.ensure(nq);
        xn.ensure(nb);
        row_norms_kernel<<<grid1d(nq * 32, 256), 256>>>(q, nq, dim, qn.p);
        row_norms_kernel<<<grid1d(nb * 32, 256), 256>>>(x, nb, dim, xn.p);
        const bool ran_tc = launch_gemm_keys(nullptr, use_tc ? 1 : 0, metric, q, x, qn.p, xn.p, (int)nq, (int)nb, dim, out_keys,
                                             ldk, nullptr, nullptr, 0);
        KB2_CUDA_CHECK(cudaGetLastError());
        KB2_CUDA_CHECK(cudaDeviceSynchronize());
        KB2_REQUIRE(!use_tc || ran_tc, KB2_INTERNAL_ERROR, "tensor-core contraction unavailable (tensor map / alignment)");
    });
}

// ---------------------------------------------------------------- introspection
int
kb2_index_last_search_counters(kb2_index_t h, int64_t* out8) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        const Counters& c = ix->last;
        out8[0] = c.launches;
        out8[1] = c.codes;
        out8[2] = c.code_bytes;
        out8[3] = c.pairs;
        out8[4] = c.h2d;
        out8[5] = c.d2h;
        out8[6] = c.survivors;
        out8[7] = c.flagged;
    });
}
int
kb2_index_enable_kernel_timing(kb2_index_t h, int on) {
    return guarded([&] { ix_of(h)->timing = (on != 0); });
}
int
kb2_index_last_kernel_ms(kb2_index_t h, float* out_ms) {
    return guarded([&] { *out_ms = ix_of(h)->last_kernel_ms; });
}
int
kb2_index_last_stage_info(kb2_index_t h, float* out4) {
    return guarded([&] {
        IndexBase* ix = ix_of(h);
        out4[0] = ix->last_stage_ms;
        out4[1] = ix->last_kernel_ms;
        out4[2] = (float)ix->last_engine;
        out4[3] = ix->last_comm_ms;
    });
}

}  // extern "C"
