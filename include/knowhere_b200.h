/*
 * knowhere_b200.h — C ABI of the B200-native ANN search core (libknowhere_b200.so).
 *
 * This is the drop-in boundary.  Every entry point is plain C (pointers + sizes, no
 * C++/torch types) and names the reference interface it stands in for.  A Knowhere
 * build binds these from an IndexNode subclass exactly like the in-tree GPU precedent
 * binds cuVS through a pimpl (reference: src/index/gpu_cuvs/gpu_cuvs.h:71-316,
 * src/common/cuvs/integration/cuvs_knowhere_index.hpp:28-75); see INTEGRATION.md.
 *
 * Conventions
 *   - Return value is a knowhere::Status integer (reference: include/knowhere/expected.h:34-68);
 *     0 = success.  kb2_last_error() returns the thread-local message of the last failure.
 *   - No exception ever crosses this boundary (reference: GuardedCall, expected.h:408-430).
 *   - Vector / query / output pointers may be HOST or DEVICE pointers; the library detects
 *     which (cudaPointerGetAttributes).  Host buffers are staged through pinned memory and
 *     copied on the index's stream inside the call (that is the end-to-end path);
 *     device buffers are used in place (that is the HBM-resident path).
 *   - Results: k entries per query, best first (L2 ascending, IP descending), ties ordered by
 *     ascending id; missing entries have id -1 and distance +FLT_MAX (L2) / -FLT_MAX (IP)
 *     (reference: F/utils/ordered_key_value.h:54-79, K/impl/HnswSearcher.h:414-428).
 *   - ids are int64 (faiss::idx_t; reference include/knowhere/dataset.h:499-510).
 *   - Bitset: bit i set => row i is filtered OUT (reference include/knowhere/bitsetview.h:166);
 *     byte i/8, bit i%8, host or device pointer, NULL = no filter.
 *   - Search on one handle is thread-safe w.r.t. other searches (internally serialised on the
 *     handle's stream; reference: IndexNodeThreadPoolWrapper, index_node_thread_pool_wrapper.cc:33-44).
 *   - There is NO CPU fallback: with no usable CUDA device every call returns
 *     KB2_CUDA_RUNTIME_ERROR (reference: index_factory.cc:29-45,62-66).
 */
#ifndef KNOWHERE_B200_H
#define KNOWHERE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* knowhere::Status values used by this library (expected.h:34-68) */
enum {
    KB2_SUCCESS = 0,
    KB2_INVALID_ARGS = 1,
    KB2_INVALID_PARAM_IN_JSON = 2,
    KB2_OUT_OF_RANGE_IN_JSON = 3,
    KB2_INVALID_METRIC_TYPE = 5,
    KB2_EMPTY_INDEX = 6,
    KB2_NOT_IMPLEMENTED = 7,
    KB2_INDEX_NOT_TRAINED = 8,
    KB2_INDEX_ALREADY_TRAINED = 9,
    KB2_MALLOC_ERROR = 13,
    KB2_INVALID_BINARY_SET = 19,
    KB2_CUDA_RUNTIME_ERROR = 22,
    KB2_INTERNAL_ERROR = 27
};

/* metric ids (reference: include/knowhere/comp/index_param.h metric names "L2","IP","COSINE") */
enum { KB2_METRIC_L2 = 0, KB2_METRIC_IP = 1, KB2_METRIC_COSINE = 2 };

typedef struct kb2_index* kb2_index_t;
typedef struct kb2_comm* kb2_comm_t;

/* ---- library ------------------------------------------------------------------------- */
const char* kb2_version(void);
const char* kb2_last_error(void);
/* number of visible CUDA devices with compute capability 10.x; <=0 => library unusable */
int kb2_device_count(void);

/* ---- index lifecycle ------------------------------------------------------------------
 * Replaces IndexFactory::Create<fp32>(name, version) + IndexNode ctor
 * (reference: include/knowhere/index/index_factory.h:27-72, src/index/index_factory.cc:48-86).
 * index_type: "FLAT" | "IVF_FLAT" | "IVF_PQ" | "HNSW"    (index_param.h:27-46)
 * json_cfg  : build-time keys of the reference configs: metric_type, dim, nlist, m, nbits,
 *             refine, refine_type ("flat"|"fp32"), M, efConstruction.  metric KB2_METRIC_COSINE: vectors are
 *             L2-normalised on entry and queries at search (then inner product); HasRawData is false.
 *             (ivf_config.h:25-128, base_hnsw_config.h:36-62).  May be NULL/"" for defaults. */
int kb2_index_create(const char* index_type, int metric, int dim, const char* json_cfg, int device,
                     kb2_index_t* out);
void kb2_index_destroy(kb2_index_t h);

/* run all of this handle's work on an externally owned cudaStream_t (e.g. torch's current
 * stream) so that callers can bracket it with their own CUDA events.  0 = legacy default stream. */
int kb2_index_set_stream(kb2_index_t h, void* cuda_stream);

/* Multi-GPU list/row sharding.  Must be called before train/add/import.
 * IVF_*: every inverted list lives on exactly one rank (greedy size-balanced packing over the global list sizes,
 * identical on all ranks; KB2_SHARD_POLICY=mod selects l % world).  FLAT: row i is kept by rank
 * floor(i * world / n) at add time.  HNSW: graph partitions — every rank builds an independent sub-graph over its
 * contiguous row slice of the (single) add() call and all of them are searched with the same ef (SURVEY §8e option 2;
 * recall >= the single graph's in practice, at world x the distance evaluations).
 * Without a communicator Search returns this shard's local top-k (merge shards with kb2_merge_topk after an
 * all-gather); with kb2_index_set_comm the gather + merge happen inside Search. */
int kb2_index_set_shard(kb2_index_t h, int rank, int world);

/* IndexNode::Train (reference: include/knowhere/index/index_node.h:131, ivf.cc:545-807):
 * k-means coarse quantizer (niter 25, <=256 pts/centroid, seed 1234) and, for IVF_PQ,
 * M residual sub-quantizer codebooks.  FLAT / HNSW: no-op. */
int kb2_index_train(kb2_index_t h, const float* x, int64_t n);
/* IndexNode::Add (index_node.h:140-146, ivf.cc:809-844, flat.cc Add).  ids==NULL => sequential
 * labels continuing from Count(). */
int kb2_index_add(kb2_index_t h, const float* x, int64_t n, const int64_t* ids);

/* Typed variants: element type of x / queries.  The reference registers FLAT / IVF_* for fp16, bf16 and int8 through a
 * wrapper that converts the dataset and every query batch to fp32 (index_factory.h:95-103,
 * index_node_data_mock_wrapper.cc:24-60); these entry points do that conversion on the device.  Distances are fp32. */
enum { KB2_DTYPE_F32 = 0, KB2_DTYPE_F16 = 1, KB2_DTYPE_BF16 = 2, KB2_DTYPE_INT8 = 3 };
int kb2_index_train_typed(kb2_index_t h, const void* x, int dtype, int64_t n);
int kb2_index_add_typed(kb2_index_t h, const void* x, int dtype, int64_t n, const int64_t* ids);
int kb2_index_search_typed(kb2_index_t h, const void* queries, int dtype, int64_t nq, int k, const char* json,
                           const uint8_t* bitset, int64_t bitset_nbits, int64_t* out_ids, float* out_dist);

/* IndexNode::Search (index_node.h:164-166; ivf.cc:887-1168; flat.cc:75-152;
 * faiss_hnsw.cc:1344-1527).  json: search keys k is passed explicitly; nprobe, ef,
 * refine_k (ivf_config.h:33-45,97-128; base_hnsw_config.h:40-71).
 * out_ids[nq*k], out_dist[nq*k] are caller-allocated (like BruteForce::SearchWithBuf,
 * include/knowhere/comp/brute_force.h:33-36). */
int kb2_index_search(kb2_index_t h, const float* queries, int64_t nq, int k, const char* json,
                     const uint8_t* bitset, int64_t bitset_nbits, int64_t* out_ids, float* out_dist);

/* IndexNode::RangeSearch (index_node.h:253-255; flat.cc:154-234; ivf.cc:1229-1500;
 * include/knowhere/range_util.h:23-26): L2 keeps radius > d >= range_filter,
 * IP keeps radius < d <= range_filter.  has_range_filter=0 => one-sided.
 * Outputs are malloc'ed HOST arrays owned by the caller (free with kb2_free):
 * lims[nq+1], ids[lims[nq]], dist[lims[nq]], each query's hits sorted best-first. */
int kb2_index_range_search(kb2_index_t h, const float* queries, int64_t nq, float radius,
                           float range_filter, int has_range_filter, const char* json,
                           const uint8_t* bitset, int64_t bitset_nbits, int64_t** out_lims,
                           int64_t** out_ids, float** out_dist);
void kb2_free(void* p);

/* IndexNode::Count / Dim / Size / HasRawData / GetVectorByIds (index_node.h:270-279,329-395) */
int64_t kb2_index_count(kb2_index_t h);
int kb2_index_dim(kb2_index_t h);
int64_t kb2_index_size_bytes(kb2_index_t h);
int kb2_index_is_trained(kb2_index_t h);
int kb2_index_has_raw_data(kb2_index_t h);
int kb2_index_get_vector_by_ids(kb2_index_t h, const int64_t* ids, int64_t n, float* out);

/* ---- importing an index built elsewhere (a Milvus/faiss CPU-built index) ---------------
 * These replace the Deserialize path for already-parsed faiss structures
 * (reference: K/impl/index_read.cpp IwFl/IwPQ/IHNf readers) and are what
 * the parity tests use so that GPU and CPU search the *same* trained index.
 * IVF: centroids[nlist*dim]; pq_centroids[M*ksub*dsub] (NULL for IVF_FLAT).  Then one
 * kb2_ivf_import_list per inverted list (codes: list_size*code_size bytes where code_size is
 * M for IVF_PQ nbits=8, dim*4 for IVF_FLAT), then kb2_ivf_import_finish.  raw (optional,
 * n*dim fp32 in label order 0..n-1) enables refine. */
int kb2_ivf_import_begin(kb2_index_t h, int64_t nlist, const float* centroids, const float* pq_centroids);
int kb2_ivf_import_list(kb2_index_t h, int64_t list_no, int64_t list_size, const int64_t* ids,
                        const uint8_t* codes);
int kb2_ivf_import_finish(kb2_index_t h, const float* raw, int64_t n_raw);
/* export (for serialisation to the faiss wire format by the caller / for oracle checks) */
int64_t kb2_ivf_nlist(kb2_index_t h);
int64_t kb2_ivf_list_size(kb2_index_t h, int64_t list_no);
int kb2_ivf_export_centroids(kb2_index_t h, float* centroids, float* pq_centroids);
int kb2_ivf_export_list(kb2_index_t h, int64_t list_no, int64_t* ids, uint8_t* codes);

/* HNSW graph import/export in the reference's own layout (K/impl/HNSW.h: levels[n] (level+1 per
 * node), offsets[n+1], neighbors[offsets[n]] int32 with -1 padding, cum_nneighbor_per_level,
 * entry_point, max_level; K/impl/HNSW.cpp:53-89,202-225).  vectors: n*dim fp32. */
int kb2_hnsw_import(kb2_index_t h, int64_t n, const float* vectors, const int32_t* levels,
                    const int64_t* offsets, const int32_t* neighbors, const int32_t* cum_nneighbor,
                    int n_cum, int32_t entry_point, int32_t max_level);
int kb2_hnsw_export_meta(kb2_index_t h, int64_t* out5 /* n, entry_point, max_level, n_links, n_cum */);
int kb2_hnsw_export(kb2_index_t h, int32_t* levels, int64_t* offsets, int32_t* neighbors, int32_t* cum);
/* per-search statistics of the last kb2_index_search on an HNSW handle
 * (reference HNSWStats: K/impl/HnswSearcher.h:284-288): out2 = {ndis, nhops} summed over queries */
int kb2_hnsw_last_stats(kb2_index_t h, int64_t* out2);

/* ---- Serialize / Deserialize (index_node.h:337-367; BinarySet payload) -------------------
 * Self-describing little-endian blob ("KB2I" container); *out is malloc'ed (kb2_free). */
int kb2_index_serialize(kb2_index_t h, uint8_t** out, size_t* out_size);
int kb2_index_deserialize(const uint8_t* blob, size_t size, int device, kb2_index_t* out);

/* ---- the reference's wire format: faiss fourcc streams, i.e. the payload Knowhere stores in a BinarySet under the
 * index type name (flat.cc:323-343, ivf.cc:1717-1741, faiss_hnsw.cc:188-217; K/impl/index_write.cpp:523-560,716-745,
 * 776-822, K/impl/index_read.cpp).  Supported: "IxF2"/"IxFI"/"IxF9" (FLAT), "IwFl" (IVF_FLAT), "IwPQ" and "IxRF" over it
 * (IVF_PQ, + flat fp32 refine store), "IHNf"/"IHN9" (HNSW over flat storage).  with_norm = 1 when the inverted lists
 * carry per-row norms (the reference's IO_FLAG_WITH_NORM).  A CPU-built Knowhere index loads straight onto the GPU and
 * a GPU-built index can be served by the reference's CPU nodes.  kb2_faiss_describe parses on the host only (no device
 * needed) and returns a one-line JSON description. */
int kb2_faiss_describe(const uint8_t* blob, size_t size, int with_norm, char* json_out, size_t cap);
/* host-only: parse the stream and write it again with this library's writer (out: malloc'ed, kb2_free) */
int kb2_faiss_rewrite(const uint8_t* blob, size_t size, int with_norm, uint8_t** out, size_t* out_size);
int kb2_index_deserialize_faiss(const uint8_t* blob, size_t size, int with_norm, int device, kb2_index_t* out);
int kb2_index_serialize_faiss(kb2_index_t h, uint8_t** out, size_t* out_size);
/* IndexNode::DeserializeFromFile / GetIndexMeta (include/knowhere/index/index_node.h:329-395): the file may hold a
 * faiss stream or this library's "KB2I" container; the meta is a JSON object (type, dim, rows, metric_type, ...) */
int kb2_index_deserialize_from_file(const char* path, int device, kb2_index_t* out);
int kb2_index_get_meta(kb2_index_t h, char* json_out, size_t cap);

/* ---- index-less exact search: knowhere::BruteForce (include/knowhere/comp/brute_force.h:26-69;
 * src/common/comp/brute_force.cc:260-392,588-710) */
int kb2_bruteforce_search(const float* base, int64_t nb, int dim, int metric, const float* queries,
                          int64_t nq, int k, const uint8_t* bitset, int64_t bitset_nbits,
                          int64_t* out_ids, float* out_dist, int device, void* cuda_stream);
int kb2_bruteforce_range_search(const float* base, int64_t nb, int dim, int metric, const float* queries,
                                int64_t nq, float radius, float range_filter, int has_range_filter,
                                const uint8_t* bitset, int64_t bitset_nbits, int64_t** out_lims,
                                int64_t** out_ids, float** out_dist, int device, void* cuda_stream);

/* ---- multi-GPU: one process per GPU, inverted lists sharded (kb2_index_set_shard), collectives over NCCL/NVLink
 * INSIDE the library (the reference has no multi-GPU path: one index per device,
 * src/common/cuvs/integration/cuvs_knowhere_index.cuh:415-460).  Bootstrap as in NCCL: rank 0 obtains a 128-byte id,
 * the host application distributes it (its own RPC; torch.distributed in the tests), every rank creates its
 * communicator on its device.  After kb2_index_set_comm, kb2_index_search on a sharded IVF index is a COLLECTIVE call:
 * every rank passes the same query batch and receives the same global top-k.  Per call: the coarse quantizer runs on
 * 1/world of the batch per rank + one all-gather of the probe lists; (tensor-core IVF_PQ engine) phase-A bounds are
 * computed by the rank owning each query's nearest list + one min all-reduce; every rank scans its own lists; ONE
 * all-gather of the per-shard top-k candidates + the merge kernel.  NCCL is resolved with dlopen("libnccl.so.2"). */
int kb2_comm_unique_id(uint8_t* out128);
int kb2_comm_create(const uint8_t* id128, int rank, int world, int device, kb2_comm_t* out);
void kb2_comm_destroy(kb2_comm_t c);
/* every rank contributes `bytes` device bytes; recv holds world*bytes, rank-major (exposed for tests / host glue) */
int kb2_comm_all_gather(kb2_comm_t c, const void* send, void* recv, size_t bytes, void* cuda_stream);
/* attach (or detach with NULL) a communicator; rank/world/device must equal the handle's shard settings.  The
 * communicator is not owned by the index and must outlive it. */
int kb2_index_set_comm(kb2_index_t h, kb2_comm_t c);

/* ---- multi-GPU candidate merge (the kernel that consumes the NCCL all-gather) ---
 * in_ids/in_dist: [world][nq][k] gathered per-shard results (device or host);
 * out: [nq][k] global top-k, same ordering rules as search. */
int kb2_merge_topk(int metric, int world, int64_t nq, int k, const int64_t* in_ids, const float* in_dist,
                   int64_t* out_ids, float* out_dist, int device, void* cuda_stream);

/* ---- introspection used by bench.py for the roofline figures --------------------------- */
/* fills out[0..7] with counters of the last search on this handle:
 * [0] kernels launched, [1] codes (rows) scanned, [2] algorithmic code bytes scanned,
 * [3] (query,list) pairs, [4] H2D bytes, [5] D2H bytes, [6] IVF_PQ tensor-core engine: codes re-evaluated exactly (survivors of
 * the bf16 filter), [7] queries redone by the LUT kernel (no bound / survivor-buffer overflow) */
int kb2_index_last_search_counters(kb2_index_t h, int64_t* out8);
/* device time in milliseconds of the dominant scan kernel of the last search, measured with
 * CUDA events on the handle's stream (valid only after kb2_index_enable_kernel_timing(h,1)) */
int kb2_index_enable_kernel_timing(kb2_index_t h, int on);
int kb2_index_last_kernel_ms(kb2_index_t h, float* out_ms);
/* out4: [0] device ms of the whole list-scan stage of the last search, [1] device ms of its dominant kernel
 * (== kb2_index_last_kernel_ms), [2] engine that served it: 0 = query-major scan kernels, 1 = list-major
 * tensor-core engine (IVF_PQ m=16 d=128 with large batches; kb2_ivfpq_tc.cuh), [3] device ms of the collectives
 * (+ merge kernel) of a sharded search with a communicator */
int kb2_index_last_stage_info(kb2_index_t h, float* out4);

/* validation hook: writes the full key matrix [nq][round_up(nb,4)] of the dense contraction
 * (|q|^2+|x|^2-2qx for L2, -qx for IP) computed by the fp32 CUDA-core kernel (use_tc=0) or by the
 * tcgen05 tensor-core kernel (use_tc=1).  Device pointers only.  Used by tests to hold the tensor-core
 * path to the fp32 one (the reference computes these distances with src/simd fvec_L2sqr_ny). */
int kb2_debug_gemm_keys(const float* q, int64_t nq, const float* x, int64_t nb, int dim, int metric, int use_tc,
                        float* out_keys, int device);

#ifdef __cplusplus
}
#endif
#endif /* KNOWHERE_B200_H */
