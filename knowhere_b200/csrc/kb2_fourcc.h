// kb2_fourcc.h — the reference's wire format: faiss fourcc streams as Knowhere stores them in a BinarySet
// (host-only code).
//
// Knowhere::Serialize writes ONE binary named after the index type whose payload is
// faiss::cppcontrib::knowhere::write_index(index) (flat.cc:323-343, ivf.cc:1717-1741, faiss_hnsw.cc:188-217):
//   FLAT      "IxF2" / "IxFI" (L2 / IP) or "IxF9" (cosine: raw vectors + L2 norms)        K/impl/index_write.cpp:537-558
//   IVF_FLAT  "IwFl"  ivf header + "ilar" inverted lists (codes = raw fp32 rows)            :716-727, :255-308
//   IVF_PQ    "IwPQ"  ivf header + by_residual + code_size + ProductQuantizer + "ilar";     :738-745
//             wrapped in "IxRF" (base, refine "IxF2"/"IxFI", k_factor) when refine is on    :776-781
//   HNSW      "IHNf" / "IHN9" header + HNSW graph + flat storage                            :782-822, :408-420
// index header (:80-103): d i32, ntotal i64, is_cosine u8, 3 x u8, u32, i64 (reserved), is_trained u8, metric i32
// (faiss: 0 = INNER_PRODUCT, 1 = L2), [metric_arg f32 when metric > 1].
// ivf header (:523-531): index header, nlist u64, nprobe u64, quantizer index, direct map (type u8, array vec<i64>).
// vectors are  u64 count + payload; "xb" vectors count floats as count/4-byte units the same way.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "kb2_blob.h"

namespace kb2 {

constexpr uint32_t
fourcc(const char (&s)[5]) {
    return (uint32_t)(uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) | ((uint32_t)(uint8_t)s[3] << 24);
}

// parsed contents of one faiss stream (only the index families of SURVEY §8)
struct FaissIndexData {
    std::string kind;   // "FLAT" | "IVF_FLAT" | "IVF_PQ" | "HNSW"
    int d = 0;
    int64_t ntotal = 0;
    int metric = KB2_METRIC_L2;   // KB2_METRIC_* (cosine flag separately)
    bool cosine = false;
    // FLAT / HNSW storage / refine store
    std::vector<float> xb;
    std::vector<float> xb_norms;   // IxF9 only
    // IVF
    int64_t nlist = 0, nprobe = 1;
    std::vector<float> centroids;
    int M = 0, nbits = 8;
    uint64_t code_size = 0;
    std::vector<float> pq_centroids;
    std::vector<std::vector<int64_t>> list_ids;
    std::vector<std::vector<uint8_t>> list_codes;
    bool has_refine = false;
    float k_factor = 1.f;
    std::vector<float> refine_xb;
    // HNSW
    std::vector<double> assign_probas;
    std::vector<int32_t> cum, levels, neighbors;
    std::vector<uint64_t> offsets;
    int32_t entry_point = -1, max_level = -1, efConstruction = 40, efSearch = 16, upper_beam = 1;
};

struct FaissReader {
    BlobReader r;
    bool with_norm = false;   // IO_FLAG_WITH_NORM: inverted lists carry one float norm per row (cosine IVF_FLAT)

    template <typename T>
    void
    read_vec(std::vector<T>& v, size_t elem_limit = ~(size_t)0) {
        const uint64_t n = r.get<uint64_t>();
        KB2_REQUIRE(n <= (r.n - r.o) / sizeof(T) && n <= elem_limit, KB2_INVALID_BINARY_SET, "faiss stream: vector longer than the stream");
        v.resize(n);
        if (n) memcpy(v.data(), r.get_bytes(n * sizeof(T)), n * sizeof(T));
    }
    struct Header {
        int d;
        int64_t ntotal;
        bool cosine, trained;
        int metric_faiss;
    };
    Header
    read_header() {
        Header h;
        h.d = r.get<int32_t>();
        h.ntotal = r.get<int64_t>();
        h.cosine = r.get<uint8_t>() != 0;
        r.get_bytes(3 + 4 + 8);
        h.trained = r.get<uint8_t>() != 0;
        h.metric_faiss = r.get<int32_t>();
        if (h.metric_faiss > 1) r.get<float>();
        KB2_REQUIRE(h.d > 0 && h.d <= (1 << 20) && h.ntotal >= 0, KB2_INVALID_BINARY_SET, "faiss stream: bad index header");
        KB2_REQUIRE(h.metric_faiss == 0 || h.metric_faiss == 1, KB2_INVALID_METRIC_TYPE, "faiss stream: only L2 / IP are supported");
        return h;
    }
    // "IxF2" / "IxFI" / "IxF9" after the fourcc
    void
    read_flat_body(uint32_t h4, Header& hd, std::vector<float>& xb, std::vector<float>* norms) {
        hd = read_header();
        read_vec(xb);
        KB2_REQUIRE((int64_t)xb.size() == hd.ntotal * hd.d, KB2_INVALID_BINARY_SET, "faiss stream: flat storage size mismatch");
        if (h4 == fourcc("IxF9")) {
            std::vector<float> nr;
            read_vec(nr);
            if (norms) *norms = std::move(nr);
            hd.cosine = true;
        }
    }
    static bool is_flat(uint32_t h) { return h == fourcc("IxF2") || h == fourcc("IxFI") || h == fourcc("IxF9") || h == fourcc("IxFl"); }

    void
    read_invlists(FaissIndexData& o, size_t expect_code_size) {
        const uint32_t h = r.get<uint32_t>();
        KB2_REQUIRE(h == fourcc("ilar"), KB2_INVALID_BINARY_SET, "faiss stream: only ArrayInvertedLists ('ilar') are supported");
        const uint64_t nl = r.get<uint64_t>();
        const uint64_t cs = r.get<uint64_t>();
        KB2_REQUIRE((int64_t)nl == o.nlist && cs == expect_code_size, KB2_INVALID_BINARY_SET, "faiss stream: inverted-list geometry mismatch");
        const uint32_t lt = r.get<uint32_t>();
        std::vector<uint64_t> sizes(nl, 0), raw;
        read_vec(raw);
        if (lt == fourcc("full")) {
            KB2_REQUIRE(raw.size() == nl, KB2_INVALID_BINARY_SET, "faiss stream: list size table");
            sizes = raw;
        } else if (lt == fourcc("sprs")) {
            KB2_REQUIRE(raw.size() % 2 == 0, KB2_INVALID_BINARY_SET, "faiss stream: sparse list size table");
            for (size_t i = 0; i < raw.size(); i += 2) {
                KB2_REQUIRE(raw[i] < nl, KB2_INVALID_BINARY_SET, "faiss stream: list number out of range");
                sizes[raw[i]] = raw[i + 1];
            }
        } else {
            throw Error(KB2_INVALID_BINARY_SET, "faiss stream: unknown list size encoding");
        }
        o.list_ids.assign(nl, {});
        o.list_codes.assign(nl, {});
        o.code_size = cs;
        uint64_t tot = 0;
        for (uint64_t l = 0; l < nl; l++) {
            const uint64_t n = sizes[l];
            if (!n) continue;
            KB2_REQUIRE(n <= (r.n - r.o) / (cs + 8), KB2_INVALID_BINARY_SET, "faiss stream: list longer than the stream");
            o.list_codes[l].resize(n * cs);
            memcpy(o.list_codes[l].data(), r.get_bytes(n * cs), n * cs);
            o.list_ids[l].resize(n);
            memcpy(o.list_ids[l].data(), r.get_bytes(n * 8), n * 8);
            if (with_norm) r.get_bytes(n * 4);   // row norms: recomputed from the vectors on import
            tot += n;
        }
        KB2_REQUIRE((int64_t)tot == o.ntotal, KB2_INVALID_BINARY_SET, "faiss stream: list sizes do not add up to ntotal");
    }
    void
    read_ivf_header(FaissIndexData& o) {
        Header hd = read_header();
        o.d = hd.d;
        o.ntotal = hd.ntotal;
        o.metric = hd.metric_faiss == 0 ? KB2_METRIC_IP : KB2_METRIC_L2;
        o.cosine = hd.cosine;
        o.nlist = (int64_t)r.get<uint64_t>();
        o.nprobe = (int64_t)r.get<uint64_t>();
        KB2_REQUIRE(o.nlist >= 1 && o.nlist <= (1ll << 26), KB2_INVALID_BINARY_SET, "faiss stream: bad nlist");
        const uint32_t qh = r.get<uint32_t>();
        KB2_REQUIRE(is_flat(qh), KB2_INVALID_BINARY_SET, "faiss stream: the coarse quantizer must be a flat index");
        Header qd;
        read_flat_body(qh, qd, o.centroids, nullptr);
        KB2_REQUIRE(qd.d == o.d && qd.ntotal == o.nlist, KB2_INVALID_BINARY_SET, "faiss stream: quantizer geometry mismatch");
        const uint8_t dm_type = r.get<uint8_t>();
        std::vector<int64_t> dm;
        read_vec(dm);
        if (dm_type == 2) {   // hashtable: vector of (idx_t, idx_t) pairs
            const uint64_t n = r.get<uint64_t>();
            r.get_bytes(n * 16);
        }
    }

    FaissIndexData
    read_index() {
        FaissIndexData o;
        uint32_t h = r.get<uint32_t>();
        if (h == fourcc("IxRF")) {
            Header hd = read_header();
            FaissIndexData base = read_index();
            KB2_REQUIRE(base.kind == "IVF_PQ", KB2_NOT_IMPLEMENTED, "faiss stream: IndexRefine is supported over IVF_PQ only");
            const uint32_t rh = r.get<uint32_t>();
            KB2_REQUIRE(rh == fourcc("IxF2") || rh == fourcc("IxFI"), KB2_NOT_IMPLEMENTED,
                        "faiss stream: only a flat fp32 refine store is supported (refine_type=flat)");
            Header rd;
            read_flat_body(rh, rd, base.refine_xb, nullptr);
            KB2_REQUIRE(rd.d == base.d && rd.ntotal == base.ntotal && hd.d == base.d, KB2_INVALID_BINARY_SET,
                        "faiss stream: refine store geometry mismatch");
            base.has_refine = true;
            base.k_factor = r.get<float>();
            return base;
        }
        if (is_flat(h)) {
            Header hd;
            read_flat_body(h, hd, o.xb, &o.xb_norms);
            o.kind = "FLAT";
            o.d = hd.d;
            o.ntotal = hd.ntotal;
            o.metric = hd.metric_faiss == 0 ? KB2_METRIC_IP : KB2_METRIC_L2;
            o.cosine = hd.cosine;
            return o;
        }
        if (h == fourcc("IwFl")) {
            o.kind = "IVF_FLAT";
            read_ivf_header(o);
            read_invlists(o, (size_t)o.d * 4);
            return o;
        }
        if (h == fourcc("IwPQ")) {
            o.kind = "IVF_PQ";
            read_ivf_header(o);
            const uint8_t by_residual = r.get<uint8_t>();
            KB2_REQUIRE(by_residual != 0, KB2_NOT_IMPLEMENTED, "faiss stream: IVF_PQ without by_residual");
            const uint64_t cs = r.get<uint64_t>();
            const uint64_t pd = r.get<uint64_t>(), pm = r.get<uint64_t>(), pb = r.get<uint64_t>();
            KB2_REQUIRE((int64_t)pd == o.d && pm >= 1 && pm <= (uint64_t)o.d && o.d % pm == 0, KB2_INVALID_BINARY_SET, "faiss stream: bad PQ geometry");
            KB2_REQUIRE(pb == 8 && cs == pm, KB2_NOT_IMPLEMENTED, "faiss stream: only nbits=8 product quantizers are supported");
            o.M = (int)pm;
            o.nbits = (int)pb;
            read_vec(o.pq_centroids);
            KB2_REQUIRE(o.pq_centroids.size() == (size_t)256 * o.d, KB2_INVALID_BINARY_SET, "faiss stream: PQ codebook size");
            read_invlists(o, cs);
            return o;
        }
        if (h == fourcc("IHNf") || h == fourcc("IHN9")) {
            Header hd = read_header();
            o.kind = "HNSW";
            o.d = hd.d;
            o.ntotal = hd.ntotal;
            o.metric = hd.metric_faiss == 0 ? KB2_METRIC_IP : KB2_METRIC_L2;
            o.cosine = hd.cosine || h == fourcc("IHN9");
            read_vec(o.assign_probas);
            read_vec(o.cum);
            read_vec(o.levels);
            read_vec(o.offsets);
            read_vec(o.neighbors);
            o.entry_point = r.get<int32_t>();
            o.max_level = r.get<int32_t>();
            o.efConstruction = r.get<int32_t>();
            o.efSearch = r.get<int32_t>();
            o.upper_beam = r.get<int32_t>();
            const uint32_t sh = r.get<uint32_t>();
            KB2_REQUIRE(is_flat(sh), KB2_NOT_IMPLEMENTED, "faiss stream: HNSW storage must be a flat fp32 index");
            Header sd;
            read_flat_body(sh, sd, o.xb, &o.xb_norms);
            KB2_REQUIRE(sd.d == o.d && sd.ntotal == o.ntotal, KB2_INVALID_BINARY_SET, "faiss stream: HNSW storage geometry mismatch");
            KB2_REQUIRE((int64_t)o.levels.size() == o.ntotal && (int64_t)o.offsets.size() == o.ntotal + 1, KB2_INVALID_BINARY_SET,
                        "faiss stream: HNSW graph arrays do not match ntotal");
            return o;
        }
        char name[5] = {(char)(h & 255), (char)((h >> 8) & 255), (char)((h >> 16) & 255), (char)((h >> 24) & 255), 0};
        throw Error(KB2_NOT_IMPLEMENTED, std::string("faiss stream: unsupported index fourcc '") + name + "'");
    }
};

struct FaissWriter {
    BlobWriter w;
    template <typename T>
    void
    write_vec(const T* p, size_t n) {
        w.put<uint64_t>((uint64_t)n);
        w.put_bytes(p, n * sizeof(T));
    }
    void
    write_header(int d, int64_t ntotal, bool cosine, int metric) {
        w.put<int32_t>(d);
        w.put<int64_t>(ntotal);
        w.put<uint8_t>(cosine ? 1 : 0);
        w.put<uint8_t>(0); w.put<uint8_t>(0); w.put<uint8_t>(0);
        w.put<uint32_t>(0);
        w.put<int64_t>(0);
        w.put<uint8_t>(1);   // is_trained
        w.put<int32_t>(metric == KB2_METRIC_IP ? 0 : 1);
    }
    void
    write_flat(int d, int64_t n, int metric, const float* xb) {
        w.put<uint32_t>(metric == KB2_METRIC_IP ? fourcc("IxFI") : fourcc("IxF2"));
        write_header(d, n, false, metric);
        write_vec(xb, (size_t)n * d);
    }
    void
    write_ivf_header(const FaissIndexData& o) {
        write_header(o.d, o.ntotal, false, o.metric);
        w.put<uint64_t>((uint64_t)o.nlist);
        w.put<uint64_t>((uint64_t)o.nprobe);
        write_flat(o.d, o.nlist, o.metric, o.centroids.data());
        w.put<uint8_t>(0);   // DirectMap::NoMap
        w.put<uint64_t>(0);
    }
    void
    write_invlists(const FaissIndexData& o) {
        w.put<uint32_t>(fourcc("ilar"));
        w.put<uint64_t>((uint64_t)o.nlist);
        w.put<uint64_t>(o.code_size);
        uint64_t non0 = 0;
        for (auto& l : o.list_ids) non0 += !l.empty();
        if (non0 > (uint64_t)o.nlist / 2) {
            w.put<uint32_t>(fourcc("full"));
            std::vector<uint64_t> sizes(o.nlist);
            for (int64_t l = 0; l < o.nlist; l++) sizes[l] = o.list_ids[l].size();
            write_vec(sizes.data(), sizes.size());
        } else {
            w.put<uint32_t>(fourcc("sprs"));
            std::vector<uint64_t> sizes;
            for (int64_t l = 0; l < o.nlist; l++)
                if (!o.list_ids[l].empty()) { sizes.push_back((uint64_t)l); sizes.push_back(o.list_ids[l].size()); }
            write_vec(sizes.data(), sizes.size());
        }
        for (int64_t l = 0; l < o.nlist; l++) {
            if (o.list_ids[l].empty()) continue;
            w.put_bytes(o.list_codes[l].data(), o.list_codes[l].size());
            w.put_bytes(o.list_ids[l].data(), o.list_ids[l].size() * 8);
        }
    }
    void
    write_index(const FaissIndexData& o) {
        if (o.kind == "FLAT") {
            write_flat(o.d, o.ntotal, o.metric, o.xb.data());
        } else if (o.kind == "IVF_FLAT") {
            w.put<uint32_t>(fourcc("IwFl"));
            write_ivf_header(o);
            write_invlists(o);
        } else if (o.kind == "IVF_PQ") {
            if (o.has_refine) {
                w.put<uint32_t>(fourcc("IxRF"));
                write_header(o.d, o.ntotal, false, o.metric);
            }
            w.put<uint32_t>(fourcc("IwPQ"));
            write_ivf_header(o);
            w.put<uint8_t>(1);   // by_residual
            w.put<uint64_t>((uint64_t)o.M);
            w.put<uint64_t>((uint64_t)o.d);
            w.put<uint64_t>((uint64_t)o.M);
            w.put<uint64_t>(8);
            write_vec(o.pq_centroids.data(), o.pq_centroids.size());
            write_invlists(o);
            if (o.has_refine) {
                write_flat(o.d, o.ntotal, o.metric, o.refine_xb.data());
                w.put<float>(o.k_factor);
            }
        } else if (o.kind == "HNSW") {
            w.put<uint32_t>(fourcc("IHNf"));
            write_header(o.d, o.ntotal, false, o.metric);
            write_vec(o.assign_probas.data(), o.assign_probas.size());
            write_vec(o.cum.data(), o.cum.size());
            write_vec(o.levels.data(), o.levels.size());
            write_vec(o.offsets.data(), o.offsets.size());
            write_vec(o.neighbors.data(), o.neighbors.size());
            w.put<int32_t>(o.entry_point);
            w.put<int32_t>(o.max_level);
            w.put<int32_t>(o.efConstruction);
            w.put<int32_t>(o.efSearch);
            w.put<int32_t>(o.upper_beam);
            write_flat(o.d, o.ntotal, o.metric, o.xb.data());
        } else {
            throw Error(KB2_NOT_IMPLEMENTED, "faiss stream: cannot write this index kind");
        }
    }
};

// one-line JSON description of a stream (CPU only; used by the ABI self-test and by hosts that route a BinarySet)
inline std::string
faiss_describe(const FaissIndexData& o) {
    std::string s = "{\"type\": \"" + o.kind + "\", \"dim\": " + std::to_string(o.d) + ", \"rows\": " + std::to_string(o.ntotal) +
                    ", \"metric_type\": \"" + (o.cosine ? "COSINE" : (o.metric == KB2_METRIC_IP ? "IP" : "L2")) + "\"";
    if (o.kind == "IVF_FLAT" || o.kind == "IVF_PQ") s += ", \"nlist\": " + std::to_string(o.nlist);
    if (o.kind == "IVF_PQ") s += ", \"m\": " + std::to_string(o.M) + ", \"nbits\": 8, \"refine\": " + (o.has_refine ? "true" : "false");
    if (o.kind == "HNSW") {
        const int M = o.cum.size() >= 2 ? o.cum[1] / 2 : 0;
        s += ", \"M\": " + std::to_string(M) + ", \"efConstruction\": " + std::to_string(o.efConstruction) +
             ", \"max_level\": " + std::to_string(o.max_level);
    }
    return s + "}";
}

}  // namespace kb2
