// kb2_blob.h — little-endian byte-stream helpers for the "KB2I" serialisation container.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "kb2_common.cuh"

namespace kb2 {

struct BlobWriter {
    std::vector<uint8_t>& b;
    template <typename T> void put(const T& v) { const uint8_t* p = (const uint8_t*)&v; b.insert(b.end(), p, p + sizeof(T)); }
    void put_bytes(const void* p, size_t n) { const uint8_t* c = (const uint8_t*)p; b.insert(b.end(), c, c + n); }
    void put_str(const std::string& s) { put<uint32_t>((uint32_t)s.size()); put_bytes(s.data(), s.size()); }
};
struct BlobReader {
    const uint8_t* p;
    size_t n, o = 0;
    template <typename T> T get() {
        KB2_REQUIRE(o + sizeof(T) <= n, KB2_INVALID_BINARY_SET, "truncated blob");
        T v; memcpy(&v, p + o, sizeof(T)); o += sizeof(T); return v;
    }
    const uint8_t* get_bytes(size_t cnt) {
        KB2_REQUIRE(o + cnt <= n, KB2_INVALID_BINARY_SET, "truncated blob");
        const uint8_t* r = p + o; o += cnt; return r;
    }
    std::string get_str() { uint32_t l = get<uint32_t>(); const uint8_t* s = get_bytes(l); return std::string((const char*)s, l); }
};


}  // namespace kb2
