// fast5_loader.cpp -- libpb_fast5.so: FAST5 (HDF5) ingest for the signal path, host only.
// ABI and reference citations: include/poreplex_b200_fast5.h.
//
// A deliberately small HDF5 reader (the same subset as poreplex_b200/hdf5_min.py): the file is
// mmap'ed read-only, every access is bounds-checked (a truncated or foreign file raises a format
// error instead of reading past the map), groups are searched through their v1 B-tree keys
// (O(log n) per lookup in a 4000-read file), and a batch of reads is decoded by a pool of
// threads straight into the caller's packed int16 buffer -- typically pinned memory that
// pb2_analyze_host then copies to the GPU.
#include "../../include/poreplex_b200_fast5.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "inflate_fast.h"

namespace {

constexpr uint64_t UNDEF = ~0ull;
// sanity limits: sizes read from a (possibly damaged) file never drive an allocation beyond these
constexpr uint64_t MAX_SIGNAL = 1ull << 31;        // samples per read
constexpr uint64_t MAX_CHUNK_BYTES = 1ull << 28;   // decoded bytes per chunk

struct FormatError : std::runtime_error { using std::runtime_error::runtime_error; };
struct NotFound : std::runtime_error { using std::runtime_error::runtime_error; };

[[noreturn]] void bad(const char *fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw FormatError(buf);
}

thread_local std::string g_error;

// ---- optional libzstd (VBZ) ------------------------------------------------------------------
typedef size_t (*zstd_decompress_fn)(void *, size_t, const void *, size_t);
typedef unsigned (*zstd_iserror_fn)(size_t);
typedef unsigned long long (*zstd_framesize_fn)(const void *, size_t);
struct Zstd {
    zstd_decompress_fn decompress = nullptr;
    zstd_iserror_fn is_error = nullptr;
    zstd_framesize_fn frame_size = nullptr;
    Zstd() {
        void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
        is_error = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
        frame_size = (zstd_framesize_fn)dlsym(h, "ZSTD_getFrameContentSize");
    }
    bool ok() const { return decompress && is_error && frame_size; }
};
const Zstd &zstd()
{
    static Zstd z;
    return z;
}

// ---- bounds-checked view of the mapped file -------------------------------------------------------
struct Map {
    const uint8_t *p = nullptr;
    uint64_t n = 0;
    const uint8_t *at(uint64_t off, uint64_t len) const {
        if (off > n || len > n - off) bad("access beyond the end of the file (offset %llu + %llu)",
                                           (unsigned long long)off, (unsigned long long)len);
        return p + off;
    }
    uint8_t u8(uint64_t o) const { return *at(o, 1); }
    uint16_t u16(uint64_t o) const { uint16_t v; memcpy(&v, at(o, 2), 2); return v; }
    uint32_t u32(uint64_t o) const { uint32_t v; memcpy(&v, at(o, 4), 4); return v; }
    uint64_t u64(uint64_t o) const { uint64_t v; memcpy(&v, at(o, 8), 8); return v; }
    bool sig(uint64_t o, const char *s) const { return memcmp(at(o, 4), s, 4) == 0; }
    std::string cstr(uint64_t o) const {
        if (o >= n) bad("string beyond the end of the file");
        const void *e = memchr(p + o, 0, n - o);
        if (!e) bad("unterminated string");
        return std::string((const char *)p + o, (const char *)e);
    }
};

inline uint64_t pad8(uint64_t x) { return (x + 7) & ~7ull; }

struct Msg { uint16_t type; uint64_t off; uint32_t size; };     // body at file offset `off`

struct Datatype {
    int cls = -1;           // 0 fixed point, 1 float, 3 string, 9 vlen
    uint32_t size = 0;
    bool is_signed = false, big_endian = false, vlen_string = false;
    uint64_t consumed = 0;
};

struct Filter { uint16_t id; std::vector<uint32_t> cd; };

struct Dataset {
    std::vector<uint64_t> dims;
    Datatype type;
    int layout = -1;                 // 0 compact, 1 contiguous, 2 chunked
    uint64_t addr = UNDEF, size = 0; // contiguous: data address; compact: offset of the bytes
    uint64_t btree = UNDEF;
    std::vector<uint32_t> chunk;     // chunk dims (without the element size)
    std::vector<Filter> filters;
    uint64_t count() const { uint64_t c = 1; for (uint64_t d : dims) c *= d; return c; }
};

struct AttrValue {
    bool is_string = false;
    std::string s;
    double f = 0;
    int64_t i = 0;
};

}  // namespace

struct pb2f_file {
    std::string path;
    int fd = -1;
    Map m;
    uint64_t root = 0;                 // object header of the root group
    bool multiread = true;
    bool names_loaded = false;
    std::vector<std::string> names;    // read names (lazy)
    std::mutex mu;                     // guards the lazy name list

    ~pb2f_file() {
        if (m.p) munmap((void *)m.p, m.n);
        if (fd >= 0) close(fd);
    }

    // ---- object headers (version 1, with continuation blocks) --------------------------------
    std::vector<Msg> messages(uint64_t addr) const {
        if (m.u8(addr) != 1) {
            if (m.n >= addr + 4 && m.sig(addr, "OHDR")) bad("version-2 object headers are not supported");
            bad("bad object header at %llu", (unsigned long long)addr);
        }
        const unsigned nmsgs = m.u16(addr + 2);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, m.u32(addr + 8)}};
        std::vector<Msg> out;
        for (size_t b = 0; b < blocks.size() && out.size() < nmsgs; b++) {
            uint64_t p = blocks[b].first;
            const uint64_t end = p + blocks[b].second;
            while (p + 8 <= end && out.size() < nmsgs) {
                const uint16_t type = m.u16(p), size = m.u16(p + 2);
                m.at(p + 8, size);
                if (type == 0x0010) blocks.push_back({m.u64(p + 8), m.u64(p + 16)});
                out.push_back({type, p + 8, size});
                p += 8 + (uint64_t)size;
            }
            if (blocks.size() > 4096) bad("object header continuation loop");
        }
        return out;
    }

    // ---- groups -------------------------------------------------------------------------------
    bool group_tables(uint64_t ohdr, uint64_t &btree, uint64_t &heap_data) const {
        for (const Msg &g : messages(ohdr))
            if (g.type == 0x0011) {
                btree = m.u64(g.off);
                const uint64_t heap = m.u64(g.off + 8);
                if (!m.sig(heap, "HEAP")) bad("bad local heap signature");
                heap_data = m.u64(heap + 24);
                return true;
            }
        return false;
    }

    // child object header of `name` in the group at `ohdr`; UNDEF when absent
    uint64_t child(uint64_t ohdr, const std::string &name) const {
        uint64_t node, heap;
        if (!group_tables(ohdr, node, heap)) throw NotFound("not a group");
        for (int depth = 0; depth < 64; depth++) {
            if (m.sig(node, "TREE")) {
                if (m.u8(node + 4) != 0) bad("group B-tree node of the wrong type");
                const unsigned used = m.u16(node + 6);
                uint64_t p = node + 24;                  // key0, child0, key1, ...
                uint64_t next = UNDEF;
                for (unsigned i = 0; i < used; i++) {
                    const std::string hi = m.cstr(heap + m.u64(p + 16 * (uint64_t)i + 16));
                    if (name <= hi) { next = m.u64(p + 16 * (uint64_t)i + 8); break; }
                }
                if (next == UNDEF) return UNDEF;
                node = next;
            } else if (m.sig(node, "SNOD")) {
                const unsigned nsym = m.u16(node + 6);
                for (unsigned i = 0; i < nsym; i++) {
                    const uint64_t e = node + 8 + 40ull * i;
                    if (m.cstr(heap + m.u64(e)) == name) return m.u64(e + 8);
                }
                return UNDEF;
            } else {
                bad("bad group node signature");
            }
        }
        bad("group B-tree too deep");
    }

    void list(uint64_t ohdr, std::vector<std::string> &out) const {
        uint64_t root_node, heap;
        if (!group_tables(ohdr, root_node, heap)) throw NotFound("not a group");
        // depth-first, children in order: names come out sorted as the file stores them
        struct Walk {
            const pb2f_file &f; uint64_t heap; std::vector<std::string> &out;
            // A crafted node that names itself (or an ancestor) as its child must not hang the
            // loader: node levels have to fall by one per step (level 0 points at symbol nodes),
            // which rules out cycles, and the whole walk has a node budget.
            uint64_t visited = 0;
            void go(uint64_t node, int depth, int want_level = -1) {
                if (depth > 64) bad("group B-tree too deep");
                if (++visited > (1u << 22)) bad("group B-tree has too many nodes");
                if (f.m.sig(node, "TREE")) {
                    const int level = f.m.u8(node + 5);
                    if (want_level == -2 || (want_level >= 0 && level != want_level))
                        bad("group B-tree levels are inconsistent");
                    const unsigned used = f.m.u16(node + 6);
                    for (unsigned i = 0; i < used; i++)
                        go(f.m.u64(node + 24 + 16ull * i + 8), depth + 1, level > 0 ? level - 1 : -2);
                } else if (f.m.sig(node, "SNOD")) {
                    if (want_level >= 0) bad("group B-tree levels are inconsistent");
                    const unsigned nsym = f.m.u16(node + 6);
                    for (unsigned i = 0; i < nsym; i++)
                        out.push_back(f.m.cstr(heap + f.m.u64(node + 8 + 40ull * i)));
                } else {
                    bad("bad group node signature");
                }
            }
        } w{*this, heap, out, 0};
        w.go(root_node, 0);
    }

    uint64_t lookup(const std::string &path) const {      // absolute path from the root
        uint64_t cur = root;
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') i++;
            size_t j = path.find('/', i);
            if (j == std::string::npos) j = path.size();
            if (j > i) {
                cur = child(cur, path.substr(i, j - i));
                if (cur == UNDEF) throw NotFound("no such node: " + path);
            }
            i = j;
        }
        return cur;
    }

    // ---- datatypes / dataspaces / attributes ---------------------------------------------------
    Datatype datatype(uint64_t p) const {
        Datatype t;
        const uint8_t b0 = m.u8(p);
        t.cls = b0 & 0x0F;
        const uint32_t bits = m.u8(p + 1) | (m.u8(p + 2) << 8) | (m.u8(p + 3) << 16);
        t.size = m.u32(p + 4);
        if (t.cls == 0) { t.big_endian = bits & 1; t.is_signed = bits & 8; t.consumed = 12; }
        else if (t.cls == 1) { t.big_endian = bits & 1; t.consumed = 20; }
        else if (t.cls == 3) { t.consumed = 8; }
        else if (t.cls == 9) {
            t.vlen_string = (bits & 0x0F) == 1;
            t.consumed = 8 + datatype(p + 8).consumed;
        } else if (t.cls == 8) {                    // enumeration: read as its base integer
            const Datatype base = datatype(p + 8);
            if (base.cls != 0) bad("unsupported enumeration base type");
            t.cls = 0; t.is_signed = base.is_signed; t.big_endian = base.big_endian;
            t.consumed = 0;
        } else {
            t.consumed = 0;                         // compound etc.: usable only as "skip"
        }
        return t;
    }

    std::vector<uint64_t> dataspace(uint64_t p) const {
        const uint8_t version = m.u8(p), rank = m.u8(p + 1);
        uint64_t q;
        if (version == 1) q = p + 8; else if (version == 2) q = p + 4;
        else bad("unsupported dataspace version %d", version);
        std::vector<uint64_t> dims(rank);
        for (unsigned i = 0; i < rank; i++) dims[i] = m.u64(q + 8ull * i);
        return dims;
    }

    std::string global_heap_object(uint64_t coll, uint32_t index) const {
        if (!m.sig(coll, "GCOL")) bad("bad global heap signature");
        const uint64_t csize = m.u64(coll + 8);
        uint64_t p = coll + 16;
        const uint64_t end = coll + csize;
        while (p + 16 <= end) {
            const uint16_t idx = m.u16(p);
            const uint64_t osize = m.u64(p + 8);
            if (idx == 0) break;
            if (idx == index) return std::string((const char *)m.at(p + 16, osize), osize);
            p += 16 + pad8(osize);
        }
        bad("global heap object not found");
    }

    bool attribute(uint64_t ohdr, const char *want, AttrValue &out) const {
        for (const Msg &g : messages(ohdr)) {
            if (g.type != 0x000C) continue;
            const uint8_t version = m.u8(g.off);
            const uint16_t nsize = m.u16(g.off + 2), tsize = m.u16(g.off + 4), ssize = m.u16(g.off + 6);
            uint64_t p;
            std::string name;
            uint64_t tp, sp, dp;
            if (version == 1) {
                p = g.off + 8; name = m.cstr(p);
                tp = p + pad8(nsize); sp = tp + pad8(tsize); dp = sp + pad8(ssize);
            } else if (version == 2 || version == 3) {
                p = g.off + (version == 2 ? 8 : 9); name = m.cstr(p);
                tp = p + nsize; sp = tp + tsize; dp = sp + ssize;
            } else {
                bad("unsupported attribute version %d", version);
            }
            if (name != want) continue;
            const Datatype t = datatype(tp);
            if (t.cls == 9 && t.vlen_string) {
                out.is_string = true;
                out.s = global_heap_object(m.u64(dp + 4), m.u32(dp + 12));
            } else if (t.cls == 3) {
                out.is_string = true;
                const char *c = (const char *)m.at(dp, t.size);
                out.s.assign(c, strnlen(c, t.size));
            } else if (t.cls == 0 && !t.big_endian && t.size <= 8) {
                uint64_t v = 0;
                memcpy(&v, m.at(dp, t.size), t.size);
                if (t.is_signed && t.size < 8 && (v >> (8 * t.size - 1)) & 1) v |= ~0ull << (8 * t.size);
                out.i = (int64_t)v;
                out.f = t.is_signed ? (double)(int64_t)v : (double)v;
            } else if (t.cls == 1 && !t.big_endian && (t.size == 4 || t.size == 8)) {
                if (t.size == 4) { float v; memcpy(&v, m.at(dp, 4), 4); out.f = v; }
                else memcpy(&out.f, m.at(dp, 8), 8);
                out.i = (int64_t)out.f;
            } else {
                bad("unsupported datatype of attribute %s", want);
            }
            // a trailing NUL inside fixed strings is dropped above; vlen strings may carry one
            while (out.is_string && !out.s.empty() && out.s.back() == '\0') out.s.pop_back();
            return true;
        }
        return false;
    }

    AttrValue need_attr(uint64_t ohdr, const char *name) const {
        AttrValue v;
        if (!attribute(ohdr, name, v)) throw NotFound(std::string("missing attribute ") + name);
        return v;
    }

    // ---- datasets ---------------------------------------------------------------------------------
    Dataset dataset(uint64_t ohdr) const {
        Dataset d;
        bool have_space = false, have_type = false;
        for (const Msg &g : messages(ohdr)) {
            if (g.type == 0x0001) { d.dims = dataspace(g.off); have_space = true; }
            else if (g.type == 0x0003) { d.type = datatype(g.off); have_type = true; }
            else if (g.type == 0x0008) {
                const uint8_t version = m.u8(g.off);
                if (version == 3) {
                    d.layout = m.u8(g.off + 1);
                    if (d.layout == 1) { d.addr = m.u64(g.off + 2); d.size = m.u64(g.off + 10); }
                    else if (d.layout == 0) { d.size = m.u16(g.off + 2); d.addr = g.off + 4; }
                    else if (d.layout == 2) {
                        const unsigned ndim = m.u8(g.off + 2);
                        if (ndim < 2) bad("chunked layout without dimensions");
                        d.btree = m.u64(g.off + 3);
                        for (unsigned i = 0; i + 1 < ndim; i++) d.chunk.push_back(m.u32(g.off + 11 + 4ull * i));
                    } else bad("unsupported layout class %d", d.layout);
                } else if (version == 1 || version == 2) {
                    const unsigned rank = m.u8(g.off + 1);
                    d.layout = m.u8(g.off + 2);
                    if (d.layout == 1) d.addr = m.u64(g.off + 8);
                    else if (d.layout == 2) {
                        d.btree = m.u64(g.off + 8);
                        for (unsigned i = 0; i + 1 < rank; i++) d.chunk.push_back(m.u32(g.off + 16 + 4ull * i));
                    } else bad("unsupported v1/v2 layout class %d", d.layout);
                } else bad("unsupported layout version %d", version);
            } else if (g.type == 0x000B) {
                const uint8_t version = m.u8(g.off), nf = m.u8(g.off + 1);
                uint64_t p = g.off + (version == 1 ? 8 : 2);
                if (version != 1 && version != 2) bad("unsupported filter pipeline version %d", version);
                for (unsigned k = 0; k < nf; k++) {
                    Filter f;
                    f.id = m.u16(p); p += 2;
                    uint16_t nlen = 0;
                    if (version == 1 || f.id >= 256) { nlen = m.u16(p); p += 2; }
                    p += 2;                                         // flags
                    const uint16_t ncd = m.u16(p); p += 2;
                    p += version == 1 ? pad8(nlen) : nlen;
                    for (unsigned c = 0; c < ncd; c++) { f.cd.push_back(m.u32(p)); p += 4; }
                    if (version == 1 && (ncd & 1)) p += 4;
                    d.filters.push_back(f);
                }
            }
        }
        if (!have_space || !have_type || d.layout < 0) throw NotFound("not a dataset");
        return d;
    }
};

namespace {

// ---- filters -------------------------------------------------------------------------------------
void unshuffle(std::vector<uint8_t> &buf, std::vector<uint8_t> &tmp, uint32_t es)
{
    if (es <= 1) return;
    const size_t n = buf.size() / es;
    tmp.resize(buf.size());
    for (uint32_t b = 0; b < es; b++) {
        const uint8_t *src = buf.data() + (size_t)b * n;
        for (size_t i = 0; i < n; i++) tmp[i * es + b] = src[i];
    }
    memcpy(tmp.data() + n * es, buf.data() + n * es, buf.size() - n * es);
    buf.swap(tmp);
}

// per-thread working memory: two byte buffers and the Huffman tables of the inflater (35 KB)
struct Scratch {
    std::vector<uint8_t> a, b;
    std::unique_ptr<pbinf::Tables> tables;
    Scratch() : tables(new pbinf::Tables) {}
};

void inflate_chunk(std::vector<uint8_t> &buf, Scratch &s, size_t expected)
{
    std::vector<uint8_t> &tmp = s.b;
    size_t cap = expected ? expected : buf.size() * 4 + 64;
    for (int attempt = 0; attempt < 8; attempt++) {
        if (cap > 2 * MAX_CHUNK_BYTES) bad("deflate: chunk larger than any plausible size");
        tmp.resize(cap);
        const int64_t got = pbinf::zlib_decompress(buf.data(), buf.size(), tmp.data(), cap, *s.tables);
        if (got >= 0) { tmp.resize((size_t)got); buf.swap(tmp); return; }
        if (got != pbinf::ERR_OVERFLOW) bad("deflate: corrupt chunk (%lld)", (long long)got);
        cap *= 2;
    }
    bad("deflate: chunk larger than expected");
}

// streamvbyte: `count` little-endian codes, keys first (KEY_BITS = 1: svb16, 1..2 bytes per value;
// KEY_BITS = 2: classic, 1..4 bytes), data after
template <int KEY_BITS>
void svb_decode_t(const uint8_t *src, size_t len, size_t count, uint32_t *out)
{
    constexpr unsigned PER = 8 / KEY_BITS, SHIFT = KEY_BITS == 1 ? 3 : 2, KMASK = (1u << KEY_BITS) - 1;
    const size_t nkeys = (count + PER - 1) / PER;
    if (nkeys > len) bad("VBZ: truncated key block");
    const uint8_t *data = src + nkeys, *end = src + len;
    size_t i = 0;
    // whole key bytes while at least 4 * PER data bytes remain: no per-value bounds check
    while (i + PER <= count && (size_t)(end - data) >= 4 * PER) {
        unsigned key = src[i >> SHIFT];
        for (unsigned k = 0; k < PER; k++, key >>= KEY_BITS) {
            const unsigned nb = (key & KMASK) + 1;
            uint32_t v;
            memcpy(&v, data, 4);
            out[i + k] = nb == 4 ? v : v & ((1u << (8 * nb)) - 1);
            data += nb;
        }
        i += PER;
    }
    for (; i < count; i++) {
        const unsigned nb = ((src[i >> SHIFT] >> ((i & (PER - 1)) * KEY_BITS)) & KMASK) + 1;
        if (data + nb > end) bad("VBZ: truncated data block");
        uint32_t v = 0;
        for (unsigned b2 = 0; b2 < nb; b2++) v |= (uint32_t)data[b2] << (8 * b2);
        data += nb;
        out[i] = v;
    }
}

void svb_decode(const uint8_t *src, size_t len, size_t count, int key_bits, std::vector<uint32_t> &out)
{
    out.resize(count);
    if (key_bits == 1) svb_decode_t<1>(src, len, count, out.data());
    else svb_decode_t<2>(src, len, count, out.data());
}

// ONT VBZ (filter 32020): uint32 uncompressed size, optional zstd frame, streamvbyte of the
// (delta, zigzag) coded integers; version 1 codes 2-byte integers with 1-bit keys.  Layout per
// ONT's published vbz_compression; NOT checked against a file written by ONT's plugin.
void vbz_decode(std::vector<uint8_t> &buf, std::vector<uint8_t> &tmp, const std::vector<uint32_t> &cd,
                size_t expected)
{
    const uint32_t version = cd.size() > 0 ? cd[0] : 0, isize = cd.size() > 1 ? cd[1] : 0;
    const uint32_t zigzag = cd.size() > 2 ? cd[2] : 0, level = cd.size() > 3 ? cd[3] : 0;
    if (version > 1) bad("unsupported VBZ version %u", version);
    if (buf.size() < 4) bad("VBZ: short chunk");
    uint32_t size;
    memcpy(&size, buf.data(), 4);
    if (size > (expected ? expected : MAX_CHUNK_BYTES)) bad("VBZ: chunk decodes to more than its size");
    const uint8_t *body = buf.data() + 4;
    size_t blen = buf.size() - 4;
    if (level != 0) {
        const Zstd &z = zstd();
        if (!z.ok()) bad("VBZ-compressed dataset: libzstd is not available");
        unsigned long long fs = z.frame_size(body, blen);
        if (fs >= (1ull << 62)) fs = (unsigned long long)size * 2 + 64;
        if (fs > 5ull * size + 1024) bad("VBZ: implausible zstd frame size");
        tmp.resize((size_t)fs ? (size_t)fs : 1);
        const size_t got = z.decompress(tmp.data(), tmp.size(), body, blen);
        if (z.is_error(got)) bad("VBZ: zstd decompression failed");
        tmp.resize(got);
        body = tmp.data();
        blen = got;
    }
    std::vector<uint8_t> out;
    if (isize == 0 || isize == 1) {
        if (blen < size) bad("VBZ: short payload");
        out.assign(body, body + size);
    } else {
        if (isize != 2 && isize != 4) bad("VBZ: unsupported integer size %u", isize);
        const size_t count = size / isize;
        const bool svb16 = version == 1 && isize == 2;
        std::vector<uint32_t> vals;
        svb_decode(body, blen, count, svb16 ? 1 : 2, vals);
        out.resize(count * isize);
        if (isize == 2) {
            uint16_t prev16 = 0;
            uint32_t prev32 = 0;
            for (size_t i = 0; i < count; i++) {
                uint16_t v;
                if (!zigzag) v = (uint16_t)vals[i];
                else if (svb16) {
                    const uint16_t c = (uint16_t)vals[i];
                    prev16 = (uint16_t)(prev16 + (uint16_t)((c >> 1) ^ (uint16_t)(0 - (c & 1))));
                    v = prev16;
                } else {
                    const uint32_t c = vals[i];
                    prev32 += (c >> 1) ^ (0u - (c & 1u));
                    v = (uint16_t)prev32;
                }
                memcpy(out.data() + 2 * i, &v, 2);
            }
        } else {
            uint32_t prev = 0;
            for (size_t i = 0; i < count; i++) {
                uint32_t v = vals[i];
                if (zigzag) { prev += (v >> 1) ^ (0u - (v & 1u)); v = prev; }
                memcpy(out.data() + 4 * i, &v, 4);
            }
        }
    }
    buf.swap(out);
}

void defilter(std::vector<uint8_t> &buf, Scratch &s, const std::vector<Filter> &filters,
              uint32_t mask, size_t expected)
{
    for (size_t k = filters.size(); k-- > 0;) {
        if (mask & (1u << k)) continue;
        const Filter &f = filters[k];
        switch (f.id) {
        case 1: inflate_chunk(buf, s, expected); break;
        case 2: unshuffle(buf, s.b, f.cd.empty() ? 1 : f.cd[0]); break;
        case 3: if (buf.size() < 4) bad("fletcher32: short chunk"); buf.resize(buf.size() - 4); break;
        case 32020: vbz_decode(buf, s.b, f.cd, expected); break;
        default: bad("unsupported filter %u", (unsigned)f.id);
        }
    }
}

// whole 1-D int16 dataset into dst[0 .. count)
void read_int16(const pb2f_file &f, const Dataset &d, int16_t *dst, Scratch &s)
{
    if (d.type.cls != 0 || d.type.size != 2 || d.type.big_endian)
        bad("Signal is not a little-endian 16-bit integer dataset");
    if (d.dims.size() != 1) bad("Signal is not one-dimensional");
    const uint64_t n = d.dims[0];
    if (d.layout == 0 || d.layout == 1) {
        if (d.layout == 1 && d.addr == UNDEF) { memset(dst, 0, n * 2); return; }
        memcpy(dst, f.m.at(d.addr, n * 2), n * 2);
        return;
    }
    if (d.chunk.size() != 1 || d.chunk[0] == 0) bad("unexpected chunk shape");
    const uint64_t clen = d.chunk[0];
    if (clen * 2 > MAX_CHUNK_BYTES) bad("implausible chunk length %llu", (unsigned long long)clen);
    memset(dst, 0, n * 2);                         // chunks never written read as the fill value
    if (d.btree == UNDEF) return;
    struct Walk {
        const pb2f_file &f; const Dataset &d; int16_t *dst; Scratch &s; uint64_t n, clen;
        uint64_t visited = 0;
        void go(uint64_t node, int depth, int want_level = -1) {
            if (depth > 32) bad("chunk B-tree too deep");
            if (++visited > (1u << 22)) bad("chunk B-tree has too many nodes");
            if (!f.m.sig(node, "TREE") || f.m.u8(node + 4) != 1) bad("bad chunk B-tree node");
            const unsigned level = f.m.u8(node + 5), used = f.m.u16(node + 6);
            if (want_level >= 0 && (int)level != want_level) bad("chunk B-tree levels are inconsistent");
            const uint64_t klen = 8 + 8 * 2;               // size, mask, offset[rank + 1], rank 1
            uint64_t p = node + 24;
            for (unsigned i = 0; i < used; i++, p += klen + 8) {
                const uint32_t size = f.m.u32(p), mask = f.m.u32(p + 4);
                const uint64_t off = f.m.u64(p + 8), child = f.m.u64(p + klen);
                if (level) { go(child, depth + 1, (int)level - 1); continue; }
                if (off >= n) continue;
                const uint64_t take = off + clen <= n ? clen : n - off;
                const uint8_t *src = f.m.at(child, size);
                if (d.filters.empty()) {
                    if (size < take * 2) bad("short chunk");
                    memcpy(dst + off, src, take * 2);
                } else {
                    s.a.assign(src, src + size);
                    defilter(s.a, s, d.filters, mask, clen * 2);
                    if (s.a.size() < take * 2) bad("short chunk after filtering");
                    memcpy(dst + off, s.a.data(), take * 2);
                }
            }
        }
    } w{f, d, dst, s, n, clen, 0};
    w.go(d.btree, 0);
}

// node paths of one read (fast5_file.py:69-82)
struct ReadNodes { uint64_t raw, channel, tracking; };

ReadNodes read_nodes(pb2f_file &f, const char *read_id)
{
    ReadNodes r;
    if (f.multiread) {
        if (!read_id) {
            if (pb2f_num_reads(&f) < 1) throw NotFound("no reads in the file");
            read_id = f.names[0].c_str();
        }
        const uint64_t g = f.child(f.root, std::string("read_") + read_id);
        if (g == UNDEF) throw NotFound(std::string("no read ") + read_id);
        r.raw = f.child(g, "Raw");
        r.channel = f.child(g, "channel_id");
        r.tracking = f.child(g, "tracking_id");
    } else {
        const uint64_t reads = f.lookup("Raw/Reads");
        std::vector<std::string> names;
        f.list(reads, names);
        if (names.empty()) throw NotFound("no reads in the file");
        r.raw = f.child(reads, names[0]);
        r.channel = f.lookup("UniqueGlobalKey/channel_id");
        r.tracking = f.lookup("UniqueGlobalKey/tracking_id");
    }
    if (r.raw == UNDEF || r.channel == UNDEF || r.tracking == UNDEF)
        throw NotFound("read without Raw / channel_id / tracking_id");
    return r;
}

void copy_str(char *dst, size_t cap, const std::string &s)
{
    const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(dst, s.data(), n);
    dst[n] = 0;
}

void load_meta(pb2f_file &f, const char *read_id, pb2f_read_meta &out, Dataset *sig_out)
{
    const ReadNodes r = read_nodes(f, read_id);
    memset(&out, 0, sizeof out);
    out.duration = f.need_attr(r.raw, "duration").i;
    out.start_time = f.need_attr(r.raw, "start_time").i;
    const std::string rid = f.need_attr(r.raw, "read_id").s;
    if (read_id && rid != read_id)                  // fast5_file.py:105-108
        bad("Unexpected read %s found", rid.c_str());
    copy_str(out.read_id, sizeof out.read_id, rid);
    copy_str(out.channel_number, sizeof out.channel_number, f.need_attr(r.channel, "channel_number").s);
    out.digitisation = f.need_attr(r.channel, "digitisation").f;
    out.offset = f.need_attr(r.channel, "offset").f;
    out.range = f.need_attr(r.channel, "range").f;
    out.sampling_rate = f.need_attr(r.channel, "sampling_rate").f;
    copy_str(out.run_id, sizeof out.run_id, f.need_attr(r.tracking, "run_id").s);
    copy_str(out.sample_id, sizeof out.sample_id, f.need_attr(r.tracking, "sample_id").s);
    const uint64_t sig = f.child(r.raw, "Signal");
    if (sig == UNDEF) throw NotFound("read without a Signal dataset");
    const Dataset d = f.dataset(sig);
    if (d.dims.size() != 1) bad("Signal is not one-dimensional");
    if (d.dims[0] > MAX_SIGNAL) bad("implausible Signal length %llu", (unsigned long long)d.dims[0]);
    if ((d.layout == 0 || (d.layout == 1 && d.addr != UNDEF)))
        f.m.at(d.addr, d.dims[0] * 2);             // contiguous data must lie inside the file
    else if (d.dims[0] * 2 > f.m.n * 256)          // no filter here packs better than 256:1
        bad("Signal length %llu is implausible for a %llu-byte file",
            (unsigned long long)d.dims[0], (unsigned long long)f.m.n);
    out.signal_length = (int64_t)d.dims[0];
    if (sig_out) *sig_out = d;
}

pb2f_file *open_file(const char *path)
{
    std::unique_ptr<pb2f_file> f(new pb2f_file);
    f->path = path;
    f->fd = open(path, O_RDONLY | O_CLOEXEC);
    if (f->fd < 0) throw std::runtime_error(std::string("cannot open ") + path);
    struct stat st;
    if (fstat(f->fd, &st) != 0) throw std::runtime_error(std::string("cannot stat ") + path);
    if (st.st_size < 96) bad("not an HDF5 file: %s", path);
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, f->fd, 0);
    if (p == MAP_FAILED) throw std::runtime_error(std::string("cannot map ") + path);
    f->m.p = (const uint8_t *)p;
    f->m.n = (uint64_t)st.st_size;
    close(f->fd);                                  // the mapping outlives the descriptor: a batch of
    f->fd = -1;                                    // single-read files must not exhaust RLIMIT_NOFILE
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (memcmp(f->m.p, sig, 8) != 0) bad("not an HDF5 file: %s", path);
    if (f->m.u8(8) != 0) bad("unsupported superblock version %d", f->m.u8(8));
    if (f->m.u8(13) != 8 || f->m.u8(14) != 8) bad("only 8-byte offsets / lengths are supported");
    if (f->m.u64(24) != 0) bad("non-zero base address");
    f->root = f->m.u64(56 + 8);
    f->multiread = f->child(f->root, "UniqueGlobalKey") == UNDEF;
    return f.release();
}

template <class F>
int guarded(F &&fn)
{
    try {
        return fn();
    } catch (const FormatError &e) {
        g_error = e.what();
        return PB2F_EFORMAT;
    } catch (const NotFound &e) {
        g_error = e.what();
        return PB2F_ENOTFOUND;
    } catch (const std::bad_alloc &) {
        g_error = "out of memory";
        return PB2F_EIO;
    } catch (const std::exception &e) {
        g_error = e.what();
        return PB2F_EIO;
    }
}

template <class F>
void parallel_for(int64_t n, int n_threads, F &&fn)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = (int)(n > 0 ? n : 1);
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        Scratch s;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            fn(i, s);
        }
    };
    if (n_threads == 1) { work(); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work);
    for (std::thread &t : pool) t.join();
}

}  // namespace

struct pb2f_batch {
    int64_t n = 0;
    std::vector<std::unique_ptr<pb2f_file>> files;
    std::vector<int> file_of;                 // per read: index into files, -1 = none
    std::vector<std::string> read_ids;
    std::vector<int32_t> status;
    std::vector<pb2f_read_meta> meta;
    std::vector<Dataset> signal;
};

extern "C" {

int pb2f_abi_version(void) { return PB2F_ABI_VERSION; }
const char *pb2f_last_error(void) { return g_error.c_str(); }

int pb2f_open(const char *path, pb2f_file **out)
{
    if (!path || !out) return PB2F_EINVAL;
    *out = nullptr;
    return guarded([&]() { *out = open_file(path); return PB2F_OK; });
}

void pb2f_close(pb2f_file *f) { delete f; }

int pb2f_is_multiread(const pb2f_file *f) { return f && f->multiread ? 1 : 0; }

int64_t pb2f_num_reads(pb2f_file *f)
{
    if (!f) return PB2F_EINVAL;
    return guarded([&]() {
        std::lock_guard<std::mutex> lock(f->mu);
        if (!f->names_loaded) {
            std::vector<std::string> all;
            if (f->multiread) {
                f->list(f->root, all);
                for (const std::string &s : all)
                    if (s.compare(0, 5, "read_") == 0) f->names.push_back(s.substr(5));
            } else {
                f->list(f->lookup("Raw/Reads"), f->names);
            }
            f->names_loaded = true;
        }
        return (int)f->names.size();
    });
}

const char *pb2f_read_name(pb2f_file *f, int64_t index)
{
    if (!f || pb2f_num_reads(f) < 0 || index < 0 || index >= (int64_t)f->names.size()) return nullptr;
    return f->names[(size_t)index].c_str();
}

int pb2f_read_meta_get(pb2f_file *f, const char *read_id, pb2f_read_meta *out)
{
    if (!f || !out) return PB2F_EINVAL;
    return guarded([&]() { load_meta(*f, read_id, *out, nullptr); return PB2F_OK; });
}

int64_t pb2f_read_signal(pb2f_file *f, const char *read_id, int16_t *dst, int64_t capacity)
{
    if (!f || (!dst && capacity > 0)) return PB2F_EINVAL;
    int64_t n = 0;
    const int rc = guarded([&]() {
        pb2f_read_meta meta;
        Dataset d;
        load_meta(*f, read_id, meta, &d);
        n = meta.signal_length;
        if (n > capacity) { g_error = "destination too small"; return PB2F_ENOSPC; }
        Scratch s;
        read_int16(*f, d, dst, s);
        return PB2F_OK;
    });
    return rc < 0 ? rc : n;
}

int pb2f_batch_open(const char *const *paths, const char *const *read_ids, int64_t n_reads,
                    int n_threads, pb2f_batch **out)
{
    if (!out || n_reads < 0 || (n_reads > 0 && (!paths || !read_ids))) return PB2F_EINVAL;
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<pb2f_batch> b(new pb2f_batch);
        b->n = n_reads;
        b->file_of.assign((size_t)n_reads, -1);
        b->status.assign((size_t)n_reads, PB2F_READ_OK);
        b->meta.resize((size_t)n_reads);
        b->signal.resize((size_t)n_reads);
        b->read_ids.resize((size_t)n_reads);
        std::map<std::string, int> index;
        std::vector<std::string> distinct;
        for (int64_t i = 0; i < n_reads; i++) {
            if (!paths[i]) { b->status[(size_t)i] = PB2F_READ_DISAPPEARED; continue; }
            if (read_ids[i]) b->read_ids[(size_t)i] = read_ids[i];
            auto it = index.find(paths[i]);
            if (it == index.end()) {
                it = index.emplace(paths[i], (int)distinct.size()).first;
                distinct.push_back(paths[i]);
            }
            b->file_of[(size_t)i] = it->second;
        }
        // open each distinct file once (0 = ok, 1 = missing, 2 = unreadable)
        b->files.resize(distinct.size());
        std::vector<int> fstat_((size_t)distinct.size(), 0);
        parallel_for((int64_t)distinct.size(), n_threads, [&](int64_t k, Scratch &) {
            if (access(distinct[(size_t)k].c_str(), F_OK) != 0) { fstat_[(size_t)k] = 1; return; }
            try { b->files[(size_t)k].reset(open_file(distinct[(size_t)k].c_str())); }
            catch (const std::exception &) { fstat_[(size_t)k] = 2; }
        });
        parallel_for(n_reads, n_threads, [&](int64_t i, Scratch &) {
            const int k = b->file_of[(size_t)i];
            if (k < 0) return;
            if (fstat_[(size_t)k]) {
                b->status[(size_t)i] = fstat_[(size_t)k] == 1 ? PB2F_READ_DISAPPEARED : PB2F_READ_IRREGULAR;
                return;
            }
            try {
                const std::string &rid = b->read_ids[(size_t)i];
                load_meta(*b->files[(size_t)k], read_ids[i] ? rid.c_str() : nullptr,
                          b->meta[(size_t)i], &b->signal[(size_t)i]);
            } catch (const std::exception &) {
                b->status[(size_t)i] = PB2F_READ_IRREGULAR;      // signal_loader.py:200-207
                memset(&b->meta[(size_t)i], 0, sizeof(pb2f_read_meta));
            }
        });
        *out = b.release();
        return PB2F_OK;
    });
}

int pb2f_batch_meta(const pb2f_batch *b, int32_t *status, int64_t *signal_length, double *range,
                    double *digitisation, double *offset, double *sampling_rate, int64_t *duration,
                    int64_t *start_time)
{
    if (!b) return PB2F_EINVAL;
    for (int64_t i = 0; i < b->n; i++) {
        const pb2f_read_meta &m = b->meta[(size_t)i];
        if (status) status[i] = b->status[(size_t)i];
        if (signal_length) signal_length[i] = m.signal_length;
        if (range) range[i] = m.range;
        if (digitisation) digitisation[i] = m.digitisation;
        if (offset) offset[i] = m.offset;
        if (sampling_rate) sampling_rate[i] = m.sampling_rate;
        if (duration) duration[i] = m.duration;
        if (start_time) start_time[i] = m.start_time;
    }
    return PB2F_OK;
}

int pb2f_batch_meta_full(const pb2f_batch *b, int64_t index, pb2f_read_meta *out)
{
    if (!b || !out || index < 0 || index >= b->n) return PB2F_EINVAL;
    *out = b->meta[(size_t)index];
    return PB2F_OK;
}

int64_t pb2f_batch_plan(const pb2f_batch *b, int64_t *raw_offsets, int64_t *raw_lengths)
{
    if (!b) return PB2F_EINVAL;
    int64_t pos = 0;
    for (int64_t i = 0; i < b->n; i++) {
        const int64_t len = b->status[(size_t)i] == PB2F_READ_OK ? b->meta[(size_t)i].signal_length : 0;
        if (raw_offsets) raw_offsets[i] = pos;
        if (raw_lengths) raw_lengths[i] = len;
        pos += (len + 7) & ~(int64_t)7;                 // next read on a 16-byte boundary
    }
    return pos;
}

int pb2f_batch_read(pb2f_batch *b, int16_t *raw, int64_t raw_capacity, const int64_t *raw_offsets,
                    int n_threads)
{
    if (!b || !raw_offsets || (!raw && raw_capacity > 0)) return PB2F_EINVAL;
    for (int64_t i = 0; i < b->n; i++) {
        if (b->status[(size_t)i] != PB2F_READ_OK) continue;
        const int64_t len = b->meta[(size_t)i].signal_length;
        if (raw_offsets[i] < 0 || raw_offsets[i] + len > raw_capacity) {
            g_error = "destination too small";
            return PB2F_ENOSPC;
        }
    }
    parallel_for(b->n, n_threads, [&](int64_t i, Scratch &s) {
        if (b->status[(size_t)i] != PB2F_READ_OK) return;
        try {
            read_int16(*b->files[(size_t)b->file_of[(size_t)i]], b->signal[(size_t)i],
                       raw + raw_offsets[i], s);
        } catch (const std::exception &) {
            b->status[(size_t)i] = PB2F_READ_IRREGULAR;
        }
    });
    return PB2F_OK;
}

void pb2f_batch_close(pb2f_batch *b) { delete b; }

int64_t pb2f_inflate(const void *src, int64_t src_len, void *dst, int64_t dst_capacity)
{
    if (!src || src_len < 0 || (!dst && dst_capacity > 0) || dst_capacity < 0) return PB2F_EINVAL;
    static thread_local std::unique_ptr<pbinf::Tables> tables;
    if (!tables) tables.reset(new pbinf::Tables);
    const int64_t got = pbinf::zlib_decompress((const uint8_t *)src, (size_t)src_len, (uint8_t *)dst,
                                               (size_t)dst_capacity, *tables);
    if (got == pbinf::ERR_OVERFLOW) return PB2F_ENOSPC;
    return got < 0 ? PB2F_EFORMAT : got;
}

// ---- streamvbyte-16 streams for the compressed upload (pb2_batch.packed) --------------------
// size of the stream of `count` samples is ceil(count / 8) + count + #(values > 255); streams
// start on 16-byte boundaries.
static inline uint16_t svb16_code(int16_t x, int16_t prev)
{
    const uint16_t d = (uint16_t)((uint16_t)x - (uint16_t)prev);
    return (uint16_t)((d << 1) ^ (uint16_t)(0 - (d >> 15)));          // zigzag of the 16-bit delta
}

int64_t pb2f_svb16_plan(const int16_t *raw, const int64_t *raw_offsets, const int64_t *raw_lengths,
                        int64_t n_reads, int n_threads, int64_t *packed_offsets)
{
    if (n_reads < 0 || !packed_offsets || (n_reads > 0 && (!raw || !raw_offsets || !raw_lengths)))
        return PB2F_EINVAL;
    std::vector<int64_t> size((size_t)n_reads);
    parallel_for(n_reads, n_threads, [&](int64_t i, Scratch &) {
        const int16_t *x = raw + raw_offsets[i];
        const int64_t n = raw_lengths[i];
        int64_t wide = 0;
        int16_t prev = 0;
        for (int64_t k = 0; k < n; k++) { wide += svb16_code(x[k], prev) > 0xFF; prev = x[k]; }
        size[(size_t)i] = (n + 7) / 8 + n + wide;
    });
    int64_t pos = 0;
    for (int64_t i = 0; i < n_reads; i++) {
        packed_offsets[i] = pos;
        pos += (size[(size_t)i] + 15) & ~(int64_t)15;
    }
    packed_offsets[n_reads] = pos;
    return pos;
}

int pb2f_svb16_encode(const int16_t *raw, const int64_t *raw_offsets, const int64_t *raw_lengths,
                      int64_t n_reads, int n_threads, const int64_t *packed_offsets, uint8_t *packed)
{
    if (n_reads < 0 || (n_reads > 0 && (!raw || !raw_offsets || !raw_lengths || !packed_offsets || !packed)))
        return PB2F_EINVAL;
    parallel_for(n_reads, n_threads, [&](int64_t i, Scratch &) {
        const int16_t *x = raw + raw_offsets[i];
        const int64_t n = raw_lengths[i], nkeys = (n + 7) / 8;
        uint8_t *keys = packed + packed_offsets[i], *data = keys + nkeys;
        uint8_t *end = packed + packed_offsets[i + 1];
        memset(keys, 0, (size_t)nkeys);
        int16_t prev = 0;
        for (int64_t k = 0; k < n; k++) {
            const uint16_t v = svb16_code(x[k], prev);
            prev = x[k];
            *data++ = (uint8_t)v;
            if (v > 0xFF) { *data++ = (uint8_t)(v >> 8); keys[k >> 3] |= (uint8_t)(1u << (k & 7)); }
        }
        while (data < end) *data++ = 0;                               // alignment padding
    });
    return PB2F_OK;
}

}  // extern "C"
