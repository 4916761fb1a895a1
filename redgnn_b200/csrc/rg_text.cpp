// Text side of the graph store: "head relation tail" lines -> int32 id triples on the
// host cores.  Restates DataLoader.read_triples' parsing (Static/transductive/load_data.py:58-67,
// Static/inductive/load_data.py:76-86): every line is `line.strip().split()` into exactly three
// names that are looked up in entity2id / relation2id.  No GPU, no stream, no allocation visible
// to the caller; the file is mapped read-only and split at line boundaries over the host threads.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "redgnn_b200.h"

namespace {

// Length in bytes of the whitespace character starting at p (0 if none): what Python's
// str.split() / str.strip() treat as a separator for text decoded from UTF-8 -- ASCII
// \t \n \v \f \r, 0x1c-0x1f, space, and U+0085 U+00A0 U+1680 U+2000-200A U+2028 U+2029 U+202F
// U+205F U+3000 in their UTF-8 encodings.
inline int ws_len(const unsigned char *p, const unsigned char *end) {
    unsigned c = *p;
    if (c <= 0x20) return (c == 0x20 || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f)) ? 1 : 0;
    if (c < 0xc2) return 0;
    if (c == 0xc2) return (end - p >= 2 && (p[1] == 0x85 || p[1] == 0xa0)) ? 2 : 0;
    if (c == 0xe1) return (end - p >= 3 && p[1] == 0x9a && p[2] == 0x80) ? 3 : 0;
    if (c == 0xe2) {
        if (end - p < 3) return 0;
        if (p[1] == 0x80) return (p[2] <= 0x8a && p[2] >= 0x80) || p[2] == 0xa8 || p[2] == 0xa9 || p[2] == 0xaf ? 3 : 0;
        return (p[1] == 0x81 && p[2] == 0x9f) ? 3 : 0;
    }
    if (c == 0xe3) return (end - p >= 3 && p[1] == 0x80 && p[2] == 0x80) ? 3 : 0;
    return 0;
}

inline uint64_t hash_bytes(const unsigned char *p, size_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (n * 0xff51afd7ed558ccdull);
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
        h ^= h >> 29;
        p += 8;
        n -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, p, n);
    h = (h ^ w) * 0xff51afd7ed558ccdull;
    return h ^ (h >> 32);
}

// name -> id, open addressing.  A slot carries the first 8 bytes of its name, so a name of up to 8 bytes
// resolves with the one cache miss of its slot; longer names compare their remaining bytes in the
// caller's blob.  Lookups are split into hash -> prefetch -> resolve so that a batch of lines keeps
// many misses in flight (the dictionaries of the large graphs do not fit any cache).
struct Slot {
    uint64_t prefix;   // first min(len, 8) bytes, zero padded
    uint64_t off;      // byte offset of the name in the blob
    uint32_t len;
    uint32_t tag;      // high hash bits
    int32_t id;
    int32_t used;
};

inline uint64_t load_prefix(const unsigned char *p, size_t n) {
    uint64_t w = 0;
    memcpy(&w, p, n < 8 ? n : 8);
    return w;
}

struct NameIndex {
    const char *bytes = nullptr;
    std::vector<Slot> slot;
    uint64_t mask = 0;

    inline bool same(const Slot &s, const unsigned char *p, size_t n, uint64_t prefix, uint32_t tag) const {
        return s.tag == tag && s.len == n && s.prefix == prefix &&
               (n <= 8 || memcmp(bytes + s.off + 8, p + 8, n - 8) == 0);
    }

    bool build(const rg_name_table *tab) {
        bytes = tab->bytes;
        uint64_t cap = 16;
        while (cap < (uint64_t)tab->n * 2 + 2) cap <<= 1;
        mask = cap - 1;
        slot.assign(cap, Slot{0, 0, 0, 0, 0, 0});
        const int64_t AHEAD = 16;
        std::vector<uint64_t> hs((size_t)std::min<int64_t>(tab->n, AHEAD));
        auto hash_of = [&](int64_t k, uint64_t *h) {
            int64_t lo = tab->off[k], hi = tab->off[k + 1];
            if (lo < 0 || hi < lo || hi - lo > (int64_t)UINT32_MAX) return false;
            *h = hash_bytes((const unsigned char *)tab->bytes + lo, (size_t)(hi - lo));
            __builtin_prefetch(&slot[*h & mask], 1);
            return true;
        };
        for (int64_t k = 0; k < (int64_t)hs.size(); k++)
            if (!hash_of(k, &hs[k])) return false;
        for (int64_t k = 0; k < tab->n; k++) {
            uint64_t h = hs[k % AHEAD];
            if (k + AHEAD < tab->n && !hash_of(k + AHEAD, &hs[k % AHEAD])) return false;
            const int64_t lo = tab->off[k];
            const size_t n = (size_t)(tab->off[k + 1] - lo);
            const unsigned char *p = (const unsigned char *)tab->bytes + lo;
            const uint64_t prefix = load_prefix(p, n);
            const uint32_t tag = (uint32_t)(h >> 32);
            uint64_t i = h & mask;
            // same name again: the later entry wins, like a dict assignment
            while (slot[i].used && !same(slot[i], p, n, prefix, tag)) i = (i + 1) & mask;
            slot[i] = Slot{prefix, (uint64_t)lo, (uint32_t)n, tag, tab->id[k], 1};
        }
        return true;
    }

    inline uint64_t prepare(const unsigned char *p, size_t n) const {
        uint64_t h = hash_bytes(p, n);
        __builtin_prefetch(&slot[h & mask], 0);
        return h;
    }

    inline bool resolve(uint64_t h, const unsigned char *p, size_t n, int32_t *id) const {
        const uint64_t prefix = load_prefix(p, n);
        const uint32_t tag = (uint32_t)(h >> 32);
        for (uint64_t i = h & mask;; i = (i + 1) & mask) {
            const Slot &s = slot[i];
            if (!s.used) return false;
            if (same(s, p, n, prefix, tag)) {
                *id = s.id;
                return true;
            }
        }
    }
};

struct Mapped {
    const unsigned char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    int open_file(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return RG_ERR_IO;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return RG_ERR_IO;
        n = (size_t)st.st_size;
        if (n == 0) return RG_OK;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);   // one pass instead of a fault per 64 KB
        if (m == MAP_FAILED) return RG_ERR_IO;
        p = (const unsigned char *)m;
        madvise(m, n, MADV_SEQUENTIAL);
        return RG_OK;
    }
    ~Mapped() {
        if (p) munmap((void *)p, n);
        if (fd >= 0) ::close(fd);
    }
};

// Lines as Python's text-mode iteration yields them (universal newlines): terminated by "\n",
// "\r\n" or a lone "\r"; a last line without terminator counts when it is not empty.
// first position after the terminator that ends the line containing `pos`
inline size_t next_line_start(const unsigned char *p, size_t n, size_t pos) {
    while (pos < n) {
        unsigned char c = p[pos++];
        if (c == '\n') return pos;
        if (c == '\r') return (pos < n && p[pos] == '\n') ? pos + 1 : pos;
    }
    return n;
}

inline int64_t count_lines(const unsigned char *p, size_t lo, size_t hi, size_t n) {
    int64_t lines = 0;
    size_t pos = lo;
    while (pos < hi) {
        size_t nx = next_line_start(p, n, pos);
        lines++;
        pos = nx;
    }
    return lines;
}

int resolve_threads(int32_t n_threads, size_t bytes) {
    int t = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    size_t by_size = bytes / (1u << 20) + 1;   // no point in a thread per few KB
    if ((size_t)t > by_size) t = (int)by_size;
    return t;
}

void chunk_starts(const Mapped &f, int T, std::vector<size_t> *start) {
    start->assign(T + 1, f.n);
    (*start)[0] = 0;
    for (int k = 1; k < T; k++) {
        size_t b = f.n / T * k;
        size_t s = b == 0 ? 0 : next_line_start(f.p, f.n, b - 1);
        (*start)[k] = std::max(s, (*start)[k - 1]);
    }
}

// fn(k) for k in [0, T): chunk 0 on the calling thread, the others on their own threads; a chunk whose
// thread cannot be started runs on the calling thread instead.  fn must not throw.
template <class F>
void run_chunks(int T, F fn) {
    std::vector<std::thread> th;
    th.reserve(T > 0 ? T : 0);
    int started = 1;
    for (int k = 1; k < T; k++) {
        try {
            th.emplace_back(fn, k);
        } catch (...) {
            break;
        }
        started = k + 1;
    }
    fn(0);
    for (int k = started; k < T; k++) fn(k);
    for (auto &t : th) t.join();
}

}  // namespace

static int count_lines_impl(const char *path, int64_t *n_lines);
static int parse_triples_impl(const char *path, const rg_name_table *ent, const rg_name_table *rel, int32_t *out,
                              int64_t cap_rows, int64_t *n_rows, int64_t *err_line, int32_t n_threads);

// nothing is thrown across the C boundary: a failed allocation becomes RG_ERR_HOST
extern "C" {

int rg_text_count_lines(const char *path, int64_t *n_lines) {
    if (!path || !n_lines) return RG_ERR_BAD_ARG;
    try {
        return count_lines_impl(path, n_lines);
    } catch (...) {
        return RG_ERR_HOST;
    }
}

int rg_text_parse_triples(const char *path, const rg_name_table *ent, const rg_name_table *rel, int32_t *out,
                          int64_t cap_rows, int64_t *n_rows, int64_t *err_line, int32_t n_threads) {
    if (!path || !ent || !rel || !n_rows || cap_rows < 0 || (cap_rows > 0 && !out)) return RG_ERR_BAD_ARG;
    try {
        return parse_triples_impl(path, ent, rel, out, cap_rows, n_rows, err_line, n_threads);
    } catch (...) {
        return RG_ERR_HOST;
    }
}

}  // extern "C"

static int count_lines_impl(const char *path, int64_t *n_lines) {
    Mapped f;
    int rc = f.open_file(path);
    if (rc != RG_OK) return rc;
    int T = resolve_threads(0, f.n);
    std::vector<size_t> start;
    chunk_starts(f, T, &start);
    std::vector<int64_t> cnt(T, 0);
    run_chunks(T, [&](int k) { cnt[k] = count_lines(f.p, start[k], start[k + 1], f.n); });
    int64_t tot = 0;
    for (int k = 0; k < T; k++) tot += cnt[k];
    *n_lines = tot;
    return RG_OK;
}

static int parse_triples_impl(const char *path, const rg_name_table *ent, const rg_name_table *rel, int32_t *out,
                              int64_t cap_rows, int64_t *n_rows, int64_t *err_line, int32_t n_threads) {
    if (ent->n < 0 || rel->n < 0 || ent->n > INT32_MAX || rel->n > INT32_MAX) return RG_ERR_BAD_ARG;
    if ((ent->n > 0 && (!ent->bytes || !ent->off || !ent->id)) || (rel->n > 0 && (!rel->bytes || !rel->off || !rel->id)))
        return RG_ERR_BAD_ARG;
    *n_rows = 0;
    if (err_line) *err_line = -1;
    Mapped f;
    int rc = f.open_file(path);
    if (rc != RG_OK) return rc;
    NameIndex E, R;
    if (!E.build(ent) || !R.build(rel)) return RG_ERR_BAD_ARG;

    int T = resolve_threads(n_threads, f.n);
    std::vector<size_t> start;
    chunk_starts(f, T, &start);
    std::vector<int64_t> cnt(T, 0), base(T + 1, 0);
    run_chunks(T, [&](int k) { cnt[k] = count_lines(f.p, start[k], start[k + 1], f.n); });
    for (int k = 0; k < T; k++) base[k + 1] = base[k] + cnt[k];
    *n_rows = base[T];
    if (base[T] > cap_rows) return RG_ERR_BAD_ARG;

    // per chunk: first failing line and its status; the smallest line over all chunks is what a
    // sequential reader (the reference) would have hit first
    std::vector<int64_t> bad_line(T, -1);
    std::vector<int> bad_rc(T, RG_OK);
    auto work = [&](int k) {
        // a batch of lines is split into names and hashed first (prefetching the dictionary slots),
        // then resolved: the cache misses of BATCH lines overlap
        constexpr int BATCH = 16;
        struct Tok {
            const unsigned char *p;
            size_t n;
            uint64_t h;
        };
        Tok tok[BATCH][3];
        int ntok[BATCH];
        size_t pos = start[k];
        const size_t hi = start[k + 1];
        int64_t row = base[k];
        while (pos < hi) {
            int nb = 0;
            for (; nb < BATCH && pos < hi; nb++) {
                size_t nx = next_line_start(f.p, f.n, pos);
                const unsigned char *p = f.p + pos, *end = f.p + nx;
                int nt = 0;
                while (p < end) {
                    int w = ws_len(p, end);
                    if (w) {
                        p += w;
                        continue;
                    }
                    const unsigned char *q = p;
                    while (q < end && !ws_len(q, end)) q++;
                    if (nt == 3) {
                        nt = 4;
                        break;
                    }
                    tok[nb][nt] = Tok{p, (size_t)(q - p), (nt == 1 ? R : E).prepare(p, (size_t)(q - p))};
                    nt++;
                    p = q;
                }
                ntok[nb] = nt;
                pos = nx;
            }
            for (int b = 0; b < nb; b++, row++) {
                // the split is checked before any lookup, and the lookups run in the reference's order
                // (h, r, t), so the first failure of a line is the one the reference raises
                int status = RG_OK;
                int32_t v[3] = {0, 0, 0};
                if (ntok[b] != 3) status = RG_ERR_PARSE;
                for (int j = 0; j < 3 && status == RG_OK; j++)
                    if (!(j == 1 ? R : E).resolve(tok[b][j].h, tok[b][j].p, tok[b][j].n, &v[j])) status = RG_ERR_UNKNOWN_NAME;
                if (status != RG_OK) {
                    bad_line[k] = row;
                    bad_rc[k] = status;
                    return;
                }
                out[3 * row + 0] = v[0];
                out[3 * row + 1] = v[1];
                out[3 * row + 2] = v[2];
            }
        }
    };
    run_chunks(T, work);
    for (int k = 0; k < T; k++)
        if (bad_rc[k] != RG_OK) {
            if (err_line) *err_line = bad_line[k];
            return bad_rc[k];
        }
    return RG_OK;
}
