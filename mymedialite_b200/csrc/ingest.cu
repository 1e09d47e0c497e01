// ingest.cu -- text rating / implicit-feedback files -> id mapping -> COO in pinned host memory.
//
// Host-side C++ (threads, no kernels): the step immediately before mml_ratings_create / mml_feedback_create.
// Reference behaviour restated here:
//   IO/StaticRatingData.cs:74-117, IO/RatingData.cs:57-88   rating files: skip lines of length 0, split every other line
//                                                            at '\t', ' ' and ',' (IO/Constants.cs:25; empty tokens are
//                                                            kept, as string.Split does), at least 3 (2) columns or a
//                                                            FormatException, float.Parse(InvariantCulture) of column 2
//   IO/ItemData.cs:59-93                                     feedback files: skip lines that are empty after Trim(),
//                                                            at least 2 columns
//   Data/IdentityMapping.cs:62-67                            internal id = int.Parse(token)
//   Data/Mapping.cs:75-85                                    internal id = order of first appearance of the token
//   TextReader.ReadLine                                      line ends at "\n", "\r" or "\r\n"
//
// The file is cut at line boundaries into one chunk per thread. Chunks are parsed independently (tokens, numbers,
// a chunk-local first-seen dictionary); the dictionaries are then merged in chunk order, which reproduces the global
// first-seen numbering exactly, and the triples are written to their final positions in parallel. The destination is
// cudaHostAlloc'ed memory when a CUDA device is present (so the upload that follows is one DMA per array), ordinary
// memory otherwise (CPU-only test boxes: parsing is host work, there is nothing to fall back from).
#include "common.cuh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <new>
#include <string_view>
#include <thread>
#include <unordered_map>

namespace mml {

namespace {

struct IdMap {                                   // Data/Mapping.cs: original_to_internal + internal_to_original
    std::unordered_map<std::string, int32_t> to_internal;
    std::vector<std::string> to_original;
};

struct Ingest {
    int32_t kind = MML_FILE_RATINGS;
    int32_t map_kind[2] = {MML_MAP_IDENTITY, MML_MAP_IDENTITY};
    int64_t n = 0;
    int32_t max_id[2] = {-1, -1};
    bool pinned = false;
    int32_t* users = nullptr;
    int32_t* items = nullptr;
    float* values = nullptr;
    IdMap map[2];
    double parse_seconds = 0.0;
    ~Ingest() {
        auto rel = [&](void* p) { if (!p) return; if (pinned) cudaFreeHost(p); else free(p); };
        rel(users); rel(items); rel(values);
    }
};

inline bool is_split(char c) { return c == '\t' || c == ' ' || c == ','; }          // IO/Constants.cs:25
inline bool is_net_white(char c) { return (c >= 0x09 && c <= 0x0D) || c == 0x20; }   // Char.IsWhiteSpace, ASCII part

inline std::string_view trim_white(std::string_view t)
{
    while (!t.empty() && is_net_white(t.front())) t.remove_prefix(1);
    while (!t.empty() && is_net_white(t.back())) t.remove_suffix(1);
    return t;
}

// int.Parse(string): NumberStyles.Integer = leading/trailing white space, leading sign, decimal digits; anything else
// is a FormatException, a value outside Int32 an OverflowException.
bool parse_int32(std::string_view t, int32_t* out)
{
    t = trim_white(t);
    if (t.empty()) return false;
    bool neg = false;
    if (t.front() == '-' || t.front() == '+') { neg = t.front() == '-'; t.remove_prefix(1); }
    if (t.empty()) return false;
    int64_t v = 0;
    for (char c : t) {
        if (c < '0' || c > '9') return false;
        v = v * 10 + (c - '0');
        if (v > (int64_t)1 << 31) return false;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return false;
    *out = (int32_t)v;
    return true;
}

// float.Parse(string, CultureInfo.InvariantCulture): NumberStyles.Float | AllowThousands = white space, leading sign,
// digits with an optional '.', optional exponent; or one of the symbols NaN / Infinity / -Infinity. The .NET Framework
// and Mono parse to a double and cast (Number.ParseSingle), which is what happens here; an infinite result from a finite
// literal is an OverflowException there, an error here.
bool parse_single(std::string_view t, float* out)
{
    t = trim_white(t);
    if (t.empty()) return false;
    if (t == "NaN") { *out = std::nanf(""); return true; }
    if (t == "Infinity") { *out = INFINITY; return true; }
    if (t == "-Infinity") { *out = -INFINITY; return true; }
    bool neg = false;
    if (t.front() == '-' || t.front() == '+') { neg = t.front() == '-'; t.remove_prefix(1); }
    size_t p = 0, digits = 0;
    while (p < t.size() && t[p] >= '0' && t[p] <= '9') { p++; digits++; }
    if (p < t.size() && t[p] == '.') {
        p++;
        while (p < t.size() && t[p] >= '0' && t[p] <= '9') { p++; digits++; }
    }
    if (digits == 0) return false;
    if (p < t.size() && (t[p] == 'e' || t[p] == 'E')) {
        size_t q = p + 1;
        if (q < t.size() && (t[q] == '-' || t[q] == '+')) q++;
        size_t ed = 0;
        while (q < t.size() && t[q] >= '0' && t[q] <= '9') { q++; ed++; }
        if (ed == 0) return false;
        p = q;
    }
    if (p != t.size()) return false;
    double d = 0.0;
    // "5." and ".5" are valid for both parsers; from_chars wants no leading '+', which was stripped above
    std::string_view body = t;
    std::string tmp;
    if (body.front() == '.') { tmp = "0"; tmp.append(body); body = tmp; }
    auto res = std::from_chars(body.data(), body.data() + body.size(), d, std::chars_format::general);
    if (res.ec == std::errc::result_out_of_range) {
        // from_chars leaves d untouched: underflow rounds to 0 in .NET, overflow throws
        bool big = false;
        size_t e = body.find_first_of("eE");
        if (e != std::string_view::npos) big = body[e + 1] != '-';
        else big = true;
        if (big) return false;
        d = 0.0;
    } else if (res.ec != std::errc()) {
        return false;
    }
    float f = (float)d;
    if (std::isinf(f)) return false;
    *out = neg ? -f : f;
    return true;
}

struct Chunk {
    const char* begin = nullptr;
    const char* end = nullptr;
    std::vector<int32_t> ids[2];                 // identity: parsed ids; first-seen: chunk-local ids
    std::vector<float> values;
    std::vector<std::string_view> local[2];      // first-seen: tokens in order of first appearance inside the chunk
    std::vector<int32_t> to_global[2];
    int64_t err_offset = -1;                     // byte offset (from the start of the text) of the first bad line
    std::string err_msg;
    int64_t out_offset = 0;
};

std::string clip_line(std::string_view line)
{
    return std::string(line.substr(0, std::min<size_t>(line.size(), 200)));
}

void parse_chunk(Chunk& c, const char* text0, int32_t kind, const int32_t map_kind[2])
{
    const int need = kind == MML_FILE_RATINGS ? 3 : 2;
    std::unordered_map<std::string_view, int32_t> dict[2];
    const size_t guess = (size_t)(c.end - c.begin) / 12 + 16;
    c.ids[0].reserve(guess); c.ids[1].reserve(guess);
    if (kind == MML_FILE_RATINGS) c.values.reserve(guess);
    const char* p = c.begin;
    while (p < c.end) {
        const char* e = p;
        while (e < c.end && *e != '\n' && *e != '\r') e++;
        std::string_view line(p, (size_t)(e - p));
        const char* next = e;
        if (next < c.end) next += (*e == '\r' && e + 1 < c.end && e[1] == '\n') ? 2 : 1;
        p = next;
        if (line.empty()) continue;                                               // StaticRatingData.cs:100-101
        if (kind == MML_FILE_FEEDBACK && trim_white(line).empty()) continue;      // ItemData.cs:73-74
        std::string_view tok[3];
        int n_tok = 0;
        size_t s = 0;
        for (size_t i = 0; i <= line.size(); i++) {
            if (i == line.size() || is_split(line[i])) {
                if (n_tok < 3) tok[n_tok] = line.substr(s, i - s);
                n_tok++;
                s = i + 1;
            }
        }
        auto fail = [&](std::string msg) {
            c.err_offset = (int64_t)(line.data() - text0);
            c.err_msg = std::move(msg);
        };
        if (n_tok < need) {                                                       // StaticRatingData.cs:105-108
            fail("Expected at least " + std::to_string(need) + " columns: " + clip_line(line));
            return;
        }
        int32_t id[2];
        for (int w = 0; w < 2; w++) {
            if (map_kind[w] == MML_MAP_IDENTITY) {
                if (!parse_int32(tok[w], &id[w])) { fail("Could not read line '" + clip_line(line) + "'"); return; }
                if (id[w] < 0) { fail("Negative entity ID in line '" + clip_line(line) + "'"); return; }
            } else {
                auto it = dict[w].find(tok[w]);
                if (it == dict[w].end()) {
                    id[w] = (int32_t)c.local[w].size();
                    dict[w].emplace(tok[w], id[w]);
                    c.local[w].push_back(tok[w]);
                } else {
                    id[w] = it->second;
                }
            }
        }
        float v = 0.f;
        if (kind == MML_FILE_RATINGS && !parse_single(tok[2], &v)) {
            fail("Could not read line '" + clip_line(line) + "'");
            return;
        }
        c.ids[0].push_back(id[0]); c.ids[1].push_back(id[1]);
        if (kind == MML_FILE_RATINGS) c.values.push_back(v);
    }
}

template <typename F>
void run_parallel(int n_threads, int n_jobs, F f)
{
    if (n_threads <= 1 || n_jobs <= 1) { for (int j = 0; j < n_jobs; j++) f(j); return; }
    std::vector<std::thread> th;
    th.reserve(n_jobs);
    for (int j = 0; j < n_jobs; j++) th.emplace_back([&f, j] { f(j); });
    for (auto& t : th) t.join();
}

template <typename T>
int32_t host_alloc(T** p, size_t count, bool* pinned, bool first)
{
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (first) {
        int n_dev = 0;
        *pinned = cudaGetDeviceCount(&n_dev) == cudaSuccess && n_dev > 0;
        if (!*pinned) (void)cudaGetLastError();
    }
    if (*pinned) {
        if (cudaHostAlloc((void**)p, bytes, cudaHostAllocDefault) == cudaSuccess) return MML_OK;
        (void)cudaGetLastError();
        if (!first) { set_error("mml_ingest: cudaHostAlloc of %zu bytes failed", bytes); return MML_ERR_CUDA; }
        *pinned = false;
    }
    *p = (T*)malloc(bytes);
    MML_CHECK(*p != nullptr, MML_ERR_ARG, "mml_ingest: out of host memory (%zu bytes)", bytes);
    return MML_OK;
}

int32_t ingest_text(const char* text, int64_t len, int32_t kind, int32_t user_mapping, int32_t item_mapping,
                    int32_t ignore_first_line, int32_t n_threads, const Ingest* prior, Ingest** out)
{
    MML_CHECK(out != nullptr, MML_ERR_ARG, "mml_ingest: out is NULL");
    MML_CHECK(len >= 0 && (text != nullptr || len == 0), MML_ERR_ARG, "mml_ingest: bad text buffer");
    MML_CHECK(kind >= MML_FILE_RATINGS && kind <= MML_FILE_FEEDBACK, MML_ERR_ARG, "mml_ingest: unknown file kind %d", kind);
    const int32_t map_kind[2] = {user_mapping, item_mapping};
    for (int w = 0; w < 2; w++) {
        MML_CHECK(map_kind[w] == MML_MAP_IDENTITY || map_kind[w] == MML_MAP_FIRST_SEEN, MML_ERR_ARG,
                  "mml_ingest: unknown mapping kind %d", map_kind[w]);
        MML_CHECK(!prior || prior->map_kind[w] == map_kind[w], MML_ERR_ARG,
                  "mml_ingest: the prior ingest used another mapping kind");
    }
    if (n_threads <= 0) n_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    n_threads = std::min(n_threads, 256);

    const char* begin = text;
    const char* end = text + len;
    // StreamReader(filename) detects and drops a UTF-8 byte order mark
    if (len >= 3 && (unsigned char)begin[0] == 0xEF && (unsigned char)begin[1] == 0xBB && (unsigned char)begin[2] == 0xBF) begin += 3;
    if (ignore_first_line) {                                                      // reader.ReadLine() once
        while (begin < end && *begin != '\n' && *begin != '\r') begin++;
        if (begin < end) begin += (*begin == '\r' && begin + 1 < end && begin[1] == '\n') ? 2 : 1;
    }
    // chunks of whole lines; at least 64 KiB each so that tiny files stay on one thread
    const int64_t body = end - begin;
    int n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, body / (64 << 10)));
    std::vector<Chunk> chunks(n_chunks);
    const char* cur = begin;
    for (int c = 0; c < n_chunks; c++) {
        chunks[c].begin = cur;
        const char* stop = end;
        if (c + 1 < n_chunks) {
            stop = std::max(cur, begin + body * (c + 1) / n_chunks);
            while (stop < end && *stop != '\n' && *stop != '\r') stop++;
            if (stop < end) stop += (*stop == '\r' && stop + 1 < end && stop[1] == '\n') ? 2 : 1;
        }
        chunks[c].end = stop;
        cur = stop;
    }
    run_parallel(n_threads, n_chunks, [&](int c) { parse_chunk(chunks[c], text, kind, map_kind); });
    for (auto& c : chunks)                         // chunks are in file order: the first one with an error holds the first bad line
        if (c.err_offset >= 0) { set_error("%s", c.err_msg.c_str()); return MML_ERR_FORMAT; }

    Ingest* g = new (std::nothrow) Ingest();
    MML_CHECK(g != nullptr, MML_ERR_ARG, "mml_ingest: out of memory");
    g->kind = kind;
    g->map_kind[0] = map_kind[0]; g->map_kind[1] = map_kind[1];
    if (prior) { g->map[0] = prior->map[0]; g->map[1] = prior->map[1]; }   // MaxUserID/MaxItemID stay the data set's own

    // merge the chunk dictionaries in file order = global order of first appearance (Data/Mapping.cs:80-83)
    run_parallel(n_threads, 2, [&](int w) {
        if (map_kind[w] != MML_MAP_FIRST_SEEN) return;
        IdMap& m = g->map[w];
        for (auto& c : chunks) {
            c.to_global[w].resize(c.local[w].size());
            for (size_t j = 0; j < c.local[w].size(); j++) {
                std::string key(c.local[w][j]);
                auto it = m.to_internal.find(key);
                if (it == m.to_internal.end()) {
                    const int32_t id = (int32_t)m.to_original.size();
                    m.to_internal.emplace(key, id);
                    m.to_original.push_back(std::move(key));
                    c.to_global[w][j] = id;
                } else {
                    c.to_global[w][j] = it->second;
                }
            }
        }
    });

    int64_t n = 0;
    for (auto& c : chunks) { c.out_offset = n; n += (int64_t)c.ids[0].size(); }
    g->n = n;
    int32_t st = host_alloc(&g->users, (size_t)n, &g->pinned, true);
    if (st == MML_OK) st = host_alloc(&g->items, (size_t)n, &g->pinned, false);
    if (st == MML_OK && kind == MML_FILE_RATINGS) st = host_alloc(&g->values, (size_t)n, &g->pinned, false);
    if (st != MML_OK) { delete g; return st; }

    std::vector<int32_t> cmax0(n_chunks, -1), cmax1(n_chunks, -1);
    run_parallel(n_threads, n_chunks, [&](int ci) {
        Chunk& c = chunks[ci];
        const size_t m = c.ids[0].size();
        int32_t* dst[2] = {g->users + c.out_offset, g->items + c.out_offset};
        int32_t mx[2] = {-1, -1};
        for (int w = 0; w < 2; w++) {
            const int32_t* src = c.ids[w].data();
            if (map_kind[w] == MML_MAP_FIRST_SEEN) {
                const int32_t* tg = c.to_global[w].data();
                for (size_t j = 0; j < m; j++) { int32_t id = tg[src[j]]; dst[w][j] = id; mx[w] = std::max(mx[w], id); }
            } else {
                for (size_t j = 0; j < m; j++) { dst[w][j] = src[j]; mx[w] = std::max(mx[w], src[j]); }
            }
        }
        if (kind == MML_FILE_RATINGS && m) memcpy(g->values + c.out_offset, c.values.data(), m * sizeof(float));
        cmax0[ci] = mx[0]; cmax1[ci] = mx[1];
    });
    for (int c = 0; c < n_chunks; c++) {
        g->max_id[0] = std::max(g->max_id[0], cmax0[c]);
        g->max_id[1] = std::max(g->max_id[1], cmax1[c]);
    }
    *out = g;
    return MML_OK;
}

inline Ingest* ingest_of(mml_ingest* h) { return reinterpret_cast<Ingest*>(h); }
inline const Ingest* ingest_of(const mml_ingest* h) { return reinterpret_cast<const Ingest*>(h); }

}  // namespace

}  // namespace mml

using namespace mml;

extern "C" {

int32_t mml_ingest_text(const char* text, int64_t len, int32_t kind, int32_t user_mapping, int32_t item_mapping,
                        int32_t ignore_first_line, int32_t n_threads, const mml_ingest* prior, mml_ingest** out)
{
    Ingest* g = nullptr;
    MML_TRY(ingest_text(text, len, kind, user_mapping, item_mapping, ignore_first_line, n_threads, ingest_of(prior), &g));
    *out = reinterpret_cast<mml_ingest*>(g);
    return MML_OK;
}

int32_t mml_ingest_file(const char* path, int32_t kind, int32_t user_mapping, int32_t item_mapping,
                        int32_t ignore_first_line, int32_t n_threads, const mml_ingest* prior, mml_ingest** out)
{
    MML_CHECK(path != nullptr && out != nullptr, MML_ERR_ARG, "mml_ingest_file: NULL argument");
    const int fd = open(path, O_RDONLY);
    MML_CHECK(fd >= 0, MML_ERR_IO, "mml_ingest_file: cannot open '%s': %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        close(fd);
        set_error("mml_ingest_file: '%s' is not a regular file", path);
        return MML_ERR_IO;
    }
    const int64_t len = (int64_t)sb.st_size;
    void* map = nullptr;
    if (len > 0) {
        map = mmap(nullptr, (size_t)len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) {
            close(fd);
            set_error("mml_ingest_file: mmap of '%s' failed: %s", path, strerror(errno));
            return MML_ERR_IO;
        }
        madvise(map, (size_t)len, MADV_SEQUENTIAL);
    }
    Ingest* g = nullptr;
    const int32_t st = ingest_text((const char*)map, len, kind, user_mapping, item_mapping, ignore_first_line, n_threads,
                                   ingest_of(prior), &g);
    if (map) munmap(map, (size_t)len);
    close(fd);
    if (st != MML_OK) return st;
    *out = reinterpret_cast<mml_ingest*>(g);
    return MML_OK;
}

int32_t mml_ingest_destroy(mml_ingest* h)
{
    delete ingest_of(h);
    return MML_OK;
}

int32_t mml_ingest_info(const mml_ingest* h, int64_t* n, int32_t* max_user, int32_t* max_item,
                        int32_t* n_user_ids, int32_t* n_item_ids, int32_t* pinned)
{
    MML_CHECK(h != nullptr, MML_ERR_ARG, "mml_ingest_info: NULL handle");
    const Ingest* g = ingest_of(h);
    if (n) *n = g->n;
    if (max_user) *max_user = g->max_id[0];
    if (max_item) *max_item = g->max_id[1];
    if (n_user_ids) *n_user_ids = (int32_t)g->map[0].to_original.size();
    if (n_item_ids) *n_item_ids = (int32_t)g->map[1].to_original.size();
    if (pinned) *pinned = g->pinned ? 1 : 0;
    return MML_OK;
}

int32_t mml_ingest_copy(const mml_ingest* h, int32_t* users, int32_t* items, float* values)
{
    MML_CHECK(h != nullptr, MML_ERR_ARG, "mml_ingest_copy: NULL handle");
    const Ingest* g = ingest_of(h);
    if (users && g->n) memcpy(users, g->users, sizeof(int32_t) * (size_t)g->n);
    if (items && g->n) memcpy(items, g->items, sizeof(int32_t) * (size_t)g->n);
    if (values) {
        MML_CHECK(g->values != nullptr, MML_ERR_STATE, "mml_ingest_copy: this file kind has no rating column");
        if (g->n) memcpy(values, g->values, sizeof(float) * (size_t)g->n);
    }
    return MML_OK;
}

int32_t mml_ingest_original_ids(const mml_ingest* h, int32_t which, int32_t first, int32_t count,
                                char* buf, int64_t buf_len, int64_t* offsets, int64_t* needed)
{
    MML_CHECK(h != nullptr && (which == 0 || which == 1), MML_ERR_ARG, "mml_ingest_original_ids: bad argument");
    const Ingest* g = ingest_of(h);
    MML_CHECK(g->map_kind[which] == MML_MAP_FIRST_SEEN, MML_ERR_STATE,
              "mml_ingest_original_ids: identity mapping keeps no table (original id = internal id)");
    const auto& t = g->map[which].to_original;
    MML_CHECK(first >= 0 && count >= 0 && (size_t)first + (size_t)count <= t.size(), MML_ERR_ARG,
              "Unknown internal ID: %d", first + count - 1);                       // Data/Mapping.cs:68
    int64_t total = 0;
    for (int32_t j = 0; j < count; j++) total += (int64_t)t[first + j].size();
    if (needed) *needed = total;
    if (!buf && !offsets) return MML_OK;
    MML_CHECK(buf_len >= total, MML_ERR_ARG, "mml_ingest_original_ids: buffer of %lld bytes, %lld needed",
              (long long)buf_len, (long long)total);
    int64_t o = 0;
    for (int32_t j = 0; j < count; j++) {
        if (offsets) offsets[j] = o;
        const std::string& s = t[first + j];
        if (buf && !s.empty()) memcpy(buf + o, s.data(), s.size());
        o += (int64_t)s.size();
    }
    if (offsets) offsets[count] = o;
    return MML_OK;
}

int32_t mml_ingest_to_ratings(mml_ctx* ctx, const mml_ingest* h, mml_ratings** out)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(h != nullptr, MML_ERR_ARG, "mml_ingest_to_ratings: NULL handle");
    const Ingest* g = ingest_of(h);
    MML_CHECK(g->kind != MML_FILE_FEEDBACK, MML_ERR_STATE, "mml_ingest_to_ratings: the handle holds a feedback file");
    if (g->values == nullptr) {                   // test files without ratings: value 0 (StaticRatingData.cs:112)
        std::vector<float> zeros((size_t)std::max<int64_t>(g->n, 1), 0.f);
        return mml_ratings_create(ctx, g->users, g->items, zeros.data(), g->n, g->max_id[0], g->max_id[1], out);
    }
    return mml_ratings_create(ctx, g->users, g->items, g->values, g->n, g->max_id[0], g->max_id[1], out);
}

int32_t mml_ingest_to_feedback(mml_ctx* ctx, const mml_ingest* h, mml_feedback** out)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(h != nullptr, MML_ERR_ARG, "mml_ingest_to_feedback: NULL handle");
    const Ingest* g = ingest_of(h);
    return mml_feedback_create(ctx, g->users, g->items, g->n, g->max_id[0], g->max_id[1], out);
}

}  // extern "C"
