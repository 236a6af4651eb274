// bdg_tsv.hpp -- host side of the rows either side of the hot path: the extraction-TSV
// reader (reference badger.py:91-111), the whitelist file reader (badger.py:82-88) and the writer of
// `<out>_output_file.tsv` (barcode_graph.py:388-410).  Plain multi-threaded C++ over a memory-mapped file; no CUDA in
// here (the 16-character records this reader hands out are packed and validated by pack16_kernel on the GPU).
// Included by bdg_api.cu inside its anonymous namespace's scope of `fail`.
//
// The reference reads the TSV with pandas.read_csv(sep="\t") and writes with DataFrame.to_csv(sep="\t", index=False).
// This reader reproduces exactly the subset of that behaviour extraction TSVs exercise and REFUSES everything else
// (BDG_ERR_UNSUPPORTED: quotes, carriage returns, NUL / non-ASCII bytes, rows with more fields than the header, ids or
// barcodes pandas would turn into numbers / booleans / NaN): the caller then takes the pandas route, so a file is
// either read with the reference's semantics or not read by this code at all.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace tsvio {

struct Mapped {
    int fd = -1;
    const char* data = nullptr;
    size_t size = 0;
    ~Mapped()
    {
        if (data && size) munmap((void*)data, size);
        if (fd >= 0) close(fd);
    }
    // 0 ok, else errno-style failure (message in err)
    bool open_file(const char* path, std::string& err)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { err = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
        struct stat st;
        if (fstat(fd, &st) != 0) { err = std::string("cannot stat ") + path; return false; }
        size = (size_t)st.st_size;
        if (size == 0) return true;
        void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { data = nullptr; err = std::string("cannot map ") + path + ": " + strerror(errno); return false; }
        data = (const char*)p;
        madvise(p, size, MADV_SEQUENTIAL);
        return true;
    }
};

inline bool field_is(const char* s, size_t n, const char* lit) { return strlen(lit) == n && memcmp(s, lit, n) == 0; }

// pandas._libs.parsers.STR_NA_VALUES (pandas 1.x - 3.x): fields read_csv turns into NaN by default
inline bool is_na_string(const char* s, size_t n)
{
    static const char* const NA[] = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
                                     "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
    for (const char* lit : NA)
        if (field_is(s, n, lit)) return true;
    return false;
}

inline bool ieq(const char* s, size_t n, const char* lower)
{
    if (strlen(lower) != n) return false;
    for (size_t i = 0; i < n; i++)
        if ((char)(s[i] | 0x20) != lower[i]) return false;
    return true;
}

// Could pandas' type inference read this field as a number or a boolean?  Deliberately generous (a superset): a column
// (or one low_memory chunk of it) only changes type when ALL its fields convert, so refusing on ANY such field is safe.
inline bool looks_typed(const char* s, size_t n)
{
    while (n && *s == ' ') { s++; n--; }
    while (n && s[n - 1] == ' ') n--;
    if (n == 0) return true;
    bool numeric = true;
    for (size_t i = 0; i < n && numeric; i++) {
        const char c = s[i];
        numeric = (c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E';
    }
    if (numeric) return true;
    if (ieq(s, n, "true") || ieq(s, n, "false")) return true;
    if (*s == '+' || *s == '-') { s++; n--; }
    return ieq(s, n, "inf") || ieq(s, n, "infinity") || ieq(s, n, "nan");
}

enum : uint8_t { ROW_BARCODE = 1, ROW_EMIT = 2 };

struct RowBlock {                      // what one thread found in its slice of the file
    std::vector<uint64_t> id_off;
    std::vector<uint32_t> id_len;
    std::vector<uint8_t> kind;
    std::vector<char> seqs;            // 16 bytes per row
    std::string why;                   // non-empty: unsupported construct met
};

struct Tsv {
    Mapped file;
    int bc_len = 16;
    size_t rows = 0, emit_rows = 0, barcode_rows = 0;
    std::vector<uint64_t> id_off;
    std::vector<uint32_t> id_len;
    std::vector<uint8_t> kind;
    std::vector<char> seqs;
};

// one slice [beg, end) of the data lines; beg is the start of a line
inline void parse_slice(const char* data, size_t beg, size_t end, int ncols, int id_col, int bc_col, int bc_len, RowBlock& out)
{
    const size_t guess = (end - beg) / 48 + 16;
    out.id_off.reserve(guess); out.id_len.reserve(guess); out.kind.reserve(guess); out.seqs.reserve(guess * 16);
    size_t pos = beg;
    while (pos < end) {
        const char* nl = (const char*)memchr(data + pos, '\n', end - pos);
        const size_t le = nl ? (size_t)(nl - data) : end;
        const size_t ls = pos;
        pos = le + 1;
        // field boundaries; byte checks ride along
        size_t id_s = 0, id_e = 0, bc_s = 0, bc_e = 0;
        bool have_id = false, have_bc = false, only_spaces = true;
        int col = 0;
        size_t fs = ls;
        for (size_t i = ls; i <= le; i++) {
            const unsigned char c = i < le ? (unsigned char)data[i] : (unsigned char)'\t';
            if (c == '\t') {
                if (i < le) only_spaces = false;
                if (col == id_col) { id_s = fs; id_e = i; have_id = true; }
                if (col == bc_col) { bc_s = fs; bc_e = i; have_bc = true; }
                col++;
                fs = i + 1;
            } else if (c != ' ') {
                only_spaces = false;
                if (c == '"' || c == '\r' || c == 0 || c >= 0x80) {
                    out.why = c == '"' ? "quote character" : c == '\r' ? "carriage return" : c == 0 ? "NUL byte" : "non-ASCII byte";
                    return;
                }
            }
        }
        if (only_spaces) continue;                          // read_csv skips blank lines (also lines of spaces only)
        if (col > ncols) { out.why = "row with more fields than the header"; return; }
        if (!have_id) { out.why = "row without a read id field"; return; }
        const char* id = data + id_s;
        const size_t idn = id_e - id_s;
        if (is_na_string(id, idn) || looks_typed(id, idn)) { out.why = "read id that pandas would not keep as text"; return; }
        uint8_t kind = 0;
        char rec[16];
        memset(rec, 'A', 16);
        bool is_hdr_bc = false;
        if (have_bc) {
            const char* bc = data + bc_s;
            const size_t bn = bc_e - bc_s;
            if (!is_na_string(bc, bn)) {                    // NaN -> '*' (fillna) / dropped (dropna): no barcode
                if (looks_typed(bc, bn)) { out.why = "barcode field that pandas would not keep as text"; return; }
                is_hdr_bc = field_is(bc, bn, "barcode");
                if ((bn == (size_t)bc_len || bn == (size_t)bc_len + 1) && !is_hdr_bc) {   // 17-mers lose their last base
                    kind |= ROW_BARCODE;
                    memcpy(rec, bc, 16);
                }
            }
        }
        if (!field_is(id, idn, "#read_id") && !is_hdr_bc) kind |= ROW_EMIT;
        out.id_off.push_back(id_s);
        out.id_len.push_back((uint32_t)idn);
        out.kind.push_back(kind);
        out.seqs.insert(out.seqs.end(), rec, rec + 16);
    }
}

inline int pick_threads(int threads, size_t bytes)
{
    int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    const size_t by_size = bytes / (1u << 20) + 1;          // no point in a thread per few kB
    return (int)std::min<size_t>((size_t)std::min(t, 64), by_size);
}

// returns "" on success, else why the file is refused; *io_error set when the failure is an I/O error, not a refusal
inline std::string tsv_parse(Tsv& t, const char* path, int bc_len, int threads, bool* io_error)
{
    *io_error = false;
    if (bc_len != 16) return "bc_len other than 16";
    std::string err;
    if (!t.file.open_file(path, err)) { *io_error = true; return err; }
    t.bc_len = bc_len;
    const char* d = t.file.data;
    const size_t n = t.file.size;
    // header = first line that is not blank
    size_t pos = 0, hs = 0, he = 0;
    bool found = false;
    while (pos < n && !found) {
        const char* nl = (const char*)memchr(d + pos, '\n', n - pos);
        const size_t le = nl ? (size_t)(nl - d) : n;
        bool blank = true;
        for (size_t i = pos; i < le; i++) blank = blank && d[i] == ' ';
        if (!blank) { hs = pos; he = le; found = true; }
        pos = le + 1;
    }
    if (!found) return "empty file";
    int ncols = 0, id_col = -1, bc_col = -1;
    {
        size_t fs = hs;
        for (size_t i = hs; i <= he; i++) {
            const unsigned char c = i < he ? (unsigned char)d[i] : (unsigned char)'\t';
            if (c == '"' || c == '\r' || c == 0 || c >= 0x80) return "header with a quote, carriage return or non-ASCII byte";
            if (c == '\t') {
                if (id_col < 0 && field_is(d + fs, i - fs, "#read_id")) id_col = ncols;
                if (bc_col < 0 && field_is(d + fs, i - fs, "barcode")) bc_col = ncols;
                if (i == fs) return "unnamed header column";
                ncols++;
                fs = i + 1;
            }
        }
    }
    if (id_col < 0 || bc_col < 0) return "header without #read_id / barcode columns";
    const size_t body = std::min(pos, n);
    const int T = pick_threads(threads, n - body);
    std::vector<size_t> cut(T + 1, n);
    cut[0] = body;
    for (int k = 1; k < T; k++) {                           // slice borders moved to the next line start
        size_t c = body + (n - body) / T * k;
        if (c < cut[k - 1]) c = cut[k - 1];
        const char* nl = c < n ? (const char*)memchr(d + c, '\n', n - c) : nullptr;
        cut[k] = nl ? (size_t)(nl - d) + 1 : n;
    }
    std::vector<RowBlock> blocks(T);
    {
        // an exception must not leave a worker thread (std::terminate): a failed allocation becomes the slice's verdict
        auto guarded = [&](int k) {
            try { parse_slice(d, cut[k], cut[k + 1], ncols, id_col, bc_col, bc_len, blocks[k]); }
            catch (const std::exception&) { blocks[k].why = "\x01out of host memory"; }
        };
        std::vector<std::thread> pool;
        for (int k = 1; k < T; k++) pool.emplace_back(guarded, k);
        guarded(0);
        for (auto& th : pool) th.join();
    }
    size_t rows = 0;
    for (auto& b : blocks) {
        if (!b.why.empty() && b.why[0] == '\x01') throw std::bad_alloc();     // reported as BDG_ERR_OOM by the caller
        if (!b.why.empty()) return b.why;
        rows += b.kind.size();
    }
    t.rows = rows;
    t.id_off.resize(rows); t.id_len.resize(rows); t.kind.resize(rows); t.seqs.resize(rows * 16);
    size_t at = 0;
    for (auto& b : blocks) {
        const size_t m = b.kind.size();
        if (m) {
            memcpy(&t.id_off[at], b.id_off.data(), m * 8);
            memcpy(&t.id_len[at], b.id_len.data(), m * 4);
            memcpy(&t.kind[at], b.kind.data(), m);
            memcpy(&t.seqs[at * 16], b.seqs.data(), m * 16);
        }
        at += m;
        RowBlock().id_off.swap(b.id_off); std::vector<char>().swap(b.seqs);
    }
    for (size_t i = 0; i < rows; i++) { t.emit_rows += (t.kind[i] & ROW_EMIT) ? 1 : 0; t.barcode_rows += (t.kind[i] & ROW_BARCODE) ? 1 : 0; }
    if (t.barcode_rows == 0) return "no 16/17-character barcode in the file";    // all-NaN / odd columns: leave it to pandas
    return "";
}

constexpr uint64_t NO_CENTRE = 1ull << 32;

// barcode_graph.py:395-410: header readID/barcode, one line per emitted row: id, then the centre (common.py:27-38 unrank)
// or '*'.  centre[i] is per ROW (all rows, emitted or not); values >= 2^32 mean "unassigned".
// Two forms of the per-row result: uint64 (>= 2^32: none) or uint32 + a "has a centre" byte (5 instead of 8 bytes per row).
inline std::string tsv_write(const Tsv& t, const char* out_path, const uint64_t* centre64, int threads, const uint32_t* centre32 = nullptr,
                             const uint8_t* has = nullptr)
{
    auto centre_of = [&](size_t i) -> uint64_t { return centre64 ? centre64[i] : (has[i] ? (uint64_t)centre32[i] : NO_CENTRE); };
    FILE* fo = fopen(out_path, "wb");
    if (!fo) return std::string("cannot create ") + out_path + ": " + strerror(errno);
    static const char HDR[] = "readID\tbarcode\n";
    bool ok = fwrite(HDR, 1, sizeof(HDR) - 1, fo) == sizeof(HDR) - 1;
    const size_t BLOCK = (size_t)1 << 22;                   // rows per buffer fill
    std::vector<char> buf;
    std::vector<size_t> off;
    const char* d = t.file.data;
    for (size_t r0 = 0; r0 < t.rows && ok; r0 += BLOCK) {
        const size_t r1 = std::min(t.rows, r0 + BLOCK), m = r1 - r0;
        const int T = std::max(1, std::min(pick_threads(threads, m * 64), (int)(m / 4096 + 1)));
        std::vector<size_t> part(T + 1, 0);
        auto span = [&](int k) { return std::make_pair(r0 + m * k / T, r0 + m * (k + 1) / T); };
        auto measure = [&](int k) {
            size_t bytes = 0;
            auto [a, b] = span(k);
            for (size_t i = a; i < b; i++)
                if (t.kind[i] & ROW_EMIT) bytes += (size_t)t.id_len[i] + 2 + (centre_of(i) < NO_CENTRE ? 16 : 1);
            part[k + 1] = bytes;
        };
        auto fill = [&](int k) {
            char* p = buf.data() + part[k];
            auto [a, b] = span(k);
            for (size_t i = a; i < b; i++) {
                if (!(t.kind[i] & ROW_EMIT)) continue;
                memcpy(p, d + t.id_off[i], t.id_len[i]);
                p += t.id_len[i];
                *p++ = '\t';
                const uint64_t c = centre_of(i);
                if (c < NO_CENTRE) {
                    for (int j = 0; j < 16; j++) *p++ = "ACGT"[(c >> (2 * j)) & 3];
                } else {
                    *p++ = '*';
                }
                *p++ = '\n';
            }
        };
        auto run = [&](auto&& fn) {
            std::vector<std::thread> pool;
            for (int k = 1; k < T; k++) pool.emplace_back([&, k] { fn(k); });
            fn(0);
            for (auto& th : pool) th.join();
        };
        run(measure);
        for (int k = 0; k < T; k++) part[k + 1] += part[k];
        buf.resize(part[T]);
        run(fill);
        ok = fwrite(buf.data(), 1, buf.size(), fo) == buf.size();
    }
    if (fclose(fo) != 0) ok = false;
    return ok ? "" : std::string("short write to ") + out_path;
}

// badger.py:82-88: `set(open(path).read().split("\n"))`.  Only entries of exactly 16 characters can ever equal an
// unranked barcode (barcode_graph.py:264), so only those are handed on (16 bytes each; letters are checked by pack16).
// Text mode translates \r\n and \r to \n: a carriage return ends an entry as well.
struct Lines16 {
    std::vector<char> seqs;
    size_t count = 0;
};

inline std::string lines16_read(Lines16& L, const char* path, bool* io_error)
{
    *io_error = false;
    Mapped f;
    std::string err;
    if (!f.open_file(path, err)) { *io_error = true; return err; }
    const char* d = f.data;
    const size_t n = f.size;
    const int T = pick_threads(0, n);
    std::vector<size_t> cut(T + 1, n);
    cut[0] = 0;
    for (int k = 1; k < T; k++) {                           // slice borders moved to the next entry start
        size_t c = std::max(cut[k - 1], n / T * k);
        while (c < n && d[c] != '\n' && d[c] != '\r') c++;
        cut[k] = c < n ? c + 1 : n;
    }
    // two sweeps per slice: count the 16-character entries, then copy them to their final place
    std::vector<size_t> found(T + 1, 0);
    auto sweep = [&](int k, char* dst) {
        const size_t beg = cut[k], end = cut[k + 1];
        size_t ls = beg, hits = 0;
        // a slice ends right after a terminator, except the last one, whose final entry may end at the end of the file
        for (size_t i = beg; i < end || (i == end && end == n && k == T - 1); i++) {
            if (i == n || d[i] == '\n' || d[i] == '\r') {
                if (i - ls == 16) {
                    if (dst) memcpy(dst + hits * 16, d + ls, 16);
                    hits++;
                }
                ls = i + 1;
            }
        }
        if (!dst) found[k + 1] = hits;
    };
    auto run = [&](bool copy) {
        std::vector<std::thread> pool;
        for (int k = 1; k < T; k++) pool.emplace_back([&, k] { sweep(k, copy ? L.seqs.data() + found[k] * 16 : nullptr); });
        sweep(0, copy ? L.seqs.data() : nullptr);
        for (auto& th : pool) th.join();
    };
    run(false);
    for (int k = 0; k < T; k++) found[k + 1] += found[k];
    L.count = found[T];
    L.seqs.resize(L.count * 16);
    run(true);
    return "";
}

}  // namespace tsvio
