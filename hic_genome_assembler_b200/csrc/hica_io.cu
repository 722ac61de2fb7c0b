// hica_io.cu - host-side loader of the HiC-Pro sparse matrix text.  No device code: the file is
// compiled by nvcc only so that it lands in libhica.so next to the kernels that consume its output.
//
// Replaces buildAdjacencyMatrix (scaffoldToChromosomes.py:70-98): the reference reads "id1\tid2\tvalue" lines with
// Python's int()/float() into a list of lists (minutes at 16k bins, 8 GB of Python floats).  Here the file is
// mmap-ed, cut at line boundaries into one chunk per thread, parsed in parallel into (row, column, value)
// triplets and scattered into the caller's dense fp64 matrix (which may be the pinned staging buffer the upload
// reads from).  Semantics kept from the reference:
//   * both [i][j] and [j][i] are written (:90-91); lines naming a bin that is not in the index are skipped (:86-89);
//   * when a cell is written more than once the LAST line in file order wins: every destination row is owned by
//     exactly one scatter thread, and that thread walks the chunks in file order;
//   * "line.strip('\r').strip('\n').split('\t')" (:84): columns beyond the third are ignored, blanks around a
//     number are accepted (int()/float() strip them), the value goes through a correctly rounded decimal ->
//     binary64 conversion (std::from_chars / strtod), the same result as Python's float();
//   * a malformed line makes the reference raise ValueError/IndexError; here the call fails with the line number.
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <math.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "hica_common.cuh"

namespace {

struct Triplet {
    int32_t i, j;
    double v;
};

struct ChunkResult {
    std::vector<Triplet> t;
    int64_t lines = 0, applied = 0;
    int64_t bad_line = -1;  // 0-based line index inside the chunk
    std::string why;
};

inline bool is_blank(char c) { return c == ' ' || c == '\r' || c == '\v' || c == '\f'; }

// int(): optional blanks, optional sign, decimal digits (Python also accepts '_' separators; not produced by HiC-Pro)
bool parse_int(const char *b, const char *e, int64_t *out)
{
    while (b < e && is_blank(*b)) b++;
    while (e > b && is_blank(e[-1])) e--;
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; b++; }
    if (b == e) return false;
    int64_t v = 0;
    for (; b < e; b++) {
        if (*b < '0' || *b > '9') return false;
        if (v > (INT64_MAX - 9) / 10) return false;
        v = v * 10 + (*b - '0');
    }
    *out = neg ? -v : v;
    return true;
}

// float(): blanks, sign, decimal / exponent forms, inf / nan
bool parse_float(const char *b, const char *e, double *out)
{
    while (b < e && is_blank(*b)) b++;
    while (e > b && is_blank(e[-1])) e--;
    if (b == e) return false;
    const char *p = b;
    if (*p == '+') {  // from_chars does not accept a leading '+'
        p++;
        if (p == e || *p == '+' || *p == '-') return false;
    }
    auto r = std::from_chars(p, e, *out, std::chars_format::general);
    if (r.ec == std::errc() && r.ptr == e) return true;
    // out-of-range values (Python gives inf / 0.0) and spellings from_chars rejects: strtod on a copy
    std::string s(b, e);
    if (s.find_first_of("xXpP") != std::string::npos) return false;  // float() has no hex literals
    char *end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end != s.c_str() + s.size()) return false;
    *out = v;
    return true;
}

void parse_chunk(const char *b, const char *e, const int32_t *index_of_id, int64_t max_id, ChunkResult *res)
{
    res->t.reserve((size_t)((e - b) / 16 + 16));
    const char *p = b;
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = nl ? nl : e;
        // three tab-separated columns; anything after a third tab is ignored
        const char *t1 = (const char *)memchr(p, '\t', (size_t)(le - p));
        const char *t2 = t1 ? (const char *)memchr(t1 + 1, '\t', (size_t)(le - t1 - 1)) : nullptr;
        if (!t1 || !t2) {
            res->bad_line = res->lines;
            res->why = "fewer than three tab-separated columns";
            return;
        }
        const char *t3 = (const char *)memchr(t2 + 1, '\t', (size_t)(le - t2 - 1));
        const char *ve = t3 ? t3 : le;
        int64_t id1, id2;
        double v;
        if (!parse_int(p, t1, &id1) || !parse_int(t1 + 1, t2, &id2)) {
            res->bad_line = res->lines;
            res->why = "bin id is not an integer";
            return;
        }
        if (!parse_float(t2 + 1, ve, &v)) {
            res->bad_line = res->lines;
            res->why = "value is not a number";
            return;
        }
        res->lines++;
        if (id1 >= 0 && id1 <= max_id && id2 >= 0 && id2 <= max_id) {
            const int32_t i = index_of_id[id1], j = index_of_id[id2];
            if (i >= 0 && j >= 0) {
                res->t.push_back({i, j, v});
                res->applied++;
            }
        }
        p = nl ? nl + 1 : e;
    }
}

std::string g_io_error;

int io_fail(const char *fmt, const char *a, long long b, const char *c)
{
    char buf[1024];
    snprintf(buf, sizeof(buf), fmt, a, b, c);
    g_io_error = buf;
    return HICA_ERR_ARG;
}

}  // namespace

extern "C" const char *hica_io_last_error(void) { return g_io_error.c_str(); }

extern "C" int hica_io_read_matrix(const char *path, const int32_t *index_of_id, int64_t max_id, int64_t n,
                                   double *A_host, int threads, int64_t *n_lines, int64_t *n_edges)
{
    if (!path || !index_of_id || !A_host || n < 0 || max_id < 0) return io_fail("%s: bad argument%.0lld%s", "hica_io_read_matrix", 0, "");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return io_fail("cannot open %s%.0lld%s", path, 0, "");
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return io_fail("cannot stat %s%.0lld%s", path, 0, ""); }
    const int64_t size = (int64_t)st.st_size;
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, 256));
    // zero fill in parallel (first touch by the thread that later owns the rows)
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([=] {
                const int64_t r0 = n * t / T, r1 = n * (t + 1) / T;
                if (r1 > r0) memset(A_host + r0 * n, 0, sizeof(double) * (size_t)((r1 - r0) * n));
            });
        for (auto &x : th) x.join();
    }
    if (n_lines) *n_lines = 0;
    if (n_edges) *n_edges = 0;
    if (size == 0) { close(fd); return HICA_OK; }
    const char *data = (const char *)mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (data == MAP_FAILED) return io_fail("cannot mmap %s%.0lld%s", path, 0, "");
    madvise((void *)data, (size_t)size, MADV_SEQUENTIAL);
    // chunk boundaries at line starts
    const int C = (int)std::max<int64_t>(1, std::min<int64_t>(T, size / 4096 + 1));
    std::vector<int64_t> cut(C + 1);
    cut[0] = 0;
    cut[C] = size;
    for (int c = 1; c < C; c++) {
        int64_t p = std::max(cut[c - 1], size * c / C);
        const char *nl = p < size ? (const char *)memchr(data + p, '\n', (size_t)(size - p)) : nullptr;
        cut[c] = nl ? (nl - data) + 1 : size;
    }
    std::vector<ChunkResult> res(C);
    {
        std::vector<std::thread> th;
        for (int c = 0; c < C; c++)
            th.emplace_back([&, c] { parse_chunk(data + cut[c], data + cut[c + 1], index_of_id, max_id, &res[c]); });
        for (auto &x : th) x.join();
    }
    munmap((void *)data, (size_t)size);
    int64_t lines = 0, edges = 0;
    for (int c = 0; c < C; c++) {
        if (res[c].bad_line >= 0) return io_fail("%s: line %lld: %s", path, (long long)(lines + res[c].bad_line + 1), res[c].why.c_str());
        lines += res[c].lines;
        edges += res[c].applied;
    }
    // scatter: thread t owns destination rows [n t / T, n (t+1) / T) and walks all chunks in file order
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                const int64_t r0 = n * t / T, r1 = n * (t + 1) / T;
                if (r1 <= r0) return;
                for (int c = 0; c < C; c++)
                    for (const Triplet &e : res[c].t) {
                        if (e.i >= r0 && e.i < r1) A_host[(int64_t)e.i * n + e.j] = e.v;
                        if (e.j >= r0 && e.j < r1) A_host[(int64_t)e.j * n + e.i] = e.v;
                    }
            });
        for (auto &x : th) x.join();
    }
    if (n_lines) *n_lines = lines;
    if (n_edges) *n_edges = edges;
    return HICA_OK;
}

// logTransformMatrix (scaffoldToChromosomes.py:165-183, orderGenome.py:160-178): an interpreted double loop over
// all N^2 entries in the reference (minutes at 16k bins, and an 8 GB list of lists) that sits on the critical path
// of the unchanged entry point although it only feeds the plots.  Same arithmetic, same libm:
//   forward  v != 0 -> math.log(float(v) + off, base) = log(v + off) / log(base)    (CPython: m_log(x) / m_log(base))
//   reverse  v != 0 -> (float(base) ** v) - off       = pow(base, v) - off
// with off = 1. in scaffoldToChromosomes.py and off = 0. (no offset at all) in orderGenome.py.
// A forward argument v + off <= 0 makes math.log raise ValueError in the reference: the call fails.
extern "C" int hica_io_log_transform(const double *in, double *out, int64_t count, double base, double offset,
                                     int reverse, int threads)
{
    if (!in || !out || count < 0) return io_fail("%s: bad argument%.0lld%s", "hica_io_log_transform", 0, "");
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min<int>(T, (int)std::min<int64_t>(256, count / 65536 + 1)));
    const double lb = log(base);
    std::atomic<int64_t> bad(-1);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            const int64_t i0 = count * t / T, i1 = count * (t + 1) / T;
            for (int64_t i = i0; i < i1; i++) {
                const double v = in[i];
                double r = 0.0;
                if (v != 0.) {
                    if (!reverse) {
                        const double x = v + offset;
                        if (!(x > 0.) && !(x != x)) { bad.store(i); r = 0.0; }
                        else r = log(x) / lb;
                    } else {
                        r = pow(base, v) - offset;
                    }
                }
                out[i] = r;
            }
        });
    for (auto &x : th) x.join();
    if (bad.load() >= 0) return io_fail("%s: math domain error at entry %lld%s", "logTransformMatrix", (long long)bad.load(), "");
    return HICA_OK;
}

// initiateLoci (scaffoldToChromosomes.py:35-68, orderGenome.py:30-63): the HiC-Pro bed file ("chrom<TAB>start<TAB>
// stop<TAB>binID") read in step with the bias file (one value per line) -> the list of genomic bins.  The reference
// builds one Python object per line; here both files are read whole, walked once and returned as columns (ids,
// starts, stops, biases, chromosome names as one byte string + offsets).  Semantics kept from the reference:
//   * text mode with universal newlines: "\n", "\r\n" and a lone "\r" all end a line; the last line may be open;
//   * a bias line is consumed for every bed line, also for bins the filter skips (:45-46), and a bias file that ends
//     early yields '' -> float('') fails -> bias 0.0 (:58-61);
//   * bed columns go through int() (:50) - a line with fewer than four columns or a non-integer field makes the
//     reference raise (IndexError / ValueError) before the filter is looked at: the call fails with the line number;
//   * binID_dict (:53-55): only ids in the set are kept;
//   * a bias that is exactly the string "nan" drops the bin (:57); anything else float() cannot parse becomes 0.0,
//     and spellings float() does parse to a NaN ("NaN", " nan", "+nan") keep the bin with a NaN bias.
namespace {

struct LociResult {
    std::vector<int64_t> id, start, stop, chrom_off;
    std::vector<double> bias;
    std::string chrom;
};
thread_local LociResult g_loci;

bool read_whole(const char *path, std::string *out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, got);
    const bool ok = !ferror(f);
    fclose(f);
    return ok;
}

// next line of [*p, e) under universal newlines; false at end of data
bool next_line(const char **p, const char *e, const char **lb, const char **le)
{
    if (*p >= e) return false;
    const char *q = *p;
    *lb = q;
    while (q < e && *q != '\n' && *q != '\r') q++;
    *le = q;
    if (q < e) q += (*q == '\r' && q + 1 < e && q[1] == '\n') ? 2 : 1;
    *p = q;
    return true;
}

inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

}  // namespace

extern "C" int hica_io_read_loci(const char *bed_path, const char *bias_path, const int64_t *keep_ids_sorted,
                                 int64_t n_keep, int64_t *n_bins, int64_t *chrom_bytes)
{
    if (!bed_path || !bias_path || !n_bins || !chrom_bytes || (n_keep > 0 && !keep_ids_sorted))
        return io_fail("%s: bad argument%.0lld%s", "hica_io_read_loci", 0, "");
    LociResult &R = g_loci;
    R = LociResult();
    std::string bed, bias;
    if (!read_whole(bed_path, &bed)) return io_fail("cannot open %s%.0lld%s", bed_path, 0, "");
    if (!read_whole(bias_path, &bias)) return io_fail("cannot open %s%.0lld%s", bias_path, 0, "");
    const char *bp = bed.data(), *be = bp + bed.size(), *ip = bias.data(), *ie = ip + bias.size();
    const char *lb, *le, *vb, *ve;
    int64_t line = 0;
    R.chrom_off.push_back(0);
    while (next_line(&bp, be, &lb, &le)) {
        line++;
        const bool have_bias = next_line(&ip, ie, &vb, &ve);
        const char *col[5];
        int nc = 0;
        col[nc++] = lb;
        for (const char *q = lb; q < le && nc < 5; q++)
            if (*q == '\t') col[nc++] = q + 1;
        if (nc < 4) return io_fail("%s: line %lld: %s", bed_path, (long long)line, "fewer than four tab-separated columns");
        const char *end3 = nc == 5 ? col[4] - 1 : le;
        int64_t start, stop, id;
        auto field = [&](const char *b, const char *e, int64_t *out) {
            while (b < e && is_space(*b)) b++;
            while (e > b && is_space(e[-1])) e--;
            return parse_int(b, e, out);
        };
        if (!field(col[1], col[2] - 1, &start) || !field(col[2], col[3] - 1, &stop) || !field(col[3], end3, &id))
            return io_fail("%s: line %lld: %s", bed_path, (long long)line, "start / stop / bin id is not an integer");
        if (n_keep >= 0 && !std::binary_search(keep_ids_sorted, keep_ids_sorted + n_keep, id)) continue;
        double bv = 0.0;
        if (have_bias) {
            if (ve - vb == 3 && memcmp(vb, "nan", 3) == 0) continue;
            const char *b = vb, *e = ve;
            while (b < e && is_space(*b)) b++;
            while (e > b && is_space(e[-1])) e--;
            if (b == e || memchr(b, '(', (size_t)(e - b)) || !parse_float(b, e, &bv)) bv = 0.0;
        }
        R.id.push_back(id);
        R.start.push_back(start);
        R.stop.push_back(stop);
        R.bias.push_back(bv);
        R.chrom.append(col[0], (size_t)(col[1] - 1 - col[0]));
        R.chrom_off.push_back((int64_t)R.chrom.size());
    }
    *n_bins = (int64_t)R.id.size();
    *chrom_bytes = (int64_t)R.chrom.size();
    return HICA_OK;
}

extern "C" int hica_io_loci_fetch(int64_t *ids, int64_t *starts, int64_t *stops, double *bias, int64_t *chrom_off,
                                  char *chrom_buf)
{
    LociResult &R = g_loci;
    if (R.chrom_off.empty()) return io_fail("%s: no result to fetch (call hica_io_read_loci first)%.0lld%s", "hica_io_loci_fetch", 0, "");
    const size_t n = R.id.size();
    if (n && (!ids || !starts || !stops || !bias)) return io_fail("%s: bad argument%.0lld%s", "hica_io_loci_fetch", 0, "");
    if (!chrom_off || (!chrom_buf && !R.chrom.empty())) return io_fail("%s: bad argument%.0lld%s", "hica_io_loci_fetch", 0, "");
    if (n) {
        memcpy(ids, R.id.data(), sizeof(int64_t) * n);
        memcpy(starts, R.start.data(), sizeof(int64_t) * n);
        memcpy(stops, R.stop.data(), sizeof(int64_t) * n);
        memcpy(bias, R.bias.data(), sizeof(double) * n);
    }
    memcpy(chrom_off, R.chrom_off.data(), sizeof(int64_t) * (n + 1));
    if (!R.chrom.empty()) memcpy(chrom_buf, R.chrom.data(), R.chrom.size());
    R = LociResult();
    return HICA_OK;
}
