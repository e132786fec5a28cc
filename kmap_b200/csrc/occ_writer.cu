// Host-side writer of the *.motif_occurence.csv rows (motif_discovery.py:1409-1418 gen_motif_occurence_file, cell format
// of get_motif_occurence :1472-1475).  The occurrence scan itself runs on the device (mask.cu); what is left of the
// reference's per-read loop is turning (read, consensus) -> positions into text, which in Python costs ~5 us per row and
// is what scan_motif then waits for.  No CUDA here: plain buffered formatting of host arrays.
#include <cstdio>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace {

inline char* put_int(char* p, long long v) {          // decimal, no sign handling needed beyond '-'
    if (v < 0) { *p++ = '-'; v = -v; }
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

}  // namespace

extern "C" int64_t kmap_write_occurrence_rows(const char* path_host, int append, int m, const int64_t* const* offsets_host,
                                              const int32_t* const* pos_host, const int64_t* seq_len_host, int64_t r0, int64_t r1) {
    if (!path_host || m < 0 || r0 < 0 || r1 < r0 || (m > 0 && (!offsets_host || !pos_host)) || !seq_len_host) {
        kmap_set_error("write_occurrence_rows: bad argument");
        return KMAP_ERR_BAD_ARG;
    }
    FILE* fh = std::fopen(path_host, append ? "ab" : "wb");
    if (!fh) { kmap_set_error("write_occurrence_rows: cannot open %s", path_host); return KMAP_ERR_BAD_ARG; }
    std::vector<char> buf((size_t)1 << 22);
    char* p = buf.data();
    char* const flush_at = buf.data() + buf.size() / 2;
    int64_t rows = 0;
    bool ok = true;
    for (int64_t r = r0; r < r1 && ok; ++r) {
        int64_t hits = 0, need = 64;
        for (int j = 0; j < m; ++j) { const int64_t c = offsets_host[j][r + 1] - offsets_host[j][r]; hits += c; need += 12 * c + 2; }
        if (hits == 0) continue;                       // only reads with a motif get a row (:1414-1417)
        if (p + need > buf.data() + buf.size() || p > flush_at) {
            ok = std::fwrite(buf.data(), 1, (size_t)(p - buf.data()), fh) == (size_t)(p - buf.data());
            p = buf.data();
            if ((size_t)need > buf.size()) buf.resize((size_t)need * 2), p = buf.data();
        }
        p = put_int(p, r);
        *p++ = ';';
        for (int j = 0; j < m; ++j) {
            const int64_t a = offsets_host[j][r], b = offsets_host[j][r + 1];
            for (int64_t i = a; i < b; ++i) {
                if (i > a) *p++ = ',';
                p = put_int(p, pos_host[j][i]);
            }
            *p++ = ';';
        }
        p = put_int(p, seq_len_host[r]);
        *p++ = '\n';
        ++rows;
    }
    if (ok && p > buf.data()) ok = std::fwrite(buf.data(), 1, (size_t)(p - buf.data()), fh) == (size_t)(p - buf.data());
    if (std::fclose(fh) != 0) ok = false;
    if (!ok) { kmap_set_error("write_occurrence_rows: write to %s failed", path_host); return KMAP_ERR_BAD_ARG; }
    return rows;
}
