// Host-side writer of the *.motif_occurence.csv rows (motif_discovery.py:1409-1418 gen_motif_occurence_file, cell format
// of get_motif_occurence :1472-1475).  The occurrence scan itself runs on the device (mask.cu); what is left of the
// reference's per-read loop is turning (read, consensus) -> positions into text, which in Python costs ~5 us per row and
// is what scan_motif then waits for.  No CUDA here: plain buffered formatting of host arrays.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>
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

namespace {

// rows of reads [r0, r1) appended to `out`; returns the number of rows
int64_t format_rows(int m, const int64_t* const* offsets, const int32_t* const* pos, const int64_t* seq_len, int64_t r0, int64_t r1,
                    std::vector<char>& out) {
    int64_t rows = 0;
    size_t used = out.size();
    for (int64_t r = r0; r < r1; ++r) {
        int64_t hits = 0, need = 64;
        for (int j = 0; j < m; ++j) { const int64_t c = offsets[j][r + 1] - offsets[j][r]; hits += c; need += 12 * c + 2; }
        if (hits == 0) continue;                       // only reads with a motif get a row (:1414-1417)
        if (out.size() < used + (size_t)need) out.resize(std::max(out.size() * 2, used + (size_t)need + 4096));
        char* p = out.data() + used;
        p = put_int(p, r);
        *p++ = ';';
        for (int j = 0; j < m; ++j) {
            const int64_t a = offsets[j][r], b = offsets[j][r + 1];
            for (int64_t i = a; i < b; ++i) {
                if (i > a) *p++ = ',';
                p = put_int(p, pos[j][i]);
            }
            *p++ = ';';
        }
        p = put_int(p, seq_len[r]);
        *p++ = '\n';
        used = (size_t)(p - out.data());
        ++rows;
    }
    out.resize(used);
    return rows;
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
    // batches of reads, each formatted by up to 16 threads into their own buffers and written in read order
    const int64_t batch = (int64_t)1 << 22;
    unsigned hw = std::thread::hardware_concurrency();
    const int n_threads = (int)std::min<unsigned>(hw ? hw : 1u, 16u);
    std::vector<std::vector<char>> bufs((size_t)n_threads);
    std::vector<int64_t> counts((size_t)n_threads);
    int64_t rows = 0;
    bool ok = true;
    for (int64_t b0 = r0; b0 < r1 && ok; b0 += batch) {
        const int64_t b1 = std::min(b0 + batch, r1);
        const int nt = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, (b1 - b0) / 65536));
        const int64_t per = (b1 - b0 + nt - 1) / nt;
        std::vector<std::thread> workers;
        for (int t = 0; t < nt; ++t) {
            bufs[(size_t)t].clear();
            const int64_t a = std::min(b0 + t * per, b1), e = std::min(a + per, b1);
            auto job = [&, t, a, e]() { counts[(size_t)t] = format_rows(m, offsets_host, pos_host, seq_len_host, a, e, bufs[(size_t)t]); };
            if (nt == 1) job(); else workers.emplace_back(job);
        }
        for (auto& w : workers) w.join();
        for (int t = 0; t < nt && ok; ++t) {
            rows += counts[(size_t)t];
            const std::vector<char>& v = bufs[(size_t)t];
            if (!v.empty()) ok = std::fwrite(v.data(), 1, v.size(), fh) == v.size();
        }
    }
    if (std::fclose(fh) != 0) ok = false;
    if (!ok) { kmap_set_error("write_occurrence_rows: write to %s failed", path_host); return KMAP_ERR_BAD_ARG; }
    return rows;
}
