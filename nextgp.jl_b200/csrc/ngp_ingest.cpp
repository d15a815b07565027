// ngp_ingest.cpp — genotype ingest of libngp: the reference's text format and PLINK .bed are turned into
// 2-bit codes (NGP_GENO_PACKED2: 4 codes per byte, LSB first, column-major) WITHOUT ever materialising the Float64 matrix
// that /root/reference/src/prepMatVec.jl:116-120 builds (CSV.read -> drop columns with missing -> Matrix{Float64}).
// Host-only code: no CUDA, no sampler arithmetic.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include "../../include/ngp.h"

namespace {

struct FileBuf {
    char* p = nullptr;
    size_t n = 0;
    bool mapped = false;
    ~FileBuf() { if (mapped) munmap(p, n); else free(p); }
    int load(const char* path)
    {
        // the file is mapped, not copied (a C3-sized text file is 120 GB); read() into a buffer only where mmap is not available
        const int fd = open(path, O_RDONLY);
        if (fd >= 0) {
            struct stat st;
            if (fstat(fd, &st) == 0 && st.st_size > 0) {
                void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) {
                    (void)madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                    p = (char*)m; n = (size_t)st.st_size; mapped = true;
                    close(fd);
                    return NGP_OK;
                }
            }
            close(fd);
        }
        FILE* f = fopen(path, "rb");
        if (!f) return NGP_EINVAL;
        if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return NGP_EINVAL; }
        const long sz = ftell(f);
        if (sz < 0) { fclose(f); return NGP_EINVAL; }
        rewind(f);
        p = (char*)malloc((size_t)sz + 1);
        if (!p) { fclose(f); return NGP_ENOMEM; }
        n = fread(p, 1, (size_t)sz, f);
        fclose(f);
        p[n] = 0;
        return NGP_OK;
    }
};

// one field of the reference's text format: -1 missing ("" / NA / NaN / missing), 0..2 a code, -2 anything else (dosage, text)
inline int parse_field(const char* a, const char* b)
{
    while (a < b && (*a == '\t' || *a == '\r')) ++a;
    while (b > a && (b[-1] == '\t' || b[-1] == '\r')) --b;
    const size_t len = (size_t)(b - a);
    if (len == 0) return -1;
    if (len == 1 && *a >= '0' && *a <= '2') return *a - '0';
    if ((len == 2 && !strncasecmp(a, "NA", 2)) || (len == 3 && !strncasecmp(a, "NaN", 3)) || (len == 7 && !strncasecmp(a, "missing", 7))) return -1;
    char tmp[64];
    if (len >= sizeof tmp) return -2;
    memcpy(tmp, a, len);
    tmp[len] = 0;
    char* end = nullptr;
    const double v = strtod(tmp, &end);
    if (end != tmp + len) return -2;
    if (v == 0.0) return 0;
    if (v == 1.0) return 1;
    if (v == 2.0) return 2;
    return (v != v) ? -1 : -2;
}

// keeps the columns with keep[j] != 0, in order, at the front of the packed matrix
int64_t compact_columns(uint8_t* packed, int64_t ld, int64_t p_total, const uint8_t* keep)
{
    int64_t k = 0;
    for (int64_t j = 0; j < p_total; ++j)
        if (keep[j]) {
            if (k != j) memmove(packed + k * ld, packed + j * ld, (size_t)ld);
            ++k;
        }
    return k;
}

// Fast path of the text reader: every row is "c c c ... c" with c in {0,1,2}, single spaces, no missing values (what a genotype file
// of prepMatVec.jl:116-120 looks like unless it holds NA).  Rows are handled in blocks of 256 (= 64 output bytes = one cache line per
// column), one block per task, tasks spread over the host threads: per column the 256 characters come from 256 sequential streams (16 KB
// of lines in L1) and the 64 packed bytes go out as one line.  Returns false (nothing trusted) if any row is not of that form.
bool pack_rows_fast(const std::vector<const char*>& line, const std::vector<uint32_t>& len, int64_t n, int64_t p, uint8_t* packed, int64_t ld)
{
    const int64_t nblk = (n + 255) / 256;
    const size_t want = (size_t)(2 * p - 1);
    for (int64_t i = 0; i < n; ++i) if (len[(size_t)i] != want) return false;
    std::atomic<int64_t> next(0);
    std::atomic<bool> ok(true);
    unsigned T = std::thread::hardware_concurrency();
    T = std::max(1u, std::min<unsigned>(T ? T : 1u, (unsigned)std::min<int64_t>(nblk, 64)));
    auto work = [&]() {
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= nblk || !ok.load(std::memory_order_relaxed)) return;
            const int64_t i0 = b * 256, rows = std::min<int64_t>(256, n - i0);
            const char* L[256];
            for (int64_t r = 0; r < 256; ++r) L[r] = line[(size_t)std::min<int64_t>(i0 + r, n - 1)];
            const int64_t nq = (rows + 3) / 4;
            unsigned bad = 0;
            // separators: every odd position a space
            for (int64_t r = 0; r < rows; ++r) {
                const char* c = L[r];
                unsigned acc = 0;
                for (int64_t k = 1; k < (int64_t)want; k += 2) acc |= (unsigned)(c[k] ^ ' ');
                bad |= acc;
            }
            uint8_t* out0 = packed + (i0 >> 2);
            for (int64_t j = 0; j < p; ++j) {
                uint8_t* out = out0 + j * ld;
                const int64_t at = 2 * j;
                for (int64_t q = 0; q < nq; ++q) {
                    unsigned byte = 0;
                    for (int64_t u = 0; u < 4; ++u) {
                        const int64_t r = 4 * q + u;
                        if (r < rows) {
                            const unsigned g = (unsigned)(unsigned char)L[r][at] - (unsigned)'0';
                            bad |= (g > 2u);
                            byte |= (g & 3u) << (2 * u);
                        }
                    }
                    out[q] = (uint8_t)byte;
                }
            }
            if (bad) ok.store(false);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return ok.load();
}

}  // namespace

extern "C" {

int ngp_read_text_genotypes(const char* path, int64_t* n_out, int64_t* p_total_out, uint8_t* packed, int64_t ld, uint8_t* keep, int64_t* p_kept)
{
    if (!path || !n_out || !p_total_out) return NGP_EINVAL;
    FileBuf fb;
    int rc = fb.load(path);
    if (rc) return rc;
    const char* s = fb.p;
    const char* const end = fb.p + fb.n;
    // pass 1: rows (non-empty lines) and fields of the first row; delim = ' ' exactly as CSV.read(...; delim=' ') sees it
    int64_t n = 0, p = 0;
    std::vector<const char*> line;                      // non-empty rows: start and length (without the line end)
    std::vector<uint32_t> llen;
    bool fast_ok = true;
    for (const char* q = s; q < end;) {
        const char* e = (const char*)memchr(q, '\n', (size_t)(end - q));
        if (!e) e = end;
        const char* le = e;
        if (le > q && le[-1] == '\r') --le;
        if (le > q) {
            if (n == 0) { p = 1; for (const char* c = q; c < le; ++c) if (*c == ' ') ++p; }
            ++n;
            if (packed) { line.push_back(q); if ((size_t)(le - q) > 0xffffffffull) fast_ok = false; llen.push_back((uint32_t)(le - q)); }
        }
        q = e + 1;
    }
    *n_out = n; *p_total_out = p;
    if (!packed) return NGP_OK;                         // sizing call
    if (!keep || !p_kept || n <= 0 || p <= 0 || ld < (n + 3) / 4) return NGP_EINVAL;
    memset(keep, 1, (size_t)p);
    if (fast_ok && pack_rows_fast(line, llen, n, p, packed, ld)) {      // all codes, no missing value: nothing to drop
        // (bytes of a column beyond ceil(n/4) are the caller's padding: zero them like the general path does)
        const int64_t used = (n + 3) / 4;
        if (ld > used) for (int64_t j = 0; j < p; ++j) memset(packed + j * ld + used, 0, (size_t)(ld - used));
        *p_kept = p;
        return NGP_OK;
    }
    memset(packed, 0, (size_t)ld * (size_t)p);
    int64_t i = 0;
    for (const char* q = s; q < end;) {
        const char* e = (const char*)memchr(q, '\n', (size_t)(end - q));
        if (!e) e = end;
        const char* le = e;
        if (le > q && le[-1] == '\r') --le;
        if (le > q) {
            int64_t j = 0;
            const char* a = q;
            for (;;) {
                const char* b = (const char*)memchr(a, ' ', (size_t)(le - a));
                if (!b) b = le;
                if (j >= p) return NGP_EDATA;            // ragged row: more fields than the first row
                const int g = parse_field(a, b);
                if (g == -2) return NGP_EDATA;           // dosage / text: packed storage holds 0/1/2 only
                if (g < 0) keep[j] = 0;                  // a missing value drops the column (prepMatVec.jl:118)
                else packed[j * ld + (i >> 2)] |= (uint8_t)(g << (2 * (i & 3)));
                ++j;
                if (b == le) break;
                a = b + 1;
            }
            if (j != p) return NGP_EDATA;                // ragged row: fewer fields
            ++i;
        }
        q = e + 1;
    }
    *p_kept = compact_columns(packed, ld, p, keep);
    return NGP_OK;
}

int ngp_read_bed_genotypes(const char* path, int64_t n, int64_t p, int count_a1, uint8_t* packed, int64_t ld, uint8_t* keep, int64_t* p_kept)
{
    if (!path || !packed || !keep || !p_kept || n <= 0 || p <= 0 || ld < (n + 3) / 4) return NGP_EINVAL;
    FILE* f = fopen(path, "rb");
    if (!f) return NGP_EINVAL;
    unsigned char magic[3];
    if (fread(magic, 1, 3, f) != 3 || magic[0] != 0x6c || magic[1] != 0x1b) { fclose(f); return NGP_EDATA; }
    if (magic[2] != 0x01) { fclose(f); return NGP_EUNSUPPORTED; }     // individual-major .bed files are obsolete
    // PLINK: 00 hom A1, 01 missing, 10 het, 11 hom A2 (sample i in bits 2(i&3) of byte i>>2: the same positions as NGP_GENO_PACKED2)
    uint8_t lut[256], miss[256];
    for (int v = 0; v < 256; ++v) {
        uint8_t o = 0, m = 0;
        for (int k = 0; k < 4; ++k) {
            const int c = (v >> (2 * k)) & 3;
            int g = 0;
            if (c == 1) m = 1;
            else if (c == 0) g = count_a1 ? 2 : 0;
            else if (c == 2) g = 1;
            else g = count_a1 ? 0 : 2;
            o |= (uint8_t)(g << (2 * k));
        }
        lut[v] = o; miss[v] = m;
    }
    const int64_t bpc = (n + 3) / 4;
    const int tail = (int)(n & 3);                       // samples in the last byte (0 = full): padding bits are ignored
    fclose(f);
    // variant-major .bed = one column of NGP_GENO_PACKED2 after the other: the file is mapped and the columns are recoded in parallel
    FileBuf fb;
    int rc = fb.load(path);
    if (rc) return rc;
    if (fb.n < 3 + (size_t)bpc * (size_t)p) return NGP_EDATA;           // truncated file
    const uint8_t* src = reinterpret_cast<const uint8_t*>(fb.p) + 3;
    std::atomic<int64_t> next(0);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(4096, (1 << 20) / bpc + 1));      // ~1 MB of columns per task
    unsigned T = std::thread::hardware_concurrency();
    T = std::max(1u, std::min<unsigned>(T ? T : 1u, (unsigned)std::min<int64_t>((p + chunk - 1) / chunk, 64)));
    auto work = [&]() {
        for (;;) {
            const int64_t j0 = next.fetch_add(chunk);
            if (j0 >= p) return;
            for (int64_t j = j0; j < std::min(p, j0 + chunk); ++j) {
                const uint8_t* col = src + j * bpc;
                uint8_t* o = packed + j * ld;
                uint8_t m = 0;
                for (int64_t b = 0; b < bpc - 1; ++b) { o[b] = lut[col[b]]; m |= miss[col[b]]; }
                uint8_t last = col[bpc - 1];
                if (tail) last |= (uint8_t)(0xff << (2 * tail));            // pad = 0b11: never "missing"
                o[bpc - 1] = lut[last]; m |= miss[last];
                if (tail) o[bpc - 1] &= (uint8_t)~(0xff << (2 * tail));     // pad rows hold code 0
                if (ld > bpc) memset(o + bpc, 0, (size_t)(ld - bpc));
                keep[j] = m ? 0 : 1;
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    *p_kept = compact_columns(packed, ld, p, keep);
    return NGP_OK;
}

}  // extern "C"
