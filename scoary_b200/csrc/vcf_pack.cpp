// vcf_pack.cpp -- SURVEY.md 8(f) rank 3: VCF 4.x -> packed bitset rows, natively.
//
// The reference converts a VCF to its CSV table first (scoary/vcf2scoary.py:50-218) and then
// parses that CSV cell by cell (scoary/methods.py:445-497).  Here the variant lines go straight
// to the uint64 rows sb_set_genes takes, with the converter's rules:
//   * one output row per ALT allele (vcf2scoary.py:178-192);
//   * genotype = the sample cell up to its first ':' (:200-202);
//   * a line with a single ALT keeps the genotype text, and the table parser then reads it as
//     present unless it is "", "0" or "-" (methods.py:476-487);
//   * a line with several ALT alleles gives, for allele c = 1, 2, ...: "." -> 0, int(genotype) == c
//     -> 1, anything else -> 0 (fixdummy, vcf2scoary.py:204-218);
//   * --types keeps a line only if the first "TYPE=<word>" of its INFO field names a wanted type
//     (re.search(r"TYPE=(\w+)"), vcf2scoary.py:170-176).
// Dialect: csv.reader(delimiter="\t", quotechar='"') as the reference opens the file
// (vcf2scoary.py:103-104).  The host only takes this path for buffers without '"' and without
// bytes >= 0x80 (quoting / Unicode word characters then cannot matter); every line this code is not
// sure about (field count differs from the header, a genotype of a multi-allelic line that is not
// "." or plain digits) is reported back and the host falls back to the Python parser for the file.
// Host code only (no CUDA); lines are parsed in parallel with OpenMP.
#include <stdint.h>
#include <string.h>

#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

inline bool is_eol(char c) { return c == '\n' || c == '\r'; }
inline bool is_word(unsigned char c)
{
    return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
}

// end of the line starting at p (position of the terminator or e)
inline const char *line_end(const char *p, const char *e)
{
    // the next '\n', unless a '\r' comes first (memchr twice: the second one only looks at this line)
    const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
    const char *le = nl ? nl : e;
    const char *cr = (const char *)memchr(p, '\r', (size_t)(le - p));
    return cr ? cr : le;
}

inline const char *field_end(const char *p, const char *le)
{
    const void *t = memchr(p, '\t', (size_t)(le - p));
    return t ? (const char *)t : le;
}

// rows a line produces: 0 (filtered out), 1, or the number of ALT alleles.  *alt_b/e: ALT field.
// Returns -1 if the line has fewer than 9 fields.
inline int line_rows(const char *p, const char *le, const char *types, int64_t types_len, const char **alt_b,
                     const char **alt_e, const char **samples)
{
    const char *f = p;
    const char *fb[9], *fe[9];
    for (int k = 0; k < 9; ++k) {
        if (f > le) return -1;
        fb[k] = f;
        fe[k] = field_end(f, le);
        f = fe[k] + 1;
    }
    *alt_b = fb[4];
    *alt_e = fe[4];
    *samples = f;                       // may be le + 1 when there are no sample columns
    if (types) {                        // first "TYPE=" followed by a word character
        const char *q = fb[7];
        const char *hit = nullptr, *hit_e = nullptr;
        while (q + 5 <= fe[7]) {
            if (memcmp(q, "TYPE=", 5) == 0 && q + 5 < fe[7] && is_word((unsigned char)q[5])) {
                hit = q + 5;
                hit_e = hit;
                while (hit_e < fe[7] && is_word((unsigned char)*hit_e)) ++hit_e;
                break;
            }
            ++q;
        }
        if (!hit) return 0;
        bool wanted = false;            // types = '\n'-separated list
        const char *t = types, *te = types + types_len;
        while (t <= te) {
            const char *u = t;
            while (u < te && *u != '\n') ++u;
            if (u - t == hit_e - hit && memcmp(t, hit, (size_t)(u - t)) == 0) { wanted = true; break; }
            t = u + 1;
        }
        if (!wanted) return 0;
    }
    int n_alt = 1;
    for (const char *q = fb[4]; q < fe[4]; ++q) n_alt += (*q == ',');
    return n_alt;
}

}  // namespace

extern "C" {

// Byte offsets of the non-empty lines that do not start with "##": the first one is the
// "#CHROM ..." header, the rest are variant lines.  line_starts may be NULL to count.
// *needs_python is set when the header or a variant line holds a '"' or a byte >= 0x80.
int64_t sb_vcf_line_starts(const char *buf, int64_t len, int64_t *line_starts, int64_t max_lines, int32_t *needs_python)
{
    if (!buf || len < 0) return -1;
    const char *p = buf, *e = buf + len;
    int64_t n = 0, first = len;
    bool odd = false;
    while (p < e) {
        const char *le = line_end(p, e);
        // "##" lines are metainformation only before the header (the converter does not look for them afterwards)
        if (le > p && !(n == 0 && le - p >= 2 && p[0] == '#' && p[1] == '#')) {
            if (line_starts) {
                if (n >= max_lines) return -1;
                line_starts[n] = p - buf;
            }
            if (n == 0) first = p - buf;
            ++n;
        }
        p = le;
        if (p < e) p += (*p == '\r' && p + 1 < e && p[1] == '\n') ? 2 : 1;
    }
    if (needs_python) {
        // the "##" block is parsed by the host's csv module (Description="..." is normal there)
        const int64_t block = 1 << 16, n_blocks = (len - first + block - 1) / block;
        int any = 0;
#pragma omp parallel for schedule(static) reduction(| : any)
        for (int64_t b = 0; b < n_blocks; ++b) {
            const unsigned char *q = (const unsigned char *)buf + first + b * block;
            const int64_t m = std::min<int64_t>(block, len - first - b * block);
            unsigned acc = 0;
            for (int64_t i = 0; i < m; ++i) acc |= (unsigned)(q[i] == '"') | (unsigned)(q[i] >> 7);
            any |= (int)acc;
        }
        odd = any != 0;
        *needs_python = odd ? 1 : 0;
    }
    return n;
}

// Output rows per variant line (0 when --types filters the line out).  types: '\n'-separated
// wanted TYPE values, or NULL for ALL.  Returns the total, or -(line + 1) for a line with fewer
// than nine fields.
int64_t sb_vcf_count_rows(const char *buf, int64_t len, const int64_t *line_starts, int64_t n_lines, const char *types,
                          int64_t types_len, int32_t *rows_per_line)
{
    if (!buf || !line_starts || !rows_per_line) return -1;
    const char *e = buf + len;
    int64_t bad = 0, total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < n_lines; ++i) {
        const char *p = buf + line_starts[i];
        const char *ab, *ae, *sm;
        const int r = line_rows(p, line_end(p, e), types, types_len, &ab, &ae, &sm);
        rows_per_line[i] = r < 0 ? 0 : r;
        if (r < 0) {
#pragma omp critical(sb_vcf_bad)
            if (bad == 0 || i + 1 < bad) bad = i + 1;
        }
        total += r < 0 ? 0 : r;
    }
    return bad ? -bad : total;
}

// Pack the rows.  row_offset[i] = first output row of line i (exclusive prefix sum of rows_per_line).
//   keep     [n_samples] bit index of each sample column, or -1 to drop it (-r)
//   bits     uint64 [n_rows][W], zero-filled here
//   ranges   int64 [n_rows][3][2]: byte ranges of CHROM, POS, ID of each row's line
//   line_of  int64 [n_rows]: the variant line each row came from
// Returns 0, or -(line + 1) for the first line the host has to hand to the Python parser (wrong
// number of sample columns, or an unusual genotype in a multi-allelic line).
int64_t sb_vcf_pack_rows(const char *buf, int64_t len, const int64_t *line_starts, int64_t n_lines,
                         const int32_t *rows_per_line, const int64_t *row_offset, const char *types, int64_t types_len,
                         const int32_t *keep, int32_t n_samples, uint64_t *bits, int32_t W, int64_t *ranges,
                         int64_t *line_of)
{
    if (!buf || !line_starts || !rows_per_line || !row_offset || !bits || (n_samples > 0 && !keep)) return -1;
    const char *e = buf + len;
    int64_t bad = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_lines; ++i) {
        const int rows = rows_per_line[i];
        if (rows == 0) continue;
        const char *p = buf + line_starts[i];
        const char *le = line_end(p, e);
        const char *ab, *ae, *sm;
        const int r = line_rows(p, le, types, types_len, &ab, &ae, &sm);
        bool ok = (r == rows);
        const int64_t row0 = row_offset[i];
        if (ok) {
            for (int a = 0; a < rows; ++a) {
                memset(bits + (row0 + a) * (int64_t)W, 0, sizeof(uint64_t) * (size_t)W);
                const char *f = p;
                for (int k = 0; k < 3; ++k) {
                    const char *fe = field_end(f, le);
                    if (ranges) {
                        ranges[((row0 + a) * 3 + k) * 2 + 0] = f - buf;
                        ranges[((row0 + a) * 3 + k) * 2 + 1] = fe - buf;
                    }
                    f = fe + 1;
                }
                if (line_of) line_of[row0 + a] = i;
            }
            const bool split = rows > 1 || memchr(ab, ',', (size_t)(ae - ab)) != nullptr;
            const char *f = sm;
            int32_t col = 0;
            for (; f <= le && col < n_samples; ++col) {
                const char *fe = field_end(f, le);
                const char *ge = (const char *)memchr(f, ':', (size_t)(fe - f));
                if (!ge) ge = fe;
                const int32_t bi = keep[col];
                if (!split) {
                    const int64_t n = ge - f;
                    const bool absent = n == 0 || (n == 1 && (f[0] == '0' || f[0] == '-'));
                    if (!absent && bi >= 0) bits[row0 * (int64_t)W + (bi >> 6)] |= 1ULL << (bi & 63);
                } else if (!(ge - f == 1 && f[0] == '.')) {
                    // int(genotype): plain digits only here, anything else goes back to Python
                    int64_t v = 0;
                    bool digits = ge > f;
                    for (const char *q = f; q < ge && digits; ++q) {
                        digits = (*q >= '0' && *q <= '9');
                        v = v < (1LL << 40) ? v * 10 + (*q - '0') : v;
                    }
                    if (!digits) { ok = false; break; }
                    if (v >= 1 && v <= rows && bi >= 0)
                        bits[(row0 + (v - 1)) * (int64_t)W + (bi >> 6)] |= 1ULL << (bi & 63);
                }
                f = fe + 1;
            }
            if (ok && (col != n_samples || f <= le)) ok = false;   // fewer or more sample columns than the header
        }
        if (!ok) {
#pragma omp critical(sb_vcf_bad2)
            if (bad == 0 || i + 1 < bad) bad = i + 1;
        }
    }
    return bad ? -bad : 0;
}

}  // extern "C"
