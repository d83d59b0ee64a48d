// csv_pack.cpp -- SURVEY.md 8(f) rank 2: gene presence/absence CSV -> packed bitset rows, natively.
//
// Replaces the per-cell Python loop of Csv_to_dic_Roary (scoary/methods.py:445-497): for every
// data row, the cells of the selected isolate columns are turned into presence bits ("present"
// unless the cell is "", "0" or "-", methods.py:476-487) and packed straight into the uint64
// rows sb_set_genes takes; the byte ranges of a few leading fields (gene name, annotation, ...)
// are returned so the host can slice them out without parsing the row again.
//
// CSV dialect = Python's csv.reader(skipinitialspace=True, delimiter=d) as the reference uses it
// (methods.py:350-351): excel dialect, '"' quoting recognised at the start of a field (after the
// skipped spaces), "" inside quotes is a literal quote, rows end at \n, \r\n or \r outside quotes.
// Host code only (no CUDA); rows are parsed in parallel with OpenMP after a sequential scan for
// the row starts.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Field {
    int64_t begin, end;   // byte range of the raw field text (quotes included when quoted)
    bool quoted;
};

// Parses one row starting at p (end of buffer e).  Calls f(col, begin, end, quoted, needs_unescape)
// for every field; returns the position after the row terminator.
template <class F>
inline const char *parse_row(const char *base, const char *p, const char *e, char delim, F &&f)
{
    int col = 0;
    for (;;) {
        // the two shapes nearly every presence cell has -- one plain character, or nothing, before the delimiter --
        // without the general scan (same result: no space to skip, no quote to open, the field ends at the delimiter)
        if (p + 1 < e && p[1] == delim && p[0] != ' ' && p[0] != '"' && p[0] != delim && p[0] != '\n' && p[0] != '\r') {
            f(col, p - base, p + 1 - base, false, false);
            ++col;
            p += 2;
            continue;
        }
        if (p < e && p[0] == '"') {                 // "text" + delimiter, no quote inside: what Roary writes
            const char *q = (const char *)memchr(p + 1, '"', (size_t)(e - (p + 1)));
            if (q && q + 1 < e && q[1] == delim) {
                f(col, p + 1 - base, q - base, true, false);
                ++col;
                p = q + 2;
                continue;
            }
        }
        while (p < e && *p == ' ') ++p;             // skipinitialspace
        const char *fb = p;
        bool quoted = false, esc = false;
        if (p < e && *p == '"') {
            quoted = true;
            ++p;
            fb = p;
            const char *fe;
            for (;;) {
                if (p >= e) { fe = e; break; }
                if (*p == '"') {
                    if (p + 1 < e && p[1] == '"') { esc = true; p += 2; continue; }
                    fe = p;
                    ++p;
                    break;
                }
                ++p;
            }
            // anything between the closing quote and the delimiter belongs to the field in Python's
            // reader; Roary files never have it: skip to the delimiter
            const char *tail = p;
            while (p < e && *p != delim && *p != '\n' && *p != '\r') ++p;
            if (p != tail) esc = true;              // flag unusual field: host re-parses it
            f(col, fb - base, (esc ? p : fe) - base, quoted, esc);
        } else {
            while (p < e && *p != delim && *p != '\n' && *p != '\r') ++p;
            f(col, fb - base, p - base, false, false);
        }
        ++col;
        if (p >= e) return e;
        if (*p == delim) { ++p; continue; }
        if (*p == '\r') { ++p; if (p < e && *p == '\n') ++p; return p; }
        if (*p == '\n') return p + 1;
    }
}

inline bool is_absent(const char *b, const char *e)
{
    const int64_t n = e - b;
    return n == 0 || (n == 1 && (b[0] == '0' || b[0] == '-'));
}

}  // namespace

extern "C" {

// Row starts of the data rows (everything after the first row).  row_starts may be NULL to count.
// Returns the number of data rows (empty trailing lines are ignored), or -1 on error.
// (The work is in scan_rows below; with grow != NULL the starts go to a malloc'ed array of the right size instead:
// sb_csv_scan_rows, one pass over the file where count-then-fill takes two.)
static int64_t scan_rows(const char *buf, int64_t len, char delimiter, int64_t *row_starts, int64_t max_rows,
                         int64_t *header_end, int64_t **grow);

int64_t sb_csv_row_starts(const char *buf, int64_t len, char delimiter, int64_t *row_starts, int64_t max_rows,
                          int64_t *header_end)
{
    return scan_rows(buf, len, delimiter, row_starts, max_rows, header_end, nullptr);
}

// As sb_csv_row_starts in ONE pass: *row_starts_out receives a malloc'ed array of the row starts (release it with
// sb_csv_free; NULL when there are no rows).  Returns the number of data rows or -1.
int64_t sb_csv_scan_rows(const char *buf, int64_t len, char delimiter, int64_t **row_starts_out, int64_t *header_end)
{
    if (!row_starts_out) return -1;
    *row_starts_out = nullptr;
    return scan_rows(buf, len, delimiter, nullptr, 0, header_end, row_starts_out);
}

void sb_csv_free(void *p) { free(p); }

static int64_t scan_rows(const char *buf, int64_t len, char delimiter, int64_t *row_starts, int64_t max_rows,
                         int64_t *header_end, int64_t **grow)
{
    if (!buf || len < 0) return -1;
    const bool collect = row_starts != nullptr || grow != nullptr;
    auto deliver = [&](int64_t n) -> bool {          // where n starts go: the caller's array, or a new one
        if (grow) {
            if (n == 0) return true;
            *grow = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
            if (!*grow) return false;
            row_starts = *grow;
            return true;
        }
        return !row_starts || n <= max_rows;
    };
    const char *p = buf, *e = buf + len;
    // The reader's state machine (Python's _csv.c, excel dialect, skipinitialspace, non-strict): a '"' opens a quoted
    // field only as the first character of a field (after the skipped spaces); anywhere else it is a literal, so an
    // unquoted cell such as `5" nuclease` does not swallow the row terminator.  Inside a quoted field "" is a literal
    // quote and newlines belong to the field; after the closing quote the rest up to the delimiter is plain text.
    enum { START_FIELD, IN_FIELD, IN_QUOTED, QUOTE_IN_QUOTED };
    auto next_row = [&](const char *q) {
        int st = START_FIELD;
        for (; q < e; ++q) {
            const char c = *q;
            if (st == IN_QUOTED) {
                if (c == '"') st = QUOTE_IN_QUOTED;
                continue;
            }
            if (c == '\n') return q + 1;
            if (c == '\r') return (q + 1 < e && q[1] == '\n') ? q + 2 : q + 1;
            if (c == delimiter) { st = START_FIELD; continue; }
            if (st == START_FIELD) {
                if (c == ' ') continue;
                st = (c == '"') ? IN_QUOTED : IN_FIELD;
            } else if (st == QUOTE_IN_QUOTED) {
                st = (c == '"') ? IN_QUOTED : IN_FIELD;
            }
        }
        return e;
    };
    p = next_row(p);   // skip the header row
    if (header_end) *header_end = p - buf;
    // rows of [from, to): appends their starts to `out` (if given), returns where the last row ended
    auto scan = [&](const char *from, const char *to, std::vector<int64_t> *out, int64_t *count) {
        const char *r = from;
        while (r < to) {
            const char *q = next_row(r);
            const char *t = r;
            while (t < q && (*t == '\n' || *t == '\r')) ++t;      // rows that are empty are ignored
            if (t < q) {
                if (out) out->push_back(r - buf);
                ++*count;
            }
            r = q;
        }
        return r;
    };
    // Large files: the scan is split at line feeds and the pieces are scanned in parallel, each ASSUMING that its first
    // byte starts a row.  Piece 0 starts at a true row start; if every piece ends exactly where the next one begins, all
    // the assumptions were true (induction).  A line feed inside a quoted field breaks the chain -- then, or for small
    // files, the plain sequential scan below decides.
    int n_pieces = 1;
#ifdef _OPENMP
    if (e - p > (4 << 20)) n_pieces = (int)std::min<int64_t>(4 * (int64_t)omp_get_max_threads(), (e - p) >> 20);
#endif
    if (n_pieces > 1) {
        std::vector<const char *> cut(n_pieces + 1);
        cut[0] = p;
        cut[n_pieces] = e;
        for (int k = 1; k < n_pieces; ++k) {
            const char *nominal = p + (e - p) / n_pieces * k;
            const char *nl = (const char *)memchr(nominal, '\n', (size_t)(e - nominal));
            cut[k] = nl ? nl + 1 : e;
        }
        std::vector<std::vector<int64_t>> starts(n_pieces);
        std::vector<int64_t> counts(n_pieces, 0);
        std::vector<const char *> ended(n_pieces);
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < n_pieces; ++k)
            ended[k] = cut[k] < cut[k + 1] ? scan(cut[k], cut[k + 1], collect ? &starts[k] : nullptr, &counts[k]) : cut[k];
        bool chain = true;
        for (int k = 0; k < n_pieces; ++k) chain = chain && (ended[k] == cut[k + 1] || cut[k] >= cut[k + 1]);
        if (chain) {
            int64_t total = 0;
            for (int k = 0; k < n_pieces; ++k) total += counts[k];
            if (!deliver(total)) return -1;
            int64_t n = 0;
            for (int k = 0; k < n_pieces; ++k) {
                if (row_starts && counts[k]) memcpy(row_starts + n, starts[k].data(), sizeof(int64_t) * (size_t)counts[k]);
                n += counts[k];
            }
            return n;
        }
    }
    std::vector<int64_t> all;
    int64_t n = 0;
    scan(p, e, collect ? &all : nullptr, &n);
    if (!deliver(n)) return -1;
    if (row_starts && n) memcpy(row_starts, all.data(), sizeof(int64_t) * (size_t)n);
    return n;
}

// Pack the presence bits of n_rows data rows.
//   keep_cols  [n_keep] column indices (0-based, ascending) whose cells become bits 0..n_keep-1
//   bits       uint64 [n_rows][W] (W >= ceil(n_keep / 64)), zero-filled by this call
//   lead_cols  [n_lead] column indices whose byte ranges are returned in lead_ranges [n_rows][n_lead][2];
//              a negative begin marks a field the host must re-parse (escaped quotes etc.)
//   row_fields [n_rows] number of fields found in each row (the host checks it against the header)
// Returns 0, or -(row + 1) for the first row that has fewer fields than the largest needed column.
int64_t sb_csv_pack_rows(const char *buf, int64_t len, char delimiter, const int64_t *row_starts, int64_t n_rows,
                         const int32_t *keep_cols, int32_t n_keep, uint64_t *bits, int32_t W, const int32_t *lead_cols,
                         int32_t n_lead, int64_t *lead_ranges, int32_t *row_fields)
{
    if (!buf || !row_starts || !bits || (n_keep > 0 && !keep_cols)) return -1;
    const char *e = buf + len;
    int32_t max_col = -1;
    for (int32_t i = 0; i < n_keep; ++i) max_col = keep_cols[i] > max_col ? keep_cols[i] : max_col;
    for (int32_t i = 0; i < n_lead; ++i) max_col = lead_cols[i] > max_col ? lead_cols[i] : max_col;
    // column -> bit index (or -1) and column -> lead slot (or -1)
    std::vector<int32_t> bit_of(max_col + 1, -1), lead_of(max_col + 1, -1);
    for (int32_t i = 0; i < n_keep; ++i) bit_of[keep_cols[i]] = i;
    for (int32_t i = 0; i < n_lead; ++i) lead_of[lead_cols[i]] = i;
    int64_t first_bad = 0;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; ++r) {
        uint64_t *row = bits + r * (int64_t)W;
        memset(row, 0, sizeof(uint64_t) * (size_t)W);
        int32_t nf = 0;
        parse_row(buf, buf + row_starts[r], e, delimiter,
                  [&](int col, int64_t b, int64_t en, bool quoted, bool esc) {
                      nf = col + 1;
                      if (col > max_col) return;
                      const int32_t bi = bit_of[col];
                      if (bi >= 0) {
                          // escaped / unusual fields are never "", "0" or "-" unless they unescape to that;
                          // a quoted field with an escaped quote has length >= 1 and contains '"': present
                          const bool absent = !esc && is_absent(buf + b, buf + en);
                          if (!absent) row[bi >> 6] |= 1ULL << (bi & 63);
                      }
                      const int32_t li = lead_of[col];
                      if (li >= 0 && lead_ranges) {
                          int64_t *o = lead_ranges + (r * (int64_t)n_lead + li) * 2;
                          o[0] = esc ? -(b + 1) : b;
                          o[1] = en;
                          (void)quoted;
                      }
                  });
        if (row_fields) row_fields[r] = nf;
        if (nf <= max_col) {
#pragma omp critical(sb_csv_bad)
            {
                if (first_bad == 0 || r + 1 < first_bad) first_bad = r + 1;
            }
        }
    }
    return first_bad ? -first_bad : 0;
}

// The text of lead field k of every row, back to back: offsets[r] .. offsets[r + 1] is row r's slice of `out`
// (empty for fields the host must re-parse, begin < 0).  With out == NULL only the offsets are filled.
// Returns the total number of bytes, or -1 if out_cap is too small.  One decode + n slices on the host instead
// of n (slice, decode) pairs: the identifier columns of a million rows in 0.2 s instead of 2.
int64_t sb_csv_gather_fields(const char *buf, const int64_t *lead_ranges, int64_t n_rows, int32_t n_lead, int32_t k,
                             char *out, int64_t out_cap, int64_t *offsets)
{
    if (!buf || !lead_ranges || !offsets || k < 0 || k >= n_lead) return -1;
    int64_t total = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t *o = lead_ranges + (r * (int64_t)n_lead + k) * 2;
        offsets[r] = total;
        if (o[0] >= 0 && o[1] > o[0]) total += o[1] - o[0];
    }
    offsets[n_rows] = total;
    if (!out) return total;
    if (total > out_cap) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t *o = lead_ranges + (r * (int64_t)n_lead + k) * 2;
        if (o[0] >= 0 && o[1] > o[0]) memcpy(out + offsets[r], buf + o[0], (size_t)(o[1] - o[0]));
    }
    return total;
}

}  // extern "C"
