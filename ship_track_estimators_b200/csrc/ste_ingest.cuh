// ste_ingest.cuh - CSV rows -> columns on the device.
//
// The reference reads one ship at a time with pandas (ship_track.py:107-195): parse the whole file,
// keep the rows of one id, build "yr-mo-dyThh:00:00" strings, parse them back into timestamps and
// difference them.  Here the raw bytes of the file are copied to the GPU once and ONE thread per
// row extracts what the filter needs - the time stamp as integer hours since 1970-01-01, latitude,
// longitude, the ship id (as a 64-bit key plus its byte span for the host to read the text of one
// row per ship) and the row label when the file carries an index column.  Grouping by ship and
// ordering inside a ship are sorts of these columns (ingest.py); nothing is parsed twice.
//
// Field syntax accepted (what pandas' C parser accepts in these files): comma separated, optional
// double quotes around a field (no embedded quotes, commas or newlines), optional trailing '\r'.
// Numbers: integers for yr/mo/dy/hr/labels; decimals with optional exponent for lat/lon, converted
// exactly (digits accumulated in a 64-bit integer, one correctly rounded division or product by a
// power of ten that is exact in fp64: the classic fast path, valid for <= 15 significant digits and
// |exponent| <= 22; anything longer is flagged for the host to re-read).  "NA", "NaN", "nan", "NULL"
// and the empty field are missing values (NaN), as in pandas.
#pragma once
#include <stdint.h>

namespace ste {

constexpr int kCsvColYr = 0, kCsvColMo = 1, kCsvColDy = 2, kCsvColHr = 3, kCsvColLat = 4, kCsvColLon = 5, kCsvColId = 6, kCsvColLabel = 7;
constexpr int kCsvTargets = 8;

// flags per row
constexpr int kCsvBadDate = 0x1;        // yr / mo / dy / hr not an integer (stray header rows, text)
constexpr int kCsvBadPosition = 0x2;    // lat / lon neither a number nor a missing value
constexpr int kCsvIdNotInteger = 0x4;   // the id is not an integer literal (ids are then compared as text)
constexpr int kCsvLabelNotInteger = 0x8;
constexpr int kCsvSlowNumber = 0x10;    // lat / lon with > 15 digits or a large exponent: host re-reads the field
constexpr int kCsvShortRow = 0x20;      // fewer fields than the columns asked for

struct CsvSpan {
    int begin, end;   // byte offsets inside the row, quotes stripped; begin > end: field absent
};

__device__ __forceinline__ bool csv_is_space(unsigned char c) { return c == ' ' || c == '\t'; }

// integer literal: [+-]digits, surrounded by optional blanks.  Returns false if anything else.
__device__ __forceinline__ bool csv_parse_int(const unsigned char *p, int n, long long &out) {
    int i = 0;
    while (i < n && csv_is_space(p[i])) ++i;
    while (n > i && csv_is_space(p[n - 1])) --n;
    if (i >= n) return false;
    bool neg = false;
    if (p[i] == '-' || p[i] == '+') neg = p[i++] == '-';
    if (i >= n || n - i > 18) return false;
    long long v = 0;
    for (; i < n; ++i) {
        const unsigned d = (unsigned)p[i] - '0';
        if (d > 9u) return false;
        v = v * 10 + d;
    }
    out = neg ? -v : v;
    return true;
}

__device__ __forceinline__ bool csv_is_missing(const unsigned char *p, int n) {
    if (n == 0) return true;
    if (n == 2) return p[0] == 'N' && p[1] == 'A';
    if (n == 3) return (p[0] == 'N' || p[0] == 'n') && (p[1] == 'a' || p[1] == 'A') && (p[2] == 'N' || p[2] == 'n');
    if (n == 4) return p[0] == 'N' && p[1] == 'U' && p[2] == 'L' && p[3] == 'L';
    return false;
}

// decimal literal -> double; 0 ok, 1 missing, 2 not a number, 3 needs the slow (host) path
__device__ __forceinline__ int csv_parse_double(const unsigned char *p, int n, double &out) {
    int i = 0;
    while (i < n && csv_is_space(p[i])) ++i;
    while (n > i && csv_is_space(p[n - 1])) --n;
    if (csv_is_missing(p + i, n - i)) {
        out = __longlong_as_double(0x7ff8000000000000ll);
        return 1;
    }
    bool neg = false;
    if (p[i] == '-' || p[i] == '+') neg = p[i++] == '-';
    unsigned long long mant = 0;
    int digits = 0, frac = 0, seen = 0;
    bool dot = false;
    for (; i < n; ++i) {
        const unsigned char c = p[i];
        if (c == '.') {
            if (dot) return 2;
            dot = true;
            continue;
        }
        const unsigned d = (unsigned)c - '0';
        if (d > 9u) break;
        ++seen;
        if (mant != 0 || d != 0) {
            if (digits >= 19) return 3;
            mant = mant * 10 + d;
            ++digits;
        }
        if (dot) ++frac;
    }
    if (seen == 0) return 2;
    int e10 = 0;
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
        ++i;
        bool eneg = false;
        if (i < n && (p[i] == '-' || p[i] == '+')) eneg = p[i++] == '-';
        if (i >= n) return 2;
        for (; i < n; ++i) {
            const unsigned d = (unsigned)p[i] - '0';
            if (d > 9u) return 2;
            if (e10 < 10000) e10 = e10 * 10 + (int)d;
        }
        if (eneg) e10 = -e10;
    }
    if (i != n) return 2;
    e10 -= frac;
    if (mant == 0) {
        out = neg ? -0.0 : 0.0;
        return 0;
    }
    if (digits > 15 || e10 > 22 || e10 < -22) return 3;
    const double pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                              1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const double m = (double)mant;   // exact: < 10^15 < 2^53
    const double v = e10 >= 0 ? m * pow10[e10] : m / pow10[-e10];   // one rounding: correctly rounded
    out = neg ? -v : v;
    return 0;
}

// days since 1970-01-01 of a proleptic Gregorian date (Hinnant's days_from_civil)
__device__ __forceinline__ long long csv_days_from_civil(long long y, long long m, long long d) {
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const long long yoe = y - era * 400;
    const long long doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const long long doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}

struct CsvArgs {
    const unsigned char *bytes;
    const int64_t *row_start;   // [n_rows + 1]: row r occupies [row_start[r], row_start[r+1] - 1) (the newline excluded)
    int64_t n_rows;
    int32_t cols[kCsvTargets];  // field index of each target, -1 = not wanted
    int64_t *hours;
    double *lat, *lon;
    uint64_t *id_key;
    int64_t *id_int;
    int32_t *id_off, *id_len;
    int64_t *label;
    int32_t *flags;
};

__global__ void __launch_bounds__(256) csv_parse_rows_kernel(const CsvArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_rows) return;
    const unsigned char *p = a.bytes + a.row_start[r];
    int n = (int)(a.row_start[r + 1] - 1 - a.row_start[r]);
    if (n > 0 && p[n - 1] == '\r') --n;
    CsvSpan span[kCsvTargets];
#pragma unroll
    for (int k = 0; k < kCsvTargets; ++k) span[k] = CsvSpan{1, 0};
    // one pass over the row: field boundaries (commas outside quotes), spans of the wanted fields
    int field = 0, begin = 0;
    bool quoted = false;
    for (int i = 0; i <= n; ++i) {
        const unsigned char c = i < n ? p[i] : ',';
        if (c == '"' && i < n) quoted = !quoted;
        if (c == ',' && !quoted) {
            int b = begin, e = i;
            if (e - b >= 2 && p[b] == '"' && p[e - 1] == '"') { ++b; --e; }
#pragma unroll
            for (int k = 0; k < kCsvTargets; ++k)
                if (a.cols[k] == field) span[k] = CsvSpan{b, e};
            ++field;
            begin = i + 1;
        }
    }
    int flags = 0;
    long long v[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (span[k].begin > span[k].end) flags |= kCsvShortRow | kCsvBadDate;
        else if (!csv_parse_int(p + span[k].begin, span[k].end - span[k].begin, v[k])) flags |= kCsvBadDate;
    }
    if (!(flags & kCsvBadDate) && (v[1] < 1 || v[1] > 12 || v[2] < 1 || v[2] > 31 || v[3] < 0 || v[3] > 99)) flags |= kCsvBadDate;
    a.hours[r] = (flags & kCsvBadDate) ? 0 : csv_days_from_civil(v[0], v[1], v[2]) * 24 + v[3];
    double pos[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const CsvSpan s = span[kCsvColLat + k];
        if (s.begin > s.end) {
            flags |= kCsvShortRow | kCsvBadPosition;
            continue;
        }
        const int rc = csv_parse_double(p + s.begin, s.end - s.begin, pos[k]);
        if (rc == 2) flags |= kCsvBadPosition;
        if (rc == 3) flags |= kCsvSlowNumber;
    }
    a.lat[r] = pos[0];
    a.lon[r] = pos[1];
    {   // id: FNV-1a of the bytes (quotes stripped), its span for the host, and its value if it is an integer literal
        const CsvSpan s = span[kCsvColId];
        uint64_t h = 1469598103934665603ull;
        long long iv = 0;
        if (s.begin > s.end) {
            flags |= kCsvShortRow | kCsvIdNotInteger;
        } else {
            for (int i = s.begin; i < s.end; ++i) h = (h ^ p[i]) * 1099511628211ull;
            if (!csv_parse_int(p + s.begin, s.end - s.begin, iv)) flags |= kCsvIdNotInteger;
        }
        a.id_key[r] = h;
        a.id_int[r] = iv;
        a.id_off[r] = s.begin;
        a.id_len[r] = s.begin > s.end ? 0 : s.end - s.begin;
    }
    long long lab = r;
    if (a.cols[kCsvColLabel] >= 0) {
        const CsvSpan s = span[kCsvColLabel];
        if (s.begin > s.end || !csv_parse_int(p + s.begin, s.end - s.begin, lab)) flags |= kCsvLabelNotInteger;
    }
    a.label[r] = lab;
    a.flags[r] = flags;
}

}  // namespace ste
