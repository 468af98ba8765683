// Batch builder: libsvm / libffm text -> COO arrays (host code, multi-threaded).
//
// Reference: rec/data/SampleParser.scala:23-51 (parseLIBSVM: "<label> <key>:<value> ...") and :53-85
// (parseLIBFFM: "<label> <field>:<key>:<value> ...").  One sample per line; non-zero i of sample r is
// emitted as (index[i] = r, feats[i] = key - 1) -- keys are 1-based in files (:37, :69) -- in file
// order, i.e. sample-major, which is the order every kernel of the path assumes.  The reference
// splits on single spaces and lets Java's parseLong / parseFloat throw on anything else; here runs of
// blanks / tabs and a trailing '\r' are tolerated and every other malformation is an error that
// names the line (B200REC_ERR_ARG, the stand-in for the NumberFormatException / AngelException).
//
// The GPU step consumes ~20 M samples/s, so the text side is parallel: the blob is cut at line ends
// into one range per host thread; pass 1 validates and counts (samples, non-zeros, lines) per range,
// a prefix sum gives every range its output offsets, pass 2 fills.  Numbers take a fast path
// (plain decimals: exact in double, then one rounding to float, with the rare double-rounding tie
// sent to strtof) so that results are bit-identical to strtof / Java parseFloat.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200rec {
namespace {

struct Cursor {
  const char* p;
  const char* end;   // end of the current line (exclusive)
};

inline bool is_blank(char ch) { return ch == ' ' || ch == '\t' || ch == '\r'; }
inline void skip_blanks(Cursor& c) {
  while (c.p < c.end && is_blank(*c.p)) ++c.p;
}

// decimal integer (optional sign); false when no digit was read or it does not fit 63 bits
inline bool read_int(Cursor& c, long long* out) {
  const char* p = c.p;
  bool neg = false;
  if (p < c.end && (*p == '-' || *p == '+')) neg = *p++ == '-';
  if (p >= c.end || *p < '0' || *p > '9') return false;
  unsigned long long v = 0;
  while (p < c.end && *p >= '0' && *p <= '9') {
    if (v > 922337203685477579ull) return false;
    v = v * 10 + (unsigned)(*p++ - '0');
  }
  *out = neg ? -(long long)v : (long long)v;
  c.p = p;
  return true;
}

const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                         1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// any form strtof takes, confined to the token (copied to a local buffer: strtof must not run past
// the line / the caller's buffer, which need not be NUL-terminated)
inline bool read_float_slow(Cursor& c, float* out) {
  char buf[64];
  int n = 0;
  while (c.p + n < c.end && n < 63) {
    const char ch = c.p[n];
    if (is_blank(ch) || ch == ':') break;
    buf[n++] = ch;
  }
  if (n == 0 || n == 63) return false;
  buf[n] = 0;
  char* stop = nullptr;
  const float v = strtof(buf, &stop);
  if (stop != buf + n) return false;
  *out = v;
  c.p += n;
  return true;
}

// [sign] digits [. digits] with at most 15 significant digits: mantissa and 10^frac are exact doubles,
// their quotient is correctly rounded, and rounding that to float equals strtof unless the double
// sits exactly half way between two floats (then strtof decides).  Everything else -> strtof.
inline bool read_float(Cursor& c, float* out) {
  const char* p = c.p;
  bool neg = false;
  if (p < c.end && (*p == '-' || *p == '+')) neg = *p++ == '-';
  unsigned long long mant = 0;
  int digits = 0, frac = 0;
  const char* d0 = p;
  while (p < c.end && *p >= '0' && *p <= '9') { mant = mant * 10 + (unsigned)(*p++ - '0'); ++digits; if (digits > 15) return read_float_slow(c, out); }
  const bool int_part = p > d0;
  if (p < c.end && *p == '.') {
    ++p;
    const char* f0 = p;
    while (p < c.end && *p >= '0' && *p <= '9') { mant = mant * 10 + (unsigned)(*p++ - '0'); ++digits; ++frac; if (digits > 15) return read_float_slow(c, out); }
    if (!int_part && p == f0) return false;
  } else if (!int_part) {
    return read_float_slow(c, out);   // inf, nan, or garbage: let strtof decide
  }
  if (p < c.end && !is_blank(*p) && *p != ':') return read_float_slow(c, out);   // exponent, hex, suffix ...
  const double d = (double)mant / kPow10[frac];
  unsigned long long bits;
  memcpy(&bits, &d, 8);
  if ((bits & 0x1fffffffull) == 0x10000000ull) return read_float_slow(c, out);   // float rounding tie
  const float v = (float)d;
  *out = neg ? -v : v;
  c.p = p;
  return true;
}

inline bool expect(Cursor& c, char ch) {
  if (c.p < c.end && *c.p == ch) { ++c.p; return true; }
  return false;
}

struct Out {
  float* targets; int* index; int* feats; int* fields; float* values;
};
struct RangeResult {
  long long ns = 0, nz = 0, lines = 0;
  int status = B200REC_OK;
  long long err_line = 0;   // 1-based, within the range
  char msg[96] = {0};
};

#define PARSE_FAIL(code, ...)                         \
  do {                                                \
    r.status = (code);                                \
    r.err_line = r.lines;                             \
    snprintf(r.msg, sizeof r.msg, __VA_ARGS__);       \
    return;                                           \
  } while (0)

// one range of whole lines; fill = false: validate and count, true: write at (ns0, nz0)
void parse_range(int format, const char* p, const char* end, bool fill, long long ns0, long long nz0,
                 const Out& o, RangeResult& r) {
  long long ns = 0, nz = 0;
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    Cursor c{p, eol ? eol : end};
    p = eol ? eol + 1 : end;
    ++r.lines;
    skip_blanks(c);
    if (c.p == c.end) continue;   // blank line
    float label;
    if (!read_float(c, &label)) PARSE_FAIL(B200REC_ERR_ARG, "label is not a number");
    if (fill) o.targets[ns0 + ns] = label;
    for (;;) {
      skip_blanks(c);
      if (c.p == c.end) break;
      long long field = 0, key = 0;
      float value = 0.f;
      if (format == 1 && !(read_int(c, &field) && expect(c, ':')))
        PARSE_FAIL(B200REC_ERR_ARG, "expected <field>:<key>:<value>");
      if (!(read_int(c, &key) && expect(c, ':') && read_float(c, &value)))
        PARSE_FAIL(B200REC_ERR_ARG, "expected %s", format == 1 ? "<field>:<key>:<value>" : "<key>:<value>");
      if (!(c.p == c.end || is_blank(*c.p))) PARSE_FAIL(B200REC_ERR_ARG, "trailing characters after a value");
      if (!(key >= 1 && key - 1 <= 2147483647LL))
        PARSE_FAIL(B200REC_ERR_INDEX, "key %lld is not a 1-based feature id that fits 31 bits", key);
      if (format == 1 && !(field >= -2147483648LL && field <= 2147483647LL))
        PARSE_FAIL(B200REC_ERR_ARG, "field %lld does not fit 32 bits", field);
      if (fill) {
        const long long i = nz0 + nz;
        o.index[i] = (int)(ns0 + ns);
        o.feats[i] = (int)(key - 1);
        if (o.fields) o.fields[i] = (int)field;
        if (o.values) o.values[i] = value;
      }
      ++nz;
    }
    ++ns;
  }
  r.ns = ns;
  r.nz = nz;
}

}  // namespace
}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_parse_samples(int format, const char* text, int64_t n_bytes, int64_t cap_samples,
                                     int64_t cap_nnz, float* targets, int* index, int* feats, int* fields,
                                     float* values, int64_t* n_samples, int64_t* nnz) {
  B200_REQUIRE(format == 0 || format == 1, B200REC_ERR_ARG, "unknown data format %d (0 libsvm, 1 libffm)", format);
  B200_REQUIRE((text || n_bytes == 0) && n_bytes >= 0 && n_samples && nnz, B200REC_ERR_ARG, "bad argument");
  const bool counting = !targets && !index && !feats && !fields && !values;
  B200_REQUIRE(counting || (targets && index && feats), B200REC_ERR_ARG,
               "targets, index and feats must be given together (or no output at all, to count)");
  // ranges of whole lines, one per thread (>= 256 KB each)
  int T = (int)std::min<long long>(std::max(1u, std::thread::hardware_concurrency()), 32);
  T = (int)std::max<long long>(1, std::min<long long>(T, n_bytes / (256 << 10)));
  std::vector<const char*> cut(T + 1);
  cut[0] = text;
  cut[T] = text + n_bytes;
  for (int t = 1; t < T; ++t) {
    const char* p = std::max(cut[t - 1], text + n_bytes * t / T);
    const char* nl = p < cut[T] ? (const char*)memchr(p, '\n', (size_t)(cut[T] - p)) : nullptr;
    cut[t] = nl ? nl + 1 : cut[T];
  }
  const Out out{targets, index, feats, fields, values};
  std::vector<RangeResult> res(T);
  auto run = [&](bool fill, const std::vector<long long>& ns0, const std::vector<long long>& nz0) {
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t)
      th.emplace_back([&, t] { RangeResult r; parse_range(format, cut[t], cut[t + 1], fill, ns0[t], nz0[t], out, r); res[t] = r; });
    RangeResult r0;
    parse_range(format, cut[0], cut[1], fill, ns0[0], nz0[0], out, r0);
    res[0] = r0;
    for (auto& x : th) x.join();
  };
  std::vector<long long> ns0(T, 0), nz0(T, 0);
  run(false, ns0, nz0);
  long long lines = 0, ns = 0, nz = 0;
  for (int t = 0; t < T; ++t) {
    if (res[t].status != B200REC_OK) {
      set_error("line %lld: %s", lines + res[t].err_line, res[t].msg);
      return res[t].status;
    }
    ns0[t] = ns; nz0[t] = nz;
    lines += res[t].lines; ns += res[t].ns; nz += res[t].nz;
  }
  *n_samples = ns;
  *nnz = nz;
  if (counting) return B200REC_OK;
  B200_REQUIRE(ns <= cap_samples && nz <= cap_nnz, B200REC_ERR_ARG,
               "%lld samples / %lld non-zeros do not fit the given arrays (%lld / %lld)", ns, nz,
               (long long)cap_samples, (long long)cap_nnz);
  B200_REQUIRE(ns <= 2147483647LL, B200REC_ERR_ARG, "too many samples for a 32-bit row index");
  run(true, ns0, nz0);
  return B200REC_OK;
}
