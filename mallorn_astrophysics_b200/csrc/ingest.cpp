// ingest.cpp -- host ingest of the reference's light-curve files (libmallorn_ingest.so, C-ABI in
// include/mallorn_ingest.h).  Plain C++17 + threads, no CUDA: it produces the flat columns K1 takes.
//
// Replaces, for the featurisation path, `load_lightcurves` (src/utils/data_loader.py:36-62: pd.read_csv of
// split_XX/{train,test}_full_lightcurves.csv, pd.concat) followed by the per-extractor
// `groupby('object_id')` / `Filter == band` masks (statistical.py:155,170 ...): the files are memory-mapped,
// cut into newline-aligned chunks, parsed by a pool of threads straight into
//     obj int32 (dense id, numbered by first appearance over the concatenated files = unique() order),
//     band uint8 (0..5 = u g r i z y, 6 = anything else), t / flux / err float64
// in file order, plus the id string table.  No DataFrame, no per-row Python objects.
//
// Numbers are converted exactly as the reference's pandas does: pd.read_csv's default C-engine converter
// ("high" precision: up to 17 significant digits accumulated in a double, one multiplication or division by
// a power of ten) -- which is NOT a correctly rounded strtod ('0.30000000000000004' reads as 0.3), so a
// correctly rounded parser would differ from the reference in the last bit.  tests/test_ingest_csv.py holds
// this file to pd.read_csv bit for bit.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <string_view>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "../../include/mallorn_ingest.h"

namespace {

thread_local std::string g_err;

const double P10[] = {
    1e0,   1e1,   1e2,   1e3,   1e4,   1e5,   1e6,   1e7,   1e8,   1e9,   1e10,  1e11,  1e12,  1e13,  1e14,  1e15,  1e16,  1e17,  1e18,  1e19,
    1e20,  1e21,  1e22,  1e23,  1e24,  1e25,  1e26,  1e27,  1e28,  1e29,  1e30,  1e31,  1e32,  1e33,  1e34,  1e35,  1e36,  1e37,  1e38,  1e39,
    1e40,  1e41,  1e42,  1e43,  1e44,  1e45,  1e46,  1e47,  1e48,  1e49,  1e50,  1e51,  1e52,  1e53,  1e54,  1e55,  1e56,  1e57,  1e58,  1e59,
    1e60,  1e61,  1e62,  1e63,  1e64,  1e65,  1e66,  1e67,  1e68,  1e69,  1e70,  1e71,  1e72,  1e73,  1e74,  1e75,  1e76,  1e77,  1e78,  1e79,
    1e80,  1e81,  1e82,  1e83,  1e84,  1e85,  1e86,  1e87,  1e88,  1e89,  1e90,  1e91,  1e92,  1e93,  1e94,  1e95,  1e96,  1e97,  1e98,  1e99,
    1e100, 1e101, 1e102, 1e103, 1e104, 1e105, 1e106, 1e107, 1e108, 1e109, 1e110, 1e111, 1e112, 1e113, 1e114, 1e115, 1e116, 1e117, 1e118, 1e119,
    1e120, 1e121, 1e122, 1e123, 1e124, 1e125, 1e126, 1e127, 1e128, 1e129, 1e130, 1e131, 1e132, 1e133, 1e134, 1e135, 1e136, 1e137, 1e138, 1e139,
    1e140, 1e141, 1e142, 1e143, 1e144, 1e145, 1e146, 1e147, 1e148, 1e149, 1e150, 1e151, 1e152, 1e153, 1e154, 1e155, 1e156, 1e157, 1e158, 1e159,
    1e160, 1e161, 1e162, 1e163, 1e164, 1e165, 1e166, 1e167, 1e168, 1e169, 1e170, 1e171, 1e172, 1e173, 1e174, 1e175, 1e176, 1e177, 1e178, 1e179,
    1e180, 1e181, 1e182, 1e183, 1e184, 1e185, 1e186, 1e187, 1e188, 1e189, 1e190, 1e191, 1e192, 1e193, 1e194, 1e195, 1e196, 1e197, 1e198, 1e199,
    1e200, 1e201, 1e202, 1e203, 1e204, 1e205, 1e206, 1e207, 1e208, 1e209, 1e210, 1e211, 1e212, 1e213, 1e214, 1e215, 1e216, 1e217, 1e218, 1e219,
    1e220, 1e221, 1e222, 1e223, 1e224, 1e225, 1e226, 1e227, 1e228, 1e229, 1e230, 1e231, 1e232, 1e233, 1e234, 1e235, 1e236, 1e237, 1e238, 1e239,
    1e240, 1e241, 1e242, 1e243, 1e244, 1e245, 1e246, 1e247, 1e248, 1e249, 1e250, 1e251, 1e252, 1e253, 1e254, 1e255, 1e256, 1e257, 1e258, 1e259,
    1e260, 1e261, 1e262, 1e263, 1e264, 1e265, 1e266, 1e267, 1e268, 1e269, 1e270, 1e271, 1e272, 1e273, 1e274, 1e275, 1e276, 1e277, 1e278, 1e279,
    1e280, 1e281, 1e282, 1e283, 1e284, 1e285, 1e286, 1e287, 1e288, 1e289, 1e290, 1e291, 1e292, 1e293, 1e294, 1e295, 1e296, 1e297, 1e298, 1e299,
    1e300, 1e301, 1e302, 1e303, 1e304, 1e305, 1e306, 1e307, 1e308};

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// eight ASCII digits at p -> their value (little-endian SWAR); false if any of the eight bytes is not a digit
inline bool eight_digits(const char* p, uint64_t& val) {
  uint64_t w;
  memcpy(&w, p, 8);
  const uint64_t hi = w + 0x4646464646464646ull, lo = w - 0x3030303030303030ull;
  if (((hi | lo) & 0x8080808080808080ull) != 0) return false;
  w = lo;
  w = (w * 10) + (w >> 8);                                                      // pairs in bytes 0, 2, 4, 6
  w = (((w & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((w >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
  val = w;
  return true;
}
inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// pandas' default converter (17 significant digits, one scaling by a power of ten), from the first character of the
// number.  Consumes the longest prefix that is a number in that grammar; returns false if there is none.
inline bool pandas_number_prefix(const char* p, const char* end, double& out, const char*& stop) {
  bool neg = false;
  if (p < end && *p == '-') { neg = true; ++p; } else if (p < end && *p == '+') ++p;
  // The reference accumulates number = number * 10 + digit in a double.  Up to 15 digits every partial value is an
  // exactly representable integer, so those are gathered in a uint64 (no floating-point dependency chain) and only
  // digits 16 and 17 -- where the double accumulation rounds -- are added the reference's way.
  double number = 0.0;
  uint64_t acc = 0;
  int exponent = 0, num_digits = 0, num_decimals = 0;
  const int max_digits = 17, exact_digits = 15;
  auto push = [&](int d) {
    if (num_digits < exact_digits) { acc = acc * 10 + (uint64_t)d; if (num_digits + 1 == exact_digits) number = (double)acc; }
    else number = number * 10.0 + d;
    ++num_digits;
  };
  while (p < end && is_digit(*p)) {
    if (num_digits < max_digits) push(*p - '0');
    else ++exponent;
    ++p;
  }
  if (p < end && *p == '.') {
    ++p;
    uint64_t eight;
    if (num_digits + 8 < exact_digits && end - p >= 8 && eight_digits(p, eight)) {   // the usual case: >= 8 decimals
      acc = acc * 100000000ull + eight; num_digits += 8; num_decimals += 8; p += 8;
    }
    while (num_digits < max_digits && p < end && is_digit(*p)) { push(*p - '0'); ++p; ++num_decimals; }
    if (num_digits >= max_digits) while (p < end && is_digit(*p)) ++p;
    exponent -= num_decimals;
  }
  if (num_digits < exact_digits) number = (double)acc;
  if (num_digits == 0) return false;
  if (neg) number = -number;
  if (p < end && (*p == 'e' || *p == 'E')) {
    const char* save = p;
    ++p;
    bool eneg = false;
    if (p < end && *p == '-') { eneg = true; ++p; } else if (p < end && *p == '+') ++p;
    int nd = 0, n = 0;
    while (nd < max_digits && p < end && is_digit(*p)) { if (n < 100000) n = n * 10 + (*p - '0'); ++nd; ++p; }
    if (nd == 0) p = save;
    else exponent += eneg ? -n : n;
  }
  if (exponent > 308) number = neg ? -HUGE_VAL : HUGE_VAL;
  else if (exponent > 0) number *= P10[exponent];
  else if (exponent < -308) {
    if (exponent < -616) number = neg ? -0.0 : 0.0;
    else { number /= P10[-308 - exponent]; number /= P10[308]; }
  } else if (exponent < 0) number /= P10[-exponent];
  out = number;
  stop = p;
  return true;
}

// a whole field: optional surrounding white space, then nothing but the number
bool pandas_strtod(const char* p, const char* end, double& out) {
  while (p < end && is_space(*p)) ++p;
  while (end > p && is_space(end[-1])) --end;
  if (p == end) return false;
  const char* stop;
  return pandas_number_prefix(p, end, out, stop) && stop == end;
}

bool ieq(std::string_view a, const char* b) {
  size_t n = strlen(b);
  if (a.size() != n) return false;
  for (size_t i = 0; i < n; ++i) { char c = a[i]; if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a'); if (c != b[i]) return false; }
  return true;
}

// pandas' default na_values (io/parsers: STR_NA_VALUES)
bool is_na_token(std::string_view s) {
  if (s.empty()) return true;
  switch (s[0]) { case '#': case '-': case '1': case '<': case 'N': case 'n': break; default: return false; }
  static const char* const NA[] = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
                                   "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
  for (const char* t : NA) if (s == t) return true;
  return false;
}

double parse_number(const char* b, const char* e, bool& bad) {
  double v;
  if (pandas_strtod(b, e, v)) return v;                    // (no NA spelling is a number in this grammar)
  std::string_view s(b, (size_t)(e - b));
  if (is_na_token(s)) return std::numeric_limits<double>::quiet_NaN();
  std::string_view t = s;
  bool neg = false;
  if (!t.empty() && (t[0] == '-' || t[0] == '+')) { neg = t[0] == '-'; t.remove_prefix(1); }
  if (ieq(t, "inf") || ieq(t, "infinity")) return neg ? -HUGE_VAL : HUGE_VAL;
  bad = true;
  return std::numeric_limits<double>::quiet_NaN();
}

struct Mapped {
  const char* data = nullptr; size_t size = 0; int fd = -1;
  ~Mapped() { if (data && size) munmap((void*)data, size); if (fd >= 0) close(fd); }
};

struct Columns { int id = -1, t = -1, f = -1, e = -1, band = -1, count = 0; };

std::string_view strip(std::string_view s) {
  while (!s.empty() && (s.back() == '\r' || s.back() == ' ')) s.remove_suffix(1);
  while (!s.empty() && s.front() == ' ') s.remove_prefix(1);
  if (s.size() >= 2 && s.front() == '"' && s.back() == '"') { s.remove_prefix(1); s.remove_suffix(1); }
  return s;
}

// id string -> dense code: open addressing, power-of-two capacity, the 64-bit hash kept beside the key so that a probe
// compares one word before touching the string (std::unordered_map spent half of the parse time here)
inline uint64_t load8(const char* p) { uint64_t w; memcpy(&w, p, 8); return w; }
inline uint64_t hash_bytes(const char* p, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdull);
  if (n >= 8) {                                     // whole words, then the last eight bytes again (overlapping): no variable-length copy
    const char* last = p + n - 8;
    for (; p < last; p += 8) { h = (h ^ load8(p)) * 0x9FB21C651E98DF25ull; h ^= h >> 29; }
    h = (h ^ load8(last)) * 0x9FB21C651E98DF25ull;
  } else {
    uint64_t w = 0;
    for (size_t k = 0; k < n; ++k) w |= (uint64_t)(unsigned char)p[k] << (8 * k);
    h = (h ^ w) * 0x9FB21C651E98DF25ull;
  }
  // full avalanche (the varying digits of an id sit in the HIGH bytes of the last word; a table index is the LOW bits)
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h;
}
inline bool same_bytes(const char* a, const char* b, size_t n) {      // n is the same for both
  if (n >= 8) {
    const char* last = a + n - 8;
    for (; a < last; a += 8, b += 8) if (load8(a) != load8(b)) return false;
    return load8(last) == load8(b + (last - a));
  }
  for (size_t k = 0; k < n; ++k) if (a[k] != b[k]) return false;
  return true;
}
struct IdTable {
  struct Slot { uint64_t h; uint32_t off; uint32_t n; int32_t code; };      // code < 0: empty
  std::vector<Slot> slots; size_t used = 0;
  std::vector<char> arena;             // key bytes copied here: comparing against the first occurrence inside the mapped
                                       // file was a cache and TLB miss per row
  IdTable() : slots(2048, Slot{0, 0, 0, -1}) { arena.reserve(1 << 16); }
  void grow() {
    std::vector<Slot> old; old.swap(slots);
    slots.assign(old.size() * 2, Slot{0, 0, 0, -1});
    for (const Slot& s : old) if (s.code >= 0) { size_t i = s.h & (slots.size() - 1); while (slots[i].code >= 0) i = (i + 1) & (slots.size() - 1); slots[i] = s; }
  }
  // returns the code of `id`, inserting it with `next_code` if new (then inserted = true)
  int32_t find_or_insert(std::string_view id, int32_t next_code, bool& inserted) {
    return find_or_insert(id, hash_bytes(id.data(), id.size()), next_code, inserted);
  }
  void prefetch(uint64_t h) const { __builtin_prefetch(&slots[h & (slots.size() - 1)]); }
  int32_t find_or_insert(std::string_view id, uint64_t h, int32_t next_code, bool& inserted) {
    size_t i = h & (slots.size() - 1);
    for (;;) {
      Slot& s = slots[i];
      if (s.code < 0) {
        s = Slot{h, (uint32_t)arena.size(), (uint32_t)id.size(), next_code};
        arena.insert(arena.end(), id.begin(), id.end());
        inserted = true;
        if (++used * 2 > slots.size()) grow();
        return next_code;
      }
      if (s.h == h && s.n == id.size() && same_bytes(arena.data() + s.off, id.data(), id.size())) { inserted = false; return s.code; }
      i = (i + 1) & (slots.size() - 1);
    }
  }
};

struct Chunk {
  const char* begin; const char* end; const Columns* cols;
  int64_t rows = 0, row0 = 0;
  std::vector<std::string_view> local_ids;                 // first-appearance order inside the chunk
  std::vector<int32_t> local_to_global;
  bool bad_number = false, bad_line = false;
};

}  // namespace

namespace {
// column storage without the zero fill of std::vector: the pages are first touched by the thread that parses into them
template <class T>
struct Buf {
  T* p = nullptr;
  bool alloc(size_t n) { p = static_cast<T*>(malloc(n ? n * sizeof(T) : 1)); return p != nullptr; }
  ~Buf() { free(p); }
  T* data() const { return p; }
  T& operator[](size_t i) const { return p[i]; }
};
}  // namespace

struct mlrn_table {
  int64_t n_rows = 0;
  Buf<int32_t> obj; Buf<uint8_t> band; Buf<double> t, flux, err;
  std::vector<int64_t> id_off; std::string id_chars;
};

namespace {

inline uint8_t band_code(std::string_view s) {
  if (s.size() == 1) switch (s[0]) { case 'u': return 0; case 'g': return 1; case 'r': return 2; case 'i': return 3; case 'z': return 4; case 'y': return 5; }
  return 6;
}

int64_t count_rows(const char* b, const char* e) {
  int64_t n = 0;
  const char* p = b;
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
    const char* le = nl ? nl : e;
    const char* q = le; while (q > p && (q[-1] == '\r')) --q;
    if (q > p) ++n;                                          // blank lines are skipped (skip_blank_lines=True)
    p = nl ? nl + 1 : e;
  }
  return n;
}

void parse_chunk(Chunk& c, mlrn_table& tb) {          // tb.obj receives chunk-local codes, made global afterwards
  IdTable seen;
  const Columns& col = *c.cols;
  int64_t r = c.row0;
  const char* p = c.begin;
  std::string_view prev_id; int32_t prev_code = -1;
  while (p < c.end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(c.end - p));
    const char* le = nl ? nl : c.end;
    const char* next = nl ? nl + 1 : c.end;
    while (le > p && le[-1] == '\r') --le;
    if (le == p) { p = next; continue; }
    // one pass over the line (the light-curve files have no quoted commas): numeric fields are converted where they
    // stand -- the number parser stops at the comma -- and only the id / Filter fields are scanned for their end
    int k = 0;
    std::string_view idv, bandv;
    double num[3] = {std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::quiet_NaN()};
    const char* q = p;
    for (;; ++k) {
      const char* fb = q;
      const int which = (k == col.t) ? 0 : (k == col.f) ? 1 : (k == col.e) ? 2 : -1;
      const char* stop;
      if (which >= 0 && pandas_number_prefix(fb, le, num[which], stop) && (stop == le || *stop == ',')) q = stop;
      else {
        while (q < le && *q != ',') ++q;
        if (which >= 0) num[which] = parse_number(fb, q, c.bad_number);         // blanks, NaN spellings, infinities
        else if (k == col.id) idv = strip(std::string_view(fb, (size_t)(q - fb)));
        else if (k == col.band) bandv = strip(std::string_view(fb, (size_t)(q - fb)));
      }
      if (q == le) { ++k; break; }
      ++q;
    }
    if (k != col.count) c.bad_line = true;
    tb.t[r] = num[0]; tb.flux[r] = num[1]; tb.err[r] = num[2];
    tb.band[r] = band_code(bandv);
    int32_t code = -1;
    if (idv == prev_id && prev_code >= 0) code = prev_code;   // real files list an object's rows together
    else if (!is_na_token(idv)) {
      bool fresh = false;
      code = seen.find_or_insert(idv, (int32_t)c.local_ids.size(), fresh);
      if (fresh) c.local_ids.push_back(idv);
    }
    tb.obj[r] = code;
    prev_id = idv; prev_code = code;
    ++r;
    p = next;
  }
}

template <class F>
void run_parallel(int n_jobs, int n_threads, F fn) {
  std::atomic<int> next{0};
  auto worker = [&]() { for (;;) { int j = next.fetch_add(1); if (j >= n_jobs) break; fn(j); } };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

const char* mlrn_ingest_last_error(void) { return g_err.c_str(); }

int mlrn_ingest_csv(const char* const* paths, int n_paths, int n_threads, mlrn_table** out) {
  if (!out || n_paths < 0 || (n_paths > 0 && !paths)) { g_err = "bad arguments"; return MLRN_INGEST_EINVAL; }
  *out = nullptr;
  if (n_threads <= 0) { n_threads = (int)std::thread::hardware_concurrency(); if (n_threads <= 0) n_threads = 1; }
  std::vector<Mapped> files((size_t)n_paths);
  std::vector<Columns> cols((size_t)n_paths);
  std::vector<Chunk> chunks;
  for (int i = 0; i < n_paths; ++i) {
    Mapped& m = files[(size_t)i];
    m.fd = open(paths[i], O_RDONLY);
    struct stat st;
    if (m.fd < 0 || fstat(m.fd, &st) != 0) { g_err = std::string("cannot open ") + paths[i] + ": " + strerror(errno); return MLRN_INGEST_EIO; }
    m.size = (size_t)st.st_size;
    if (m.size == 0) { g_err = std::string("empty file ") + paths[i]; return MLRN_INGEST_EFORMAT; }
    void* a = mmap(nullptr, m.size, PROT_READ, MAP_PRIVATE, m.fd, 0);
    if (a == MAP_FAILED) { m.size = 0; g_err = std::string("mmap failed for ") + paths[i]; return MLRN_INGEST_EIO; }
    m.data = (const char*)a;
    madvise(a, m.size, MADV_SEQUENTIAL);
    // header: column positions by name (structural error if one is missing, as a KeyError would be upstream)
    const char* nl = (const char*)memchr(m.data, '\n', m.size);
    const char* he = nl ? nl : m.data + m.size;
    Columns& c = cols[(size_t)i];
    int k = 0; const char* fb = m.data;
    for (const char* q = m.data;; ++q) {
      if (q == he || *q == ',') {
        std::string_view name = strip(std::string_view(fb, (size_t)(q - fb)));
        if (name == "object_id") c.id = k; else if (name == "Time (MJD)") c.t = k; else if (name == "Flux") c.f = k;
        else if (name == "Flux_err") c.e = k; else if (name == "Filter") c.band = k;
        ++k; fb = q + 1;
        if (q == he) break;
      }
    }
    c.count = k;
    if (c.id < 0 || c.t < 0 || c.f < 0 || c.e < 0 || c.band < 0) {
      g_err = std::string(paths[i]) + ": header must name object_id, Time (MJD), Flux, Flux_err, Filter"; return MLRN_INGEST_EFORMAT;
    }
    const char* body = nl ? nl + 1 : m.data + m.size;
    const char* fend = m.data + m.size;
    // newline-aligned chunks of ~8 MiB
    const size_t target = 8u << 20;
    while (body < fend) {
      const char* ce = (size_t)(fend - body) > target ? body + target : fend;
      if (ce < fend) { const char* q = (const char*)memchr(ce, '\n', (size_t)(fend - ce)); ce = q ? q + 1 : fend; }
      Chunk ch; ch.begin = body; ch.end = ce; ch.cols = &c;
      chunks.push_back(std::move(ch));
      body = ce;
    }
  }
  const int nc = (int)chunks.size();
  run_parallel(nc, n_threads, [&](int j) { chunks[(size_t)j].rows = count_rows(chunks[(size_t)j].begin, chunks[(size_t)j].end); });
  int64_t total = 0;
  for (auto& c : chunks) { c.row0 = total; total += c.rows; }
  if (total > 2147483647LL) { g_err = "more than 2^31-1 rows: ingest in several batches"; return MLRN_INGEST_EINVAL; }
  mlrn_table* tb = new (std::nothrow) mlrn_table();
  if (!tb) { g_err = "out of memory"; return MLRN_INGEST_EIO; }
  tb->n_rows = total;
  if (!(tb->obj.alloc((size_t)total) && tb->band.alloc((size_t)total) && tb->t.alloc((size_t)total) &&
        tb->flux.alloc((size_t)total) && tb->err.alloc((size_t)total))) { delete tb; g_err = "out of memory"; return MLRN_INGEST_EIO; }
  run_parallel(nc, n_threads, [&](int j) { parse_chunk(chunks[(size_t)j], *tb); });
  for (auto& c : chunks) {
    if (c.bad_line) { delete tb; g_err = "a data line does not have the header's number of fields"; return MLRN_INGEST_EFORMAT; }
    if (c.bad_number) { delete tb; g_err = "a Time/Flux/Flux_err field is not a number"; return MLRN_INGEST_EFORMAT; }
  }
  // global ids in order of first appearance over the concatenated files (== pd.concat(...)['object_id'].unique())
  IdTable global;
  tb->id_off.push_back(0);
  for (auto& c : chunks) {
    c.local_to_global.resize(c.local_ids.size());
    for (size_t k = 0; k < c.local_ids.size(); ++k) {
      bool fresh = false;
      const int32_t gid = global.find_or_insert(c.local_ids[k], (int32_t)(tb->id_off.size() - 1), fresh);
      if (fresh) {
        tb->id_chars.append(c.local_ids[k].data(), c.local_ids[k].size());
        tb->id_off.push_back((int64_t)tb->id_chars.size());
      }
      c.local_to_global[k] = gid;
    }
  }
  run_parallel(nc, n_threads, [&](int j) {
    const Chunk& c = chunks[(size_t)j];
    for (int64_t r = c.row0; r < c.row0 + c.rows; ++r) {
      const int32_t lc = tb->obj[(size_t)r];
      tb->obj[(size_t)r] = lc < 0 ? -1 : c.local_to_global[(size_t)lc];
    }
  });
  *out = tb;
  return MLRN_INGEST_OK;
}

// ---- factorisation of an Arrow string column (include/mallorn_ingest.h) -------------------------------------------------
namespace {
struct Utf8Chunk {                   // one Arrow (large_)string array: rows [0, n) are logical rows offset .. offset+n of its buffers
  const void* offsets; const char* data; const uint8_t* validity; int64_t offset, n, row0; int width;
  inline bool valid(int64_t i) const { const int64_t k = offset + i; return !validity || ((validity[k >> 3] >> (k & 7)) & 1); }
  inline std::string_view at(int64_t i) const {
    const int64_t k = offset + i;
    if (width == 8) { const int64_t* o = (const int64_t*)offsets; return std::string_view(data + o[k], (size_t)(o[k + 1] - o[k])); }
    const int32_t* o = (const int32_t*)offsets; return std::string_view(data + o[k], (size_t)(o[k + 1] - o[k]));
  }
};
struct Utf8Job { int chunk; int64_t lo, hi; std::vector<std::string_view> local_ids; std::vector<int64_t> local_first; std::vector<int32_t> local_to_global; };

bool utf8_chunks(int n_chunks, const void* const* offsets, const uint8_t* const* data, const uint8_t* const* validity,
                 const int64_t* array_offset, const int64_t* length, int offset_bytes, std::vector<Utf8Chunk>& out, int64_t& rows) {
  if (n_chunks < 0 || (offset_bytes != 4 && offset_bytes != 8)) return false;
  rows = 0;
  for (int c = 0; c < n_chunks; ++c) {
    if (length[c] < 0 || array_offset[c] < 0 || (length[c] > 0 && (!offsets[c]))) return false;
    out.push_back(Utf8Chunk{offsets[c], (const char*)data[c], validity ? validity[c] : nullptr, array_offset[c], length[c], rows, offset_bytes});
    rows += length[c];
  }
  return rows <= 0x7fffffffLL;
}
}  // namespace

int mlrn_factorize_utf8(int n_chunks, const void* const* offsets, const uint8_t* const* data, const uint8_t* const* validity,
                        const int64_t* array_offset, const int64_t* length, int offset_bytes, int n_threads,
                        int32_t* codes, int64_t* first_row, int32_t* n_unique) {
  std::vector<Utf8Chunk> chunks; int64_t rows = 0;
  if (!codes || !first_row || !n_unique || !utf8_chunks(n_chunks, offsets, data, validity, array_offset, length, offset_bytes, chunks, rows)) {
    g_err = "mlrn_factorize_utf8: bad argument (or more than 2^31-1 rows)"; return MLRN_INGEST_EINVAL;
  }
  if (n_threads <= 0) { n_threads = (int)std::thread::hardware_concurrency(); if (n_threads <= 0) n_threads = 1; }
  // jobs: contiguous row ranges inside one chunk, about rows / n_threads each (few jobs: every job's distinct ids are merged serially)
  const int64_t per = std::max<int64_t>(16384, (rows + n_threads - 1) / n_threads);
  std::vector<Utf8Job> jobs;
  for (int c = 0; c < n_chunks; ++c)
    for (int64_t lo = 0; lo < chunks[(size_t)c].n; lo += per) jobs.push_back(Utf8Job{c, lo, std::min(lo + per, chunks[(size_t)c].n), {}, {}, {}});
  try {
    run_parallel((int)jobs.size(), n_threads, [&](int j) {
      Utf8Job& job = jobs[(size_t)j];
      const Utf8Chunk& ch = chunks[(size_t)job.chunk];
      IdTable local;
      std::string_view prev; int32_t prev_code = -1;
      constexpr int BLOCK = 16;                     // hashes of a block first, with a prefetch of each slot: the probes of a shuffled
      uint64_t hs[BLOCK];                           // file are cache misses, and a mispredicted probe branch serialises them otherwise
      for (int64_t b0 = job.lo; b0 < job.hi; b0 += BLOCK) {
        const int nb = (int)std::min<int64_t>(BLOCK, job.hi - b0);
        for (int k = 0; k < nb; ++k)
          if (ch.valid(b0 + k)) { const std::string_view id = ch.at(b0 + k); hs[k] = hash_bytes(id.data(), id.size()); local.prefetch(hs[k]); }
        for (int k = 0; k < nb; ++k) {
          const int64_t i = b0 + k;
          int32_t code = -1;
          if (ch.valid(i)) {
            const std::string_view id = ch.at(i);
            if (prev_code >= 0 && id.size() == prev.size() && same_bytes(id.data(), prev.data(), id.size())) code = prev_code;     // files grouped by object
            else {
              bool fresh = false;
              code = local.find_or_insert(id, hs[k], (int32_t)job.local_ids.size(), fresh);
              if (fresh) { job.local_ids.push_back(id); job.local_first.push_back(ch.row0 + i); }
              prev = id; prev_code = code;
            }
          }
          codes[(size_t)(ch.row0 + i)] = code;
        }
      }
    });
    // global codes in order of first appearance over all rows (== pandas.factorize(sort=False)): jobs are in row order
    IdTable global; int32_t n_glob = 0;
    for (auto& job : jobs) {
      job.local_to_global.resize(job.local_ids.size());
      for (size_t k = 0; k < job.local_ids.size(); ++k) {
        bool fresh = false;
        const int32_t gid = global.find_or_insert(job.local_ids[k], n_glob, fresh);
        if (fresh) { first_row[(size_t)n_glob] = job.local_first[k]; ++n_glob; }
        job.local_to_global[k] = gid;
      }
    }
    run_parallel((int)jobs.size(), n_threads, [&](int j) {
      const Utf8Job& job = jobs[(size_t)j];
      const int64_t r0 = chunks[(size_t)job.chunk].row0;
      for (int64_t r = r0 + job.lo; r < r0 + job.hi; ++r) { const int32_t lc = codes[(size_t)r]; if (lc >= 0) codes[(size_t)r] = job.local_to_global[(size_t)lc]; }
    });
    *n_unique = n_glob;
  } catch (const std::bad_alloc&) { g_err = "mlrn_factorize_utf8: out of memory"; return MLRN_INGEST_EIO; }
  return MLRN_INGEST_OK;
}

int mlrn_band_codes_utf8(int n_chunks, const void* const* offsets, const uint8_t* const* data, const uint8_t* const* validity,
                         const int64_t* array_offset, const int64_t* length, int offset_bytes, int n_threads, uint8_t* band) {
  std::vector<Utf8Chunk> chunks; int64_t rows = 0;
  if (!band || !utf8_chunks(n_chunks, offsets, data, validity, array_offset, length, offset_bytes, chunks, rows)) {
    g_err = "mlrn_band_codes_utf8: bad argument (or more than 2^31-1 rows)"; return MLRN_INGEST_EINVAL;
  }
  if (n_threads <= 0) { n_threads = (int)std::thread::hardware_concurrency(); if (n_threads <= 0) n_threads = 1; }
  struct Job { int chunk; int64_t lo, hi; };
  const int64_t per = std::max<int64_t>(65536, (rows + n_threads - 1) / n_threads);
  std::vector<Job> jobs;
  for (int c = 0; c < n_chunks; ++c)
    for (int64_t lo = 0; lo < chunks[(size_t)c].n; lo += per) jobs.push_back(Job{c, lo, std::min(lo + per, chunks[(size_t)c].n)});
  run_parallel((int)jobs.size(), n_threads, [&](int j) {
    const Job& job = jobs[(size_t)j];
    const Utf8Chunk& ch = chunks[(size_t)job.chunk];
    for (int64_t i = job.lo; i < job.hi; ++i) band[(size_t)(ch.row0 + i)] = ch.valid(i) ? band_code(ch.at(i)) : (uint8_t)6;
  });
  return MLRN_INGEST_OK;
}

int64_t mlrn_table_rows(const mlrn_table* t) { return t ? t->n_rows : 0; }
int32_t mlrn_table_objects(const mlrn_table* t) { return t ? (int32_t)(t->id_off.size() - 1) : 0; }
const int32_t* mlrn_table_obj(const mlrn_table* t) { return t ? t->obj.data() : nullptr; }
const uint8_t* mlrn_table_band(const mlrn_table* t) { return t ? t->band.data() : nullptr; }
const double* mlrn_table_time(const mlrn_table* t) { return t ? t->t.data() : nullptr; }
const double* mlrn_table_flux(const mlrn_table* t) { return t ? t->flux.data() : nullptr; }
const double* mlrn_table_err(const mlrn_table* t) { return t ? t->err.data() : nullptr; }
int64_t mlrn_table_id_bytes(const mlrn_table* t) { return t ? (int64_t)t->id_chars.size() : 0; }
const int64_t* mlrn_table_id_offsets(const mlrn_table* t) { return t ? t->id_off.data() : nullptr; }
const char* mlrn_table_id_chars(const mlrn_table* t) { return t ? t->id_chars.data() : nullptr; }
void mlrn_table_free(mlrn_table* t) { delete t; }

}  // extern "C"
