// Host-side audio ingest: FLAC and RIFF/WAVE files -> float32 samples in caller-owned
// (pinned) memory, many files at a time on host threads.
//
// Replaces the per-item Python decode of the reference's loaders (raw_dataset.py:20-28 `librosa.load(path, sr=16000)`
// with a `soundfile.read` fallback, :61-66; preprocess.py reads the same files before LFCC extraction): integer PCM is
// scaled by 2^-(bits-1) as libsndfile does, and multi-channel files are averaged to mono as librosa.load(mono=True)
// does.  There is no resampler: the ASVspoof corpora are 16 kHz and a file at another rate is an error the caller
// sees (sample_rate is returned), not something silently converted.
//
// The FLAC decoder is written from the format specification (RFC 9639): STREAMINFO, frame headers (fixed / variable
// blocking, every block-size and sample-rate code, CRC-8), CONSTANT / VERBATIM / FIXED / LPC subframes with wasted
// bits, partitioned Rice residuals (4- and 5-bit parameters, escape partitions), left-side / right-side / mid-side
// decorrelation, CRC-16 per frame and, on request, the MD5 signature of the decoded audio against STREAMINFO.
// No third-party code; plain C++17 + pthreads, no CUDA.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "air_b200.h"              // status codes + the declarations this file implements

namespace air_audio {

// ------------------------------------------------------------------------------------------------ file -> bytes
static int read_file(const char* path, std::vector<uint8_t>& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return AIR_ERR_IO;
  if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return AIR_ERR_IO; }
  const long n = ftell(f);
  if (n < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); return AIR_ERR_IO; }
  buf.resize((size_t)n);
  const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n ? AIR_OK : AIR_ERR_IO;
}

struct Pcm {                       // decoded audio: interleaved int32 samples
  int sample_rate = 0, channels = 0, bits = 0;
  long long frames = 0;
  std::vector<int32_t> data;
};

// ------------------------------------------------------------------------------------------------ checksums
static uint8_t crc8_table[256];
static uint16_t crc16_table[256];
static std::atomic<int> tables_ready{0};
static void init_tables() {
  if (tables_ready.load(std::memory_order_acquire)) return;
  for (int i = 0; i < 256; ++i) {
    uint8_t c = (uint8_t)i;
    for (int k = 0; k < 8; ++k) c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
    crc8_table[i] = c;
    uint16_t d = (uint16_t)(i << 8);
    for (int k = 0; k < 8; ++k) d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1));
    crc16_table[i] = d;
  }
  tables_ready.store(1, std::memory_order_release);
}
static uint8_t crc8(const uint8_t* p, size_t n) {
  uint8_t c = 0;
  for (size_t i = 0; i < n; ++i) c = crc8_table[c ^ p[i]];
  return c;
}
static uint16_t crc16(const uint8_t* p, size_t n) {
  uint16_t c = 0;
  for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ crc16_table[(c >> 8) ^ p[i]]);
  return c;
}

struct Md5 {                       // RFC 1321
  uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
  uint64_t total = 0;
  uint8_t block[64];
  size_t fill = 0;
  static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
  void compress(const uint8_t* p) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
        0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
        0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
        0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
        0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
        0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
        0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                              14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                              4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t m[16];
    for (int i = 0; i < 16; ++i)
      m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
    uint32_t A = a, B = b, C = c, D = d;
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      if (i < 16) { f = (B & C) | (~B & D); g = i; }
      else if (i < 32) { f = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
      else if (i < 48) { f = B ^ C ^ D; g = (3 * i + 5) & 15; }
      else { f = C ^ (B | ~D); g = (7 * i) & 15; }
      const uint32_t t = D;
      D = C;
      C = B;
      B = B + rol(A + f + K[i] + m[g], S[i]);
      A = t;
    }
    a += A; b += B; c += C; d += D;
  }
  void update(const uint8_t* p, size_t n) {
    total += n;
    while (n) {
      const size_t take = n < 64 - fill ? n : 64 - fill;
      memcpy(block + fill, p, take);
      fill += take; p += take; n -= take;
      if (fill == 64) { compress(block); fill = 0; }
    }
  }
  void finish(uint8_t out[16]) {
    const uint64_t bits = total * 8;
    const uint8_t one = 0x80, zero = 0;
    update(&one, 1);
    while (fill != 56) update(&zero, 1);
    uint8_t len[8];
    for (int i = 0; i < 8; ++i) len[i] = (uint8_t)(bits >> (8 * i));
    update(len, 8);
    const uint32_t v[4] = {a, b, c, d};
    for (int i = 0; i < 16; ++i) out[i] = (uint8_t)(v[i / 4] >> (8 * (i % 4)));
  }
};

// ------------------------------------------------------------------------------------------------ FLAC
struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;               // pos: next byte to load
  uint64_t acc = 0;                // unread bits, left-aligned
  int have = 0;
  bool bad = false;
  BitReader(const uint8_t* data, size_t len) : p(data), n(len) {}
  void refill() {
    while (have <= 56 && pos < n) { acc |= (uint64_t)p[pos++] << (56 - have); have += 8; }
  }
  uint32_t bits(int k) {           // 0 <= k <= 32
    if (k == 0) return 0;
    if (have < k) { refill(); if (have < k) { bad = true; return 0; } }
    const uint32_t v = (uint32_t)(acc >> (64 - k));
    acc <<= k; have -= k;
    return v;
  }
  int32_t sbits(int k) {           // two's complement, 1 <= k <= 32
    const uint32_t v = bits(k);
    if (k == 32) return (int32_t)v;
    const uint32_t m = 1u << (k - 1);
    return (int32_t)((v ^ m) - m);
  }
  int64_t sbits_wide(int k) {      // up to 33 bits (side channel of 32-bit audio)
    if (k <= 32) return sbits(k);
    const uint64_t hi = bits(k - 32), lo = bits(32);
    const uint64_t v = (hi << 32) | lo, m = 1ull << (k - 1);
    return (int64_t)((v ^ m) - m);
  }
  uint32_t unary() {               // number of 0 bits before the next 1 bit
    uint32_t q = 0;
    for (;;) {
      if (have == 0) { refill(); if (have == 0) { bad = true; return 0; } }
      if (acc == 0) { q += (uint32_t)have; have = 0; continue; }
      const int z = __builtin_clzll(acc);
      if (z >= have) { q += (uint32_t)have; acc = 0; have = 0; continue; }
      q += (uint32_t)z;
      acc = z == 63 ? 0 : acc << (z + 1);
      have -= z + 1;
      return q;
    }
  }
  void align() { const int r = have & 7; acc <<= r; have -= r; }
  size_t byte_pos() const { return pos - (size_t)(have >> 3); }      // valid when aligned
};

struct StreamInfo {
  int min_block = 0, max_block = 0, sample_rate = 0, channels = 0, bits = 0;
  long long total = 0;
  uint8_t md5[16];
  bool has_md5 = false;
};

static int parse_flac_header(const std::vector<uint8_t>& buf, StreamInfo& si, size_t& audio_start) {
  size_t o = 0;
  if (buf.size() >= 10 && memcmp(buf.data(), "ID3", 3) == 0) {          // an ID3v2 tag some taggers prepend
    o = 10 + (((size_t)buf[6] & 127) << 21 | ((size_t)buf[7] & 127) << 14 | ((size_t)buf[8] & 127) << 7 | ((size_t)buf[9] & 127));
  }
  if (buf.size() < o + 4 + 4 + 34 || memcmp(buf.data() + o, "fLaC", 4) != 0) return AIR_ERR_FORMAT;
  o += 4;
  bool last = false, seen_info = false;
  while (!last) {
    if (o + 4 > buf.size()) return AIR_ERR_FORMAT;
    last = (buf[o] & 0x80) != 0;
    const int type = buf[o] & 0x7f;
    const size_t len = ((size_t)buf[o + 1] << 16) | ((size_t)buf[o + 2] << 8) | buf[o + 3];
    o += 4;
    if (o + len > buf.size() || type == 127) return AIR_ERR_FORMAT;
    if (type == 0) {
      if (len != 34 || seen_info) return AIR_ERR_FORMAT;
      const uint8_t* s = buf.data() + o;
      si.min_block = (s[0] << 8) | s[1];
      si.max_block = (s[2] << 8) | s[3];
      si.sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
      si.channels = ((s[12] >> 1) & 7) + 1;
      si.bits = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      si.total = ((long long)(s[13] & 15) << 32) | ((long long)s[14] << 24) | (s[15] << 16) | (s[16] << 8) | s[17];
      memcpy(si.md5, s + 18, 16);
      si.has_md5 = false;
      for (int i = 0; i < 16; ++i) si.has_md5 |= si.md5[i] != 0;
      seen_info = true;
    } else if (!seen_info) {
      return AIR_ERR_FORMAT;                                            // STREAMINFO must come first
    }
    o += len;
  }
  if (!seen_info || si.sample_rate == 0 || si.bits < 4 || si.bits > 32) return AIR_ERR_FORMAT;
  audio_start = o;
  return AIR_OK;
}

// Partitioned Rice residual of one subframe: fills res[order .. block).
static int read_residual(BitReader& br, int block, int order, int32_t* res) {
  const int method = (int)br.bits(2);
  if (method > 1) return AIR_ERR_FORMAT;
  const int pbits = method == 0 ? 4 : 5, escape = method == 0 ? 15 : 31;
  const int porder = (int)br.bits(4);
  const int parts = 1 << porder;
  if ((block & (parts - 1)) != 0 && porder != 0) return AIR_ERR_FORMAT;
  if ((block >> porder) < order) return AIR_ERR_FORMAT;
  int i = order;
  for (int p = 0; p < parts; ++p) {
    const int count = (block >> porder) - (p == 0 ? order : 0);
    const int k = (int)br.bits(pbits);
    if (k == escape) {
      const int raw = (int)br.bits(5);
      for (int j = 0; j < count; ++j) res[i++] = raw ? br.sbits(raw) : 0;
    } else {
      for (int j = 0; j < count; ++j) {
        const uint32_t q = br.unary();
        const uint64_t u = ((uint64_t)q << k) | br.bits(k);
        if (u > 0xffffffffull) return AIR_ERR_FORMAT;
        res[i++] = (int32_t)((uint32_t)(u >> 1) ^ (0u - (uint32_t)(u & 1)));
      }
    }
    if (br.bad) return AIR_ERR_FORMAT;
  }
  return AIR_OK;
}

// One subframe into out[0 .. block) as 64-bit samples (a 33-bit side channel must not wrap).
static int read_subframe(BitReader& br, int block, int bps, int64_t* out, std::vector<int32_t>& res) {
  if (br.bits(1) != 0) return AIR_ERR_FORMAT;
  const int type = (int)br.bits(6);
  int wasted = 0;
  if (br.bits(1)) wasted = (int)br.unary() + 1;
  if (br.bad || wasted >= bps) return AIR_ERR_FORMAT;
  bps -= wasted;
  if (type == 0) {                                                      // CONSTANT
    const int64_t v = br.sbits_wide(bps);
    for (int i = 0; i < block; ++i) out[i] = v;
  } else if (type == 1) {                                               // VERBATIM
    for (int i = 0; i < block; ++i) out[i] = br.sbits_wide(bps);
  } else if (type >= 8 && type <= 12) {                                 // FIXED, order 0..4
    const int order = type - 8;
    if (order > block) return AIR_ERR_FORMAT;
    for (int i = 0; i < order; ++i) out[i] = br.sbits_wide(bps);
    res.resize((size_t)block);
    const int st = read_residual(br, block, order, res.data());
    if (st != AIR_OK) return st;
    // Sums wrap in uint64: a valid stream stays far inside 64 bits, a damaged one (caught by the CRC-16 afterwards)
    // must not run into signed-overflow UB on the way.
    uint64_t* u = reinterpret_cast<uint64_t*>(out);
    switch (order) {
      case 0: for (int i = 0; i < block; ++i) out[i] = res[i]; break;
      case 1: for (int i = 1; i < block; ++i) u[i] = (uint64_t)(int64_t)res[i] + u[i - 1]; break;
      case 2: for (int i = 2; i < block; ++i) u[i] = (uint64_t)(int64_t)res[i] + 2 * u[i - 1] - u[i - 2]; break;
      case 3: for (int i = 3; i < block; ++i) u[i] = (uint64_t)(int64_t)res[i] + 3 * u[i - 1] - 3 * u[i - 2] + u[i - 3]; break;
      default: for (int i = 4; i < block; ++i) u[i] = (uint64_t)(int64_t)res[i] + 4 * u[i - 1] - 6 * u[i - 2] + 4 * u[i - 3] - u[i - 4];
    }
  } else if (type >= 32) {                                              // LPC, order 1..32
    const int order = type - 31;
    if (order > block) return AIR_ERR_FORMAT;
    for (int i = 0; i < order; ++i) out[i] = br.sbits_wide(bps);
    const int prec = (int)br.bits(4) + 1;
    if (prec == 16) return AIR_ERR_FORMAT;
    const int shift = br.sbits(5);
    if (shift < 0) return AIR_ERR_FORMAT;
    int32_t coef[32];
    for (int j = 0; j < order; ++j) coef[j] = br.sbits(prec);
    res.resize((size_t)block);
    const int st = read_residual(br, block, order, res.data());
    if (st != AIR_OK) return st;
    for (int i = order; i < block; ++i) {
      uint64_t acc = 0;                                                 // wraps, see the FIXED predictor above
      for (int j = 0; j < order; ++j) acc += (uint64_t)(int64_t)coef[j] * (uint64_t)out[i - 1 - j];
      out[i] = (int64_t)((uint64_t)(int64_t)res[i] + (uint64_t)((int64_t)acc >> shift));
    }
  } else {
    return AIR_ERR_FORMAT;                                              // reserved subframe type
  }
  if (br.bad) return AIR_ERR_FORMAT;
  if (wasted) for (int i = 0; i < block; ++i) out[i] = (int64_t)((uint64_t)out[i] << wasted);
  return AIR_OK;
}

// Buffers a worker thread reuses from file to file: fresh multi-hundred-KB allocations per file are mmap / munmap /
// page-fault traffic that serialises the decoder threads in the kernel.
struct Scratch {
  std::vector<uint8_t> file, md5_bytes;
  std::vector<int64_t> ch[8];
  std::vector<int32_t> res;
};

static int decode_flac(const std::vector<uint8_t>& buf, Pcm& pcm, bool verify_md5, Scratch& sc) {
  init_tables();
  StreamInfo si;
  size_t o = 0;
  int st = parse_flac_header(buf, si, o);
  if (st != AIR_OK) return st;
  pcm.sample_rate = si.sample_rate; pcm.channels = si.channels; pcm.bits = si.bits;
  pcm.data.clear();
  // Reserve what the header announces only as far as the file could plausibly hold it (a damaged 36-bit total must
  // not become a 100 GB allocation); beyond that the vector grows as frames arrive.
  {
    const unsigned long long announced = (unsigned long long)si.total * (unsigned)si.channels;
    const unsigned long long plausible = (unsigned long long)buf.size() * 16ull + 65536ull;
    if (si.total > 0) pcm.data.reserve((size_t)(announced < plausible ? announced : plausible));
  }
  std::vector<int64_t>* ch = sc.ch;
  std::vector<int32_t>& res = sc.res;
  long long frames = 0;
  const uint8_t* base = buf.data();
  while (o + 2 <= buf.size()) {
    if (!(base[o] == 0xff && (base[o + 1] & 0xfe) == 0xf8)) {
      // trailing bytes that are not a frame (padding, ID3v1 tag): tolerated only after all announced samples
      if (si.total > 0 && frames >= si.total) break;
      return AIR_ERR_FORMAT;
    }
    BitReader br(base + o, buf.size() - o);
    br.bits(15);                                                        // sync + reserved 0
    br.bits(1);                                                         // blocking strategy: only changes what the number means
    const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4);
    const int ch_code = (int)br.bits(4), sz_code = (int)br.bits(3);
    if (br.bits(1) != 0) return AIR_ERR_FORMAT;
    {                                                                   // UTF-8 style frame / sample number (1..7 bytes)
      const uint32_t first = br.bits(8);
      int extra = 0;
      if (first >= 0xfe) extra = 6;
      else if (first >= 0xfc) extra = 5;
      else if (first >= 0xf8) extra = 4;
      else if (first >= 0xf0) extra = 3;
      else if (first >= 0xe0) extra = 2;
      else if (first >= 0xc0) extra = 1;
      else if (first >= 0x80) return AIR_ERR_FORMAT;
      for (int i = 0; i < extra; ++i) if ((br.bits(8) & 0xc0) != 0x80) return AIR_ERR_FORMAT;
    }
    int block;
    if (bs_code == 0) return AIR_ERR_FORMAT;
    else if (bs_code == 1) block = 192;
    else if (bs_code <= 5) block = 576 << (bs_code - 2);
    else if (bs_code == 6) block = (int)br.bits(8) + 1;
    else if (bs_code == 7) block = (int)br.bits(16) + 1;
    else block = 256 << (bs_code - 8);
    if (sr_code == 12) br.bits(8);
    else if (sr_code == 13 || sr_code == 14) br.bits(16);
    else if (sr_code == 15) return AIR_ERR_FORMAT;
    static const int size_of_code[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    int bps = size_of_code[sz_code];
    if (bps < 0) return AIR_ERR_FORMAT;
    if (bps == 0) bps = si.bits;
    if (bps != si.bits) return AIR_ERR_UNSUPPORTED;                     // sample size changing mid-stream
    int nch;
    if (ch_code < 8) nch = ch_code + 1;
    else if (ch_code <= 10) nch = 2;
    else return AIR_ERR_FORMAT;
    if (nch != si.channels) return AIR_ERR_UNSUPPORTED;
    if (br.bad) return AIR_ERR_FORMAT;
    const size_t hdr_len = br.byte_pos();
    const uint32_t want8 = br.bits(8);
    if (br.bad || crc8(base + o, hdr_len) != want8) return AIR_ERR_CHECKSUM;
    for (int c = 0; c < nch; ++c) {
      ch[c].resize((size_t)block);
      const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
      st = read_subframe(br, block, bps + (side ? 1 : 0), ch[c].data(), res);
      if (st != AIR_OK) return st;
    }
    br.align();
    const size_t body_len = br.byte_pos();
    const uint32_t want16 = br.bits(16);
    if (br.bad || crc16(base + o, body_len) != want16) return AIR_ERR_CHECKSUM;
    if (ch_code == 8) for (int i = 0; i < block; ++i) ch[1][i] = (int64_t)((uint64_t)ch[0][i] - (uint64_t)ch[1][i]);
    else if (ch_code == 9) for (int i = 0; i < block; ++i) ch[0][i] = (int64_t)((uint64_t)ch[1][i] + (uint64_t)ch[0][i]);
    else if (ch_code == 10)
      for (int i = 0; i < block; ++i) {
        const uint64_t side = (uint64_t)ch[1][i], mid = ((uint64_t)ch[0][i] << 1) | (side & 1);
        ch[0][i] = (int64_t)(mid + side) >> 1;
        ch[1][i] = (int64_t)(mid - side) >> 1;
      }
    const size_t at = pcm.data.size();
    pcm.data.resize(at + (size_t)block * nch);
    for (int i = 0; i < block; ++i)
      for (int c = 0; c < nch; ++c) pcm.data[at + (size_t)i * nch + c] = (int32_t)ch[c][i];
    frames += block;
    if ((si.total > 0 && frames > si.total) || frames > 0x7fffffffll) return AIR_ERR_FORMAT;
    o += body_len + 2;
  }
  if (si.total > 0 && frames != si.total) return AIR_ERR_FORMAT;
  pcm.frames = frames;
  if (verify_md5 && si.has_md5) {
    Md5 h;
    const int bytes = (si.bits + 7) / 8;
    std::vector<uint8_t>& tmp = sc.md5_bytes;
    tmp.clear();
    tmp.reserve(65536 + 8);
    for (size_t i = 0; i < pcm.data.size(); ++i) {
      const uint32_t v = (uint32_t)pcm.data[i];
      for (int b = 0; b < bytes; ++b) tmp.push_back((uint8_t)(v >> (8 * b)));
      if (tmp.size() >= 65536) { h.update(tmp.data(), tmp.size()); tmp.clear(); }
    }
    h.update(tmp.data(), tmp.size());
    uint8_t got[16];
    h.finish(got);
    if (memcmp(got, si.md5, 16) != 0) return AIR_ERR_CHECKSUM;
  }
  return AIR_OK;
}

// ------------------------------------------------------------------------------------------------ RIFF / WAVE
static uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// WAV keeps float data as float: `fdata` is used instead of pcm.data when the file holds IEEE samples.
static int decode_wav(const std::vector<uint8_t>& buf, Pcm& pcm, std::vector<float>& fdata) {
  if (buf.size() < 12 || memcmp(buf.data(), "RIFF", 4) != 0 || memcmp(buf.data() + 8, "WAVE", 4) != 0) return AIR_ERR_FORMAT;
  size_t o = 12;
  int format = 0, block_align = 0;
  bool have_fmt = false;
  while (o + 8 <= buf.size()) {
    const uint32_t len = le32(buf.data() + o + 4);
    const uint8_t* body = buf.data() + o + 8;
    const size_t avail = buf.size() - (o + 8);
    if (memcmp(buf.data() + o, "fmt ", 4) == 0) {
      if (len < 16 || avail < 16) return AIR_ERR_FORMAT;
      format = le16(body);
      pcm.channels = le16(body + 2);
      pcm.sample_rate = (int)le32(body + 4);
      block_align = le16(body + 12);
      pcm.bits = le16(body + 14);
      if (format == 0xfffe && len >= 26 && avail >= 26) format = le16(body + 24);     // WAVE_FORMAT_EXTENSIBLE
      have_fmt = true;
    } else if (memcmp(buf.data() + o, "data", 4) == 0) {
      if (!have_fmt || pcm.channels < 1 || pcm.channels > 8) return AIR_ERR_FORMAT;
      const int bytes = pcm.bits / 8;
      if (bytes < 1 || bytes > 4 || pcm.bits % 8 || block_align != bytes * pcm.channels) return AIR_ERR_UNSUPPORTED;
      size_t n = len <= avail ? len : avail;                            // a streamed file may claim 0xffffffff
      const size_t total = n / (size_t)bytes;
      pcm.frames = (long long)(total / pcm.channels);
      if (format == 1) {
        pcm.data.resize(total);
        for (size_t i = 0; i < total; ++i) {
          const uint8_t* s = body + i * bytes;
          int32_t v;
          if (bytes == 1) v = (int32_t)s[0] - 128;                      // 8-bit WAV is unsigned
          else if (bytes == 2) v = (int16_t)le16(s);
          else if (bytes == 3) v = (int32_t)((uint32_t)s[0] << 8 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 24) >> 8;
          else v = (int32_t)le32(s);
          pcm.data[i] = v;
        }
      } else if (format == 3 && bytes == 4) {
        fdata.resize(total);
        memcpy(fdata.data(), body, total * 4);
      } else {
        return AIR_ERR_UNSUPPORTED;
      }
      return AIR_OK;
    }
    o += 8 + (size_t)len + (len & 1);
  }
  return AIR_ERR_FORMAT;
}

// ------------------------------------------------------------------------------------------------ file -> mono float
struct Decoded {
  Pcm pcm;
  std::vector<float> fdata;       // interleaved float samples when the container held floats
};

static int decode_any(const char* path, Decoded& d, int flags, Scratch& sc) {
  try {                                                                 // no exception crosses the C boundary
    std::vector<uint8_t>& buf = sc.file;
    d.fdata.clear();
    d.pcm.data.clear();
    int st = read_file(path, buf);
    if (st != AIR_OK) return st;
    if (buf.size() >= 4 && memcmp(buf.data(), "RIFF", 4) == 0) return decode_wav(buf, d.pcm, d.fdata);
    return decode_flac(buf, d.pcm, (flags & 1) != 0, sc);
  } catch (const std::bad_alloc&) {
    return AIR_ERR_NOMEM;
  } catch (...) {
    return AIR_ERR_FORMAT;
  }
}
static int decode_any(const char* path, Decoded& d, int flags) {
  Scratch sc;
  return decode_any(path, d, flags, sc);
}

// mono = mean over channels (librosa.load(mono=True)); integer PCM scaled by 2^-(bits-1) (libsndfile)
static void to_mono_f32(const Decoded& d, float* out, long long n) {
  const int C = d.pcm.channels;
  if (!d.fdata.empty()) {
    for (long long i = 0; i < n; ++i) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += d.fdata[(size_t)i * C + c];
      out[i] = C == 1 ? s : s / (float)C;
    }
    return;
  }
  const float scale = 1.0f / (float)(1ll << (d.pcm.bits - 1));
  for (long long i = 0; i < n; ++i) {
    if (C == 1) {
      out[i] = (float)d.pcm.data[(size_t)i] * scale;
    } else {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += (float)d.pcm.data[(size_t)i * C + c] * scale;
      out[i] = s / (float)C;
    }
  }
}

}  // namespace air_audio

using namespace air_audio;

// Header-only when the container announces its length (FLAC STREAMINFO total samples, WAV data chunk size);
// a streamed FLAC without a total is decoded to count its frames.
extern "C" int air_audio_info(const char* path, int* sample_rate, int* channels, int* bits, long long* frames) {
  if (!path) return AIR_ERR_ARG;
  Pcm pcm;
  {
    FILE* f = fopen(path, "rb");
    if (!f) return AIR_ERR_IO;
    std::vector<uint8_t> head(1 << 16);
    head.resize(fread(head.data(), 1, head.size(), f));
    long long file_size = 0;
    if (fseek(f, 0, SEEK_END) == 0) file_size = ftell(f);
    fclose(f);
    bool done = false;
    if (head.size() >= 12 && memcmp(head.data(), "RIFF", 4) == 0) {
      size_t o = 12;
      int block_align = 0;
      while (o + 8 <= head.size() && !done) {
        const uint32_t len = le32(head.data() + o + 4);
        if (memcmp(head.data() + o, "fmt ", 4) == 0 && o + 24 <= head.size()) {
          pcm.channels = le16(head.data() + o + 10);
          pcm.sample_rate = (int)le32(head.data() + o + 12);
          block_align = le16(head.data() + o + 20);
          pcm.bits = le16(head.data() + o + 22);
        } else if (memcmp(head.data() + o, "data", 4) == 0 && block_align > 0) {
          const long long avail = file_size - (long long)(o + 8);      // what the file really holds (as the decoder does)
          pcm.frames = ((long long)len < avail ? (long long)len : (avail > 0 ? avail : 0)) / block_align;
          done = true;
        }
        o += 8 + (size_t)len + (len & 1);
      }
    } else {
      StreamInfo si;
      size_t start = 0;
      // the metadata may be longer than the 64 KB read here (cover art): then fall through to the full decode
      // ... and so does a total no file of this size can hold (a 65 535-sample frame takes >= 14 bytes), which a
      // caller would otherwise turn into an allocation: only the decoder's own count is trusted then
      if (parse_flac_header(head, si, start) == AIR_OK && si.total > 0 && si.total <= file_size * 4700 + 65536) {
        pcm.sample_rate = si.sample_rate; pcm.channels = si.channels; pcm.bits = si.bits; pcm.frames = si.total;
        done = true;
      }
    }
    if (!done) {
      Decoded d;
      const int st = decode_any(path, d, 0);
      if (st != AIR_OK) return st;
      pcm.sample_rate = d.pcm.sample_rate; pcm.channels = d.pcm.channels; pcm.bits = d.pcm.bits; pcm.frames = d.pcm.frames;
    }
  }
  if (sample_rate) *sample_rate = pcm.sample_rate;
  if (channels) *channels = pcm.channels;
  if (bits) *bits = pcm.bits;
  if (frames) *frames = pcm.frames;
  return AIR_OK;
}

// flags bit 0: verify the FLAC MD5 signature.  Writes min(frames, capacity) samples; *frames = the file's length.
extern "C" int air_audio_decode_f32(const char* path, float* out, long long capacity, long long* frames,
                                    int* sample_rate, int flags) {
  if (!path || !out || capacity < 0 || !frames) return AIR_ERR_ARG;
  Decoded d;
  const int st = decode_any(path, d, flags);
  if (st != AIR_OK) return st;
  *frames = d.pcm.frames;
  if (sample_rate) *sample_rate = d.pcm.sample_rate;
  to_mono_f32(d, out, d.pcm.frames < capacity ? d.pcm.frames : capacity);
  return AIR_OK;
}

// Interleaved integer samples exactly as stored (tests and tools): out has capacity int32 values.
extern "C" int air_audio_decode_i32(const char* path, int* out, long long capacity, long long* frames, int* channels,
                                    int* bits, int* sample_rate, int flags) {
  if (!path || !out || capacity < 0 || !frames) return AIR_ERR_ARG;
  Decoded d;
  const int st = decode_any(path, d, flags);
  if (st != AIR_OK) return st;
  if (!d.fdata.empty()) return AIR_ERR_UNSUPPORTED;
  *frames = d.pcm.frames;
  if (channels) *channels = d.pcm.channels;
  if (bits) *bits = d.pcm.bits;
  if (sample_rate) *sample_rate = d.pcm.sample_rate;
  const long long total = d.pcm.frames * d.pcm.channels;
  memcpy(out, d.pcm.data.data(), (size_t)(total < capacity ? total : capacity) * sizeof(int32_t));
  return AIR_OK;
}

// n files -> rows of a (pinned) float matrix with row stride ld: row i holds min(length, ld) mono samples followed by
// zeros; lengths[i] = the file's own length (may exceed ld), sample_rates[i] its rate, status[i] its error code.
// Decoding runs on `threads` host threads (<= 0: hardware concurrency).  Returns the first non-zero status.
extern "C" int air_audio_decode_batch_f32(const char* const* paths, int n, float* out, long long ld, int* lengths,
                                          int* sample_rates, int* status, int threads, int flags) {
  if (!paths || n < 0 || !out || ld < 1 || !lengths) return AIR_ERR_ARG;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  std::vector<int> st((size_t)n, AIR_OK);
  std::atomic<int> next{0};
  auto work = [&]() {
    Decoded d;
    Scratch sc;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) return;
      float* row = out + (long long)i * ld;
      st[i] = paths[i] ? decode_any(paths[i], d, flags, sc) : AIR_ERR_ARG;
      long long got = 0;
      if (st[i] == AIR_OK) {
        got = d.pcm.frames < ld ? d.pcm.frames : ld;
        to_mono_f32(d, row, got);
        lengths[i] = (int)(d.pcm.frames > 0x7fffffffll ? 0x7fffffffll : d.pcm.frames);
        if (sample_rates) sample_rates[i] = d.pcm.sample_rate;
      } else {
        lengths[i] = 0;
        if (sample_rates) sample_rates[i] = 0;
      }
      memset(row + got, 0, (size_t)(ld - got) * sizeof(float));
    }
  };
  if (threads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  int first = AIR_OK;
  for (int i = 0; i < n; ++i) {
    if (status) status[i] = st[i];
    if (first == AIR_OK && st[i] != AIR_OK) first = st[i];
  }
  return first;
}

// Rows of a (pinned) float matrix from a packed int16 corpus (data.PackedWaves): row i = blob[offsets[i] ..
// offsets[i] + min(lengths[i], ld)) * 2^-15, zero-padded to ld.  `blob_samples` bounds every (offset, length) pair -- a
// stale index must not read past the mapping.  Memory-bound; `threads` host threads (<= 0: all, capped at 16).
extern "C" int air_audio_gather_i16_f32(const short* blob, long long blob_samples, const long long* offsets,
                                        const int* lengths, int n, float* out, long long ld, int threads) {
  if (!blob || blob_samples < 0 || !offsets || !lengths || n < 0 || !out || ld < 1) return AIR_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (offsets[i] < 0 || lengths[i] < 0 || offsets[i] > blob_samples || lengths[i] > blob_samples - offsets[i]) return AIR_ERR_ARG;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads > 16) threads = 16;
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  std::atomic<int> next{0};
  auto work = [&]() {
    const float scale = 1.0f / 32768.0f;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) return;
      const short* src = blob + offsets[i];
      float* row = out + (long long)i * ld;
      const long long k = lengths[i] < ld ? lengths[i] : ld;
      for (long long j = 0; j < k; ++j) row[j] = (float)src[j] * scale;
      memset(row + k, 0, (size_t)(ld - k) * sizeof(float));
    }
  };
  if (threads > 1) {
    std::vector<std::thread> pool;
    try {
      for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    } catch (...) {
      // thread creation failed (resource limits): the calling thread and the workers that did start drain the queue
    }
    work();
    for (auto& t : pool) t.join();
  } else {
    work();
  }
  return AIR_OK;
}
