// rt_png.cpp — PNG reader / writer with its own inflate (no zlib dependency), for compiled callers that do not
// have the `image` crate the reference uses (texture.rs:17 `image::open`, tracing.rs:546 `save_with_format`).
// Reader: non-interlaced, colour types 0/2/3/4/6, bit depths 1/2/4/8/16 -> RGB8, row 0 = top (to_rgb() semantics:
// grey replicated, alpha dropped, 16-bit truncated to the high byte).  Writer: RGB8, stored (uncompressed) deflate.
// Host-side asset code, never on the per-ray path.

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rt_lower.h"

namespace rt {
namespace {

// ---------------------------------------------------------------- inflate (RFC 1951)
struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;
  uint32_t bitbuf = 0;
  int bitcnt = 0;
  bool fail = false;
  int bits(int need) {
    uint32_t v = bitbuf;
    while (bitcnt < need) {
      if (pos >= n) {
        fail = true;
        return 0;
      }
      v |= (uint32_t)p[pos++] << bitcnt;
      bitcnt += 8;
    }
    bitbuf = v >> need;
    bitcnt -= need;
    return (int)(v & ((1u << need) - 1u));
  }
};
struct Huffman {
  uint16_t count[16];
  uint16_t symbol[288];
  void build(const uint8_t* length, int n) {
    std::memset(count, 0, sizeof count);
    for (int i = 0; i < n; ++i) count[length[i]]++;
    count[0] = 0;
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; ++len) offs[len + 1] = offs[len] + count[len];
    for (int i = 0; i < n; ++i)
      if (length[i]) symbol[offs[length[i]]++] = (uint16_t)i;
  }
  int decode(BitReader& br) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
      code |= br.bits(1);
      if (br.fail) return -1;
      int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};
const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint16_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint16_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflate_codes(BitReader& br, std::vector<uint8_t>& out, const Huffman& lc, const Huffman& dc, size_t max_out) {
  for (;;) {
    int sym = lc.decode(br);
    if (sym < 0) return false;
    if (out.size() > max_out) return false;  // more data than the image can hold: refuse (zip bomb)
    if (sym < 256) {
      out.push_back((uint8_t)sym);
    } else if (sym == 256) {
      return true;
    } else {
      sym -= 257;
      if (sym >= 29) return false;
      int len = LBASE[sym] + br.bits(LEXT[sym]);
      int ds = dc.decode(br);
      if (ds < 0 || ds >= 30) return false;
      size_t dist = (size_t)DBASE[ds] + (size_t)br.bits(DEXT[ds]);
      if (br.fail || dist > out.size()) return false;
      size_t from = out.size() - dist;
      for (int i = 0; i < len; ++i) out.push_back(out[from + i]);
    }
  }
}

// `max_out`: the caller knows how many bytes the stream may legitimately produce; anything beyond that is refused
bool inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t max_out) {
  BitReader br{src, n};
  int last;
  do {
    last = br.bits(1);
    int type = br.bits(2);
    if (br.fail) return false;
    if (type == 0) {
      br.bitbuf = 0;
      br.bitcnt = 0;
      if (br.pos + 4 > n) return false;
      uint32_t len = src[br.pos] | (src[br.pos + 1] << 8), nlen = src[br.pos + 2] | (src[br.pos + 3] << 8);
      br.pos += 4;
      if ((len ^ 0xFFFFu) != nlen || br.pos + len > n || out.size() + len > max_out + 258) return false;
      out.insert(out.end(), src + br.pos, src + br.pos + len);
      br.pos += len;
    } else if (type == 1) {
      uint8_t l[288];
      for (int i = 0; i < 144; ++i) l[i] = 8;
      for (int i = 144; i < 256; ++i) l[i] = 9;
      for (int i = 256; i < 280; ++i) l[i] = 7;
      for (int i = 280; i < 288; ++i) l[i] = 8;
      Huffman lc, dc;
      lc.build(l, 288);
      uint8_t d[30];
      for (int i = 0; i < 30; ++i) d[i] = 5;
      dc.build(d, 30);
      if (!inflate_codes(br, out, lc, dc, max_out)) return false;
    } else if (type == 2) {
      int nlen = br.bits(5) + 257, ndist = br.bits(5) + 1, ncode = br.bits(4) + 4;
      if (br.fail || nlen > 286 || ndist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t lengths[320];
      std::memset(lengths, 0, sizeof lengths);
      for (int i = 0; i < ncode; ++i) lengths[order[i]] = (uint8_t)br.bits(3);
      Huffman cl;
      cl.build(lengths, 19);
      uint8_t ll[320];
      std::memset(ll, 0, sizeof ll);
      int idx = 0;
      while (idx < nlen + ndist) {
        int sym = cl.decode(br);
        if (sym < 0) return false;
        if (sym < 16) {
          ll[idx++] = (uint8_t)sym;
        } else {
          int prev = 0, rep;
          if (sym == 16) {
            if (idx == 0) return false;
            prev = ll[idx - 1];
            rep = 3 + br.bits(2);
          } else if (sym == 17) {
            rep = 3 + br.bits(3);
          } else {
            rep = 11 + br.bits(7);
          }
          if (br.fail || idx + rep > nlen + ndist) return false;
          while (rep--) ll[idx++] = (uint8_t)prev;
        }
      }
      if (ll[256] == 0) return false;
      Huffman lc, dc;
      lc.build(ll, nlen);
      dc.build(ll + nlen, ndist);
      if (!inflate_codes(br, out, lc, dc, max_out)) return false;
    } else {
      return false;
    }
  } while (!last);
  return true;
}

uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32(const uint8_t* p, size_t n, uint32_t c = 0) {
  if (!crc_ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t v = i;
      for (int k = 0; k < 8; ++k) v = (v & 1) ? 0xEDB88320u ^ (v >> 1) : v >> 1;
      crc_table[i] = v;
    }
    crc_ready = true;
  }
  c ^= 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 255] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
inline void put32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
inline int paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

int png_decode(const uint8_t* b, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h, std::string& err) {
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (len < 33 || std::memcmp(b, sig, 8) != 0) { err = "not a PNG"; return RT_ERR_IO; }
  uint32_t W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  size_t pos = 8;
  bool end = false;
  while (!end && pos + 12 <= len) {
    uint32_t n = be32(b + pos);
    const uint8_t* type = b + pos + 4;
    const uint8_t* data = b + pos + 8;
    if (pos + 12 + (size_t)n > len) { err = "PNG truncated"; return RT_ERR_IO; }
    if (!std::memcmp(type, "IHDR", 4) && n >= 13) {
      W = be32(data); H = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
    } else if (!std::memcmp(type, "PLTE", 4)) {
      plte.assign(data, data + n);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + n);
    } else if (!std::memcmp(type, "IEND", 4)) {
      end = true;
    }
    pos += 12 + (size_t)n;
  }
  if (!W || !H || W > (1u << 16) || H > (1u << 16)) { err = "bad PNG header"; return RT_ERR_IO; }
  if (interlace) { err = "interlaced PNG is not supported"; return RT_ERR_IO; }
  int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!channels || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16) ||
      ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) || (ctype == 3 && depth == 16)) {
    err = "unsupported PNG colour type / bit depth";
    return RT_ERR_IO;
  }
  if (idat.size() < 6) { err = "PNG has no image data"; return RT_ERR_IO; }
  size_t bpp = std::max<size_t>(1, (size_t)channels * depth / 8);
  size_t stride = ((size_t)W * channels * depth + 7) / 8;
  const size_t expected = (stride + 1) * (size_t)H;
  // deflate expands by at most 1032:1, so a header that promises more pixels than the IDAT bytes can produce is
  // refused before anything is allocated for it
  if (expected / 1032 > idat.size()) { err = "PNG image data too short"; return RT_ERR_IO; }
  std::vector<uint8_t> raw;
  raw.reserve(expected + 258);
  if (!inflate(idat.data() + 2, idat.size() - 2, raw, expected)) { err = "PNG inflate failed"; return RT_ERR_IO; }
  if (raw.size() < (stride + 1) * (size_t)H) { err = "PNG image data too short"; return RT_ERR_IO; }
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  uint8_t* out = (uint8_t*)std::malloc((size_t)W * H * 3);
  if (!out) { err = "out of memory"; return RT_ERR_IO; }
  for (uint32_t y = 0; y < H; ++y) {
    const uint8_t* line = &raw[(stride + 1) * (size_t)y];
    uint8_t ft = line[0];
    for (size_t i = 0; i < stride; ++i) {
      int x = line[1 + i], a = i >= bpp ? cur[i - bpp] : 0, up = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
      switch (ft) {
        case 0: break;
        case 1: x += a; break;
        case 2: x += up; break;
        case 3: x += (a + up) >> 1; break;
        case 4: x += paeth(a, up, c); break;
        default: std::free(out); err = "bad PNG filter"; return RT_ERR_IO;
      }
      cur[i] = (uint8_t)x;
    }
    for (uint32_t x = 0; x < W; ++x) {
      uint8_t* d = out + ((size_t)y * W + x) * 3;
      auto sample = [&](uint32_t idx) -> uint32_t {  // idx-th sample of the scanline
        if (depth == 8) return cur[idx];
        if (depth == 16) return cur[2 * idx];
        uint32_t bit = idx * depth;
        return (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
      };
      if (ctype == 3) {
        uint32_t pi = sample(x);
        if (3 * pi + 2 < plte.size()) { d[0] = plte[3 * pi]; d[1] = plte[3 * pi + 1]; d[2] = plte[3 * pi + 2]; }
        else d[0] = d[1] = d[2] = 0;
      } else if (ctype == 0 || ctype == 4) {
        uint32_t g = sample(x * channels);
        if (depth < 8) g = g * 255u / ((1u << depth) - 1u);
        d[0] = d[1] = d[2] = (uint8_t)g;
      } else {
        d[0] = (uint8_t)sample(x * channels); d[1] = (uint8_t)sample(x * channels + 1); d[2] = (uint8_t)sample(x * channels + 2);
      }
    }
    prev.swap(cur);
  }
  *rgb = out;
  *w = W;
  *h = H;
  return RT_OK;
}

int png_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len) {
  if (!rgb || !w || !h) return RT_ERR_INVALID;
  std::vector<uint8_t> raw;
  raw.reserve((size_t)h * (3 * (size_t)w + 1));
  for (uint32_t y = 0; y < h; ++y) {
    raw.push_back(0);
    raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
  }
  std::vector<uint8_t> z;
  z.push_back(0x78);
  z.push_back(0x01);
  uint32_t a = 1, bsum = 0;
  for (size_t off = 0; off < raw.size() || off == 0;) {
    size_t n = std::min<size_t>(65535, raw.size() - off);
    bool last = off + n >= raw.size();
    z.push_back(last ? 1 : 0);
    z.push_back(n & 255); z.push_back(n >> 8); z.push_back(~n & 255); z.push_back((~n >> 8) & 255);
    z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    for (size_t i = 0; i < n; ++i) {
      a = (a + raw[off + i]) % 65521u;
      bsum = (bsum + a) % 65521u;
    }
    off += n;
    if (last) break;
  }
  put32(z, (bsum << 16) | a);
  std::vector<uint8_t> o = {137, 80, 78, 71, 13, 10, 26, 10};
  auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
    put32(o, (uint32_t)data.size());
    size_t start = o.size();
    o.insert(o.end(), type, type + 4);
    o.insert(o.end(), data.begin(), data.end());
    put32(o, crc32(o.data() + start, o.size() - start));
  };
  std::vector<uint8_t> ihdr;
  put32(ihdr, w);
  put32(ihdr, h);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk("IHDR", ihdr);
  chunk("IDAT", z);
  chunk("IEND", {});
  uint8_t* out = (uint8_t*)std::malloc(o.size());
  if (!out) return RT_ERR_IO;
  std::memcpy(out, o.data(), o.size());
  *bytes = out;
  *len = o.size();
  return RT_OK;
}

}  // namespace rt
