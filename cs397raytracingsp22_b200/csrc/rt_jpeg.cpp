// rt_jpeg.cpp — JPEG (ITU-T T.81) decoder for texture assets: baseline / extended sequential (SOF0, SOF1) and
// progressive (SOF2) Huffman, 8-bit, 1 or 3 components, any sampling factors, restart intervals, JFIF YCbCr and Adobe
// RGB.  Replaces what `image::open` (texture.rs:17) does for the reference's .jpg textures through jpeg-decoder 0.1.22
// (Cargo.lock:431); texture/magenta.jpg is progressive 4:2:0, normal_test.jpg baseline 4:2:0, earthmap.jpg 4:4:4.
// One-time asset work on the host, never in the per-ray loop.
//
// The algorithms are the standard's: Huffman decoding by code length (T.81 F.2.2.3), sequential and progressive
// coefficient decoding (F.2.2, G.1.2), dequantisation, a separable 8x8 inverse DCT in float (A.3.3) and the JFIF colour
// transform.  Chroma is upsampled with the triangle ("fancy") filter for 2:1 horizontal and 2x2 subsampling, by
// replication otherwise.  Decoders legitimately differ by an LSB or two in IDCT rounding and upsampling; the oracle and
// the GPU always consume the same decoded array, so that never enters a parity comparison.

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "rt_lower.h"

namespace rt {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman {
  bool present = false;
  uint8_t vals[256];
  int mincode[17], maxcode[18], valptr[17];
  uint8_t look_len[256], look_val[256];  // codes of up to 8 bits resolved in one step
  // false: the code lengths over-subscribe the code space (not a prefix code)
  bool build(const uint8_t counts[16], const uint8_t* symbols) {
    present = false;
    int code = 0, k = 0;
    std::memset(look_len, 0, sizeof look_len);
    for (int len = 1; len <= 16; ++len) {
      valptr[len] = k;
      mincode[len] = code;
      if (code + counts[len - 1] > (1 << len)) return false;
      for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
        vals[k] = symbols[k];
        if (len <= 8) {
          int first = code << (8 - len), n = 1 << (8 - len);
          for (int j = 0; j < n; ++j) {
            look_len[first + j] = (uint8_t)len;
            look_val[first + j] = symbols[k];
          }
        }
      }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7FFFFFFF;
    present = true;
    return true;
  }
};

// entropy-coded segment reader: removes the stuffed zero after 0xFF, stops feeding at a marker
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  int marker = 0;  // marker that ended the segment (0 while none seen)
  const uint8_t* marker_at = nullptr;  // its 0xFF
  void fill() {
    while (n <= 24) {
      int byte = 0;
      if (!marker && p < end) {
        byte = *p++;
        if (byte == 0xFF) {
          int next = p < end ? *p : 0xD9;
          if (next == 0) {
            ++p;
          } else {
            while (p < end && *p == 0xFF) ++p;  // fill bytes
            marker_at = p < end ? p - 1 : end;
            marker = p < end ? *p++ : 0xD9;
            byte = 0;
          }
        }
      } else if (!marker) {
        marker = 0xD9;
        marker_at = end;
      }
      acc |= (uint32_t)byte << (24 - n);
      n += 8;
    }
  }
  int peek(int k) {
    if (n < k) fill();
    return (int)(acc >> (32 - k));
  }
  void skip(int k) {
    acc <<= k;
    n -= k;
  }
  int bits(int k) {
    if (k == 0) return 0;
    int v = peek(k);
    skip(k);
    return v;
  }
  int bit() { return bits(1); }
};

inline int decode_symbol(BitReader& br, const Huffman& h) {
  int look = br.peek(8);
  if (h.look_len[look]) {
    br.skip(h.look_len[look]);
    return h.look_val[look];
  }
  int code = br.peek(16);
  for (int len = 9; len <= 16; ++len) {
    int c = code >> (16 - len);
    if (h.maxcode[len] >= 0 && c <= h.maxcode[len] && c >= h.mincode[len]) {
      br.skip(len);
      return h.vals[h.valptr[len] + c - h.mincode[len]];
    }
  }
  return -1;
}
// T.81 F.2.2.1 EXTEND
inline int extend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;        // Huffman table selectors of the current scan
  int bw = 0, bh = 0;        // blocks per row / column actually covering the component
  int pbw = 0, pbh = 0;      // padded to whole MCUs (buffer stride)
  int dc_pred = 0;
  std::vector<int16_t> coef;  // pbw * pbh * 64, natural order
  std::vector<uint8_t> plane;  // pbw*8 x pbh*8 samples
};

struct Decoder {
  const uint8_t* data;
  size_t len;
  std::string& err;
  int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  bool progressive = false, have_frame = false;
  int adobe_transform = -1;
  int restart_interval = 0;
  uint16_t qt[4][64];  // natural order
  bool qt_present[4] = {false, false, false, false};
  Huffman dc[4], ac[4];
  Component comp[3];
  int eobrun = 0;

  Decoder(const uint8_t* d, size_t l, std::string& e) : data(d), len(l), err(e) {}
  int bad(const char* why) {
    err = std::string("jpeg: ") + why;
    return RT_ERR_IO;
  }

  int parse_dqt(const uint8_t* p, int n) {
    while (n > 0) {
      int pq = p[0] >> 4, tq = p[0] & 15;
      if (tq > 3 || pq > 1) return bad("bad quantisation table");
      int need = 1 + 64 * (pq + 1);
      if (n < need) return bad("short DQT");
      for (int i = 0; i < 64; ++i) qt[tq][kZigzag[i]] = pq ? (uint16_t)((p[1 + 2 * i] << 8) | p[2 + 2 * i]) : p[1 + i];
      qt_present[tq] = true;
      p += need;
      n -= need;
    }
    return RT_OK;
  }
  int parse_dht(const uint8_t* p, int n) {
    while (n > 0) {
      if (n < 17) return bad("short DHT");
      int tc = p[0] >> 4, th = p[0] & 15;
      if (tc > 1 || th > 3) return bad("bad Huffman table id");
      int total = 0;
      for (int i = 0; i < 16; ++i) total += p[1 + i];
      if (total > 256 || n < 17 + total) return bad("bad Huffman table");
      if (!(tc ? ac[th] : dc[th]).build(p + 1, p + 17)) return bad("Huffman table is not a prefix code");
      p += 17 + total;
      n -= 17 + total;
    }
    return RT_OK;
  }
  int parse_sof(const uint8_t* p, int n, int marker) {
    if (have_frame) return bad("more than one frame");
    if (n < 6) return bad("short SOF");
    if (p[0] != 8) return bad("only 8-bit samples are supported");
    height = (p[1] << 8) | p[2];
    width = (p[3] << 8) | p[4];
    ncomp = p[5];
    if (width == 0 || height == 0) return bad("empty image");
    if (ncomp != 1 && ncomp != 3) return bad("only 1 or 3 components are supported");
    if (n < 6 + 3 * ncomp) return bad("short SOF");
    progressive = marker == 0xC2;
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      c.id = p[6 + 3 * i];
      c.h = p[7 + 3 * i] >> 4;
      c.v = p[7 + 3 * i] & 15;
      c.tq = p[8 + 3 * i];
      if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return bad("bad component");
      hmax = std::max(hmax, c.h);
      vmax = std::max(vmax, c.v);
    }
    if (ncomp == 1) comp[0].h = comp[0].v = hmax = vmax = 1;  // a single component is never interleaved
    mcux = (width + 8 * hmax - 1) / (8 * hmax);
    mcuy = (height + 8 * vmax - 1) / (8 * vmax);
    if ((uint64_t)mcux * mcuy * hmax * vmax * 64 * 3 > (1ull << 31)) return bad("image too large");
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
      c.bw = (cw + 7) / 8;
      c.bh = (ch + 7) / 8;
      c.pbw = mcux * c.h;
      c.pbh = mcuy * c.v;
      c.coef.assign((size_t)c.pbw * c.pbh * 64, 0);
    }
    have_frame = true;
    return RT_OK;
  }

  // ---- one 8x8 block of one scan
  int block_sequential(BitReader& br, Component& c, int16_t* b) {
    const Huffman& hd = dc[c.td];
    const Huffman& ha = ac[c.ta];
    int t = decode_symbol(br, hd);
    if (t < 0 || t > 15) return bad("bad DC code");
    c.dc_pred += extend(br.bits(t), t);
    b[0] = (int16_t)c.dc_pred;
    for (int k = 1; k < 64;) {
      int rs = decode_symbol(br, ha);
      if (rs < 0) return bad("bad AC code");
      int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r != 15) break;
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) return bad("AC run past the block");
      b[kZigzag[k++]] = (int16_t)extend(br.bits(s), s);
    }
    return RT_OK;
  }
  int block_dc_progressive(BitReader& br, Component& c, int16_t* b, int ah, int al) {
    if (ah == 0) {
      int t = decode_symbol(br, dc[c.td]);
      if (t < 0 || t > 15) return bad("bad DC code");
      c.dc_pred += extend(br.bits(t), t);
      b[0] = (int16_t)(c.dc_pred * (1 << al));
    } else if (br.bit()) {
      b[0] = (int16_t)(b[0] | (1 << al));
    }
    return RT_OK;
  }
  int block_ac_first(BitReader& br, Component& c, int16_t* b, int ss, int se, int al) {
    if (eobrun > 0) {
      --eobrun;
      return RT_OK;
    }
    const Huffman& ha = ac[c.ta];
    for (int k = ss; k <= se;) {
      int rs = decode_symbol(br, ha);
      if (rs < 0) return bad("bad AC code");
      int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r < 15) {
          eobrun = (1 << r) - 1;
          if (r) eobrun += br.bits(r);
          break;
        }
        k += 16;
        continue;
      }
      k += r;
      if (k > se) return bad("AC run past the band");
      b[kZigzag[k++]] = (int16_t)(extend(br.bits(s), s) * (1 << al));
    }
    return RT_OK;
  }
  // T.81 G.1.2.3: successive-approximation refinement of the AC band
  int block_ac_refine(BitReader& br, Component& c, int16_t* b, int ss, int se, int al) {
    const int p1 = 1 << al, m1 = -(1 << al);
    const Huffman& ha = ac[c.ta];
    int k = ss;
    auto refine = [&](int16_t& v) {
      if (br.bit() && (v & p1) == 0) v = (int16_t)(v + (v >= 0 ? p1 : m1));
    };
    if (eobrun == 0) {
      for (; k <= se; ++k) {
        int rs = decode_symbol(br, ha);
        if (rs < 0) return bad("bad AC code");
        int r = rs >> 4, s = rs & 15, value = 0;
        if (s) {
          value = br.bit() ? p1 : m1;  // the size of a newly significant coefficient is always 1
        } else if (r != 15) {
          eobrun = 1 << r;
          if (r) eobrun += br.bits(r);
          break;
        }
        // pass over r still-zero coefficients, correcting the already significant ones on the way
        for (; k <= se; ++k) {
          int16_t& v = b[kZigzag[k]];
          if (v != 0) {
            refine(v);
          } else if (--r < 0) {
            break;
          }
        }
        if (value && k <= se) b[kZigzag[k]] = (int16_t)value;
      }
    }
    if (eobrun > 0) {
      for (; k <= se; ++k) {
        int16_t& v = b[kZigzag[k]];
        if (v != 0) refine(v);
      }
      --eobrun;
    }
    return RT_OK;
  }

  int decode_scan(const uint8_t* hdr, int n, const uint8_t*& pos) {
    if (!have_frame) return bad("scan before frame");
    if (n < 1) return bad("short SOS");
    int ns = hdr[0];
    if (ns < 1 || ns > ncomp || n < 1 + 2 * ns + 3) return bad("bad SOS");
    Component* sc[3];
    for (int i = 0; i < ns; ++i) {
      int id = hdr[1 + 2 * i];
      sc[i] = nullptr;
      for (int j = 0; j < ncomp; ++j)
        if (comp[j].id == id) sc[i] = &comp[j];
      if (!sc[i]) return bad("scan names an unknown component");
      sc[i]->td = hdr[2 + 2 * i] >> 4;
      sc[i]->ta = hdr[2 + 2 * i] & 15;
      if (sc[i]->td > 3 || sc[i]->ta > 3) return bad("bad table selector");
    }
    int ss = hdr[1 + 2 * ns], se = hdr[2 + 2 * ns], ah = hdr[3 + 2 * ns] >> 4, al = hdr[3 + 2 * ns] & 15;
    if (!progressive) {
      ss = 0;
      se = 63;
      ah = al = 0;
    } else {
      if (ss > se || se > 63 || al > 13 || ah > 13) return bad("bad spectral selection");
      if (ss == 0 && se != 0) return bad("DC and AC in one progressive scan");
      if (ss > 0 && ns != 1) return bad("interleaved AC scan");
    }
    for (int i = 0; i < ns; ++i) {
      bool need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || ss > 0;
      if (need_dc && !dc[sc[i]->td].present) return bad("missing DC Huffman table");
      if (need_ac && !ac[sc[i]->ta].present) return bad("missing AC Huffman table");
    }
    BitReader br{pos, data + len};
    auto one_block = [&](Component& c, int bx, int by) -> int {
      int16_t* b = c.coef.data() + ((size_t)by * c.pbw + bx) * 64;
      if (!progressive) return block_sequential(br, c, b);
      if (ss == 0) return block_dc_progressive(br, c, b, ah, al);
      return ah == 0 ? block_ac_first(br, c, b, ss, se, al) : block_ac_refine(br, c, b, ss, se, al);
    };
    auto restart = [&]() -> int {
      // the entropy segment ends at a byte boundary with RSTn
      br.n = 0;
      br.acc = 0;
      if (!br.marker) br.fill();
      if (br.marker < 0xD0 || br.marker > 0xD7) return bad("missing restart marker");
      const uint8_t* resume = br.p;
      br = BitReader{resume, data + len};
      for (int j = 0; j < ncomp; ++j) comp[j].dc_pred = 0;
      eobrun = 0;
      return RT_OK;
    };
    for (int j = 0; j < ncomp; ++j) comp[j].dc_pred = 0;
    eobrun = 0;
    int rc, todo = restart_interval;
    if (ns == 1) {  // non-interleaved: the component's own blocks, row by row
      Component& c = *sc[0];
      for (int by = 0; by < c.bh; ++by)
        for (int bx = 0; bx < c.bw; ++bx) {
          if ((rc = one_block(c, bx, by)) != RT_OK) return rc;
          if (restart_interval && --todo == 0 && !(by == c.bh - 1 && bx == c.bw - 1)) {
            if ((rc = restart()) != RT_OK) return rc;
            todo = restart_interval;
          }
        }
    } else {
      for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
          for (int i = 0; i < ns; ++i)
            for (int vy = 0; vy < sc[i]->v; ++vy)
              for (int hx = 0; hx < sc[i]->h; ++hx)
                if ((rc = one_block(*sc[i], mx * sc[i]->h + hx, my * sc[i]->v + vy)) != RT_OK) return rc;
          if (restart_interval && --todo == 0 && !(my == mcuy - 1 && mx == mcux - 1)) {
            if ((rc = restart()) != RT_OK) return rc;
            todo = restart_interval;
          }
        }
    }
    // continue parsing at the marker that ends the segment
    if (br.marker) {
      pos = br.marker_at;
    } else {
      const uint8_t* q = br.p;
      while (q + 1 < data + len && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF)) ++q;
      pos = q + 1 < data + len ? q : data + len;
    }
    return RT_OK;
  }

  // dequantise + inverse DCT every block into the component planes
  void reconstruct() {
    float cs[8][8];  // cs[x][u] = C(u)/2 * cos((2x+1) u pi / 16)
    for (int x = 0; x < 8; ++x)
      for (int u = 0; u < 8; ++u)
        cs[x][u] = (float)((u == 0 ? std::sqrt(0.5) : 1.0) * 0.5 * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      const uint16_t* q = qt[c.tq];
      const int stride = c.pbw * 8;
      c.plane.assign((size_t)stride * c.pbh * 8, 0);
      for (int by = 0; by < c.pbh; ++by)
        for (int bx = 0; bx < c.pbw; ++bx) {
          const int16_t* b = c.coef.data() + ((size_t)by * c.pbw + bx) * 64;
          float f[64], t[64];
          for (int k = 0; k < 64; ++k) f[k] = (float)(b[k] * (int)q[k]);
          for (int y = 0; y < 8; ++y)  // rows: t[y][x] = sum_u cs[x][u] f[y][u]
            for (int x = 0; x < 8; ++x) {
              float s = 0.0f;
              for (int u = 0; u < 8; ++u) s += cs[x][u] * f[y * 8 + u];
              t[y * 8 + x] = s;
            }
          uint8_t* out = c.plane.data() + (size_t)by * 8 * stride + bx * 8;
          for (int x = 0; x < 8; ++x)  // columns
            for (int y = 0; y < 8; ++y) {
              float s = 0.0f;
              for (int v = 0; v < 8; ++v) s += cs[y][v] * t[v * 8 + x];
              int val = (int)std::lrintf(s) + 128;
              out[(size_t)y * stride + x] = (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
            }
        }
    }
  }

  // component plane -> full-resolution plane (width x height)
  void upsample(const Component& c, std::vector<uint8_t>& out) const {
    out.resize((size_t)width * height);
    const int stride = c.pbw * 8;
    const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
    const uint8_t* in = c.plane.data();
    if (c.h == hmax && c.v == vmax) {
      for (int y = 0; y < height; ++y) std::memcpy(&out[(size_t)y * width], in + (size_t)y * stride, (size_t)width);
    } else if (c.h * 2 == hmax && c.v == vmax) {  // 2:1 horizontally, triangle filter
      for (int y = 0; y < height; ++y) {
        const uint8_t* r = in + (size_t)y * stride;
        for (int x = 0; x < width; ++x) {
          int i = x >> 1, j = (x & 1) ? std::min(i + 1, cw - 1) : std::max(i - 1, 0);
          out[(size_t)y * width + x] = (uint8_t)((3 * r[i] + r[j] + ((x & 1) ? 2 : 1)) >> 2);
        }
      }
    } else if (c.h * 2 == hmax && c.v * 2 == vmax) {  // 2x2, triangle filter in both directions
      for (int y = 0; y < height; ++y) {
        int iy = y >> 1, jy = (y & 1) ? std::min(iy + 1, ch - 1) : std::max(iy - 1, 0);
        const uint8_t* near = in + (size_t)iy * stride;
        const uint8_t* far = in + (size_t)jy * stride;
        for (int x = 0; x < width; ++x) {
          int i = x >> 1, j = (x & 1) ? std::min(i + 1, cw - 1) : std::max(i - 1, 0);
          int a = 3 * near[i] + far[i], b = 3 * near[j] + far[j];
          out[(size_t)y * width + x] = (uint8_t)((3 * a + b + ((x & 1) ? 7 : 8)) >> 4);
        }
      }
    } else {  // anything else: replication
      for (int y = 0; y < height; ++y) {
        int iy = std::min(y * c.v / vmax, ch - 1);
        for (int x = 0; x < width; ++x) out[(size_t)y * width + x] = in[(size_t)iy * stride + std::min(x * c.h / hmax, cw - 1)];
      }
    }
  }

  int run(uint8_t** rgb, uint32_t* w, uint32_t* h) {
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return bad("not a JPEG file");
    const uint8_t* p = data + 2;
    const uint8_t* end = data + len;
    bool seen_scan = false, eoi = false;
    int rc;
    while (!eoi && p + 2 <= end) {
      if (p[0] != 0xFF) {  // tolerate garbage between segments
        ++p;
        continue;
      }
      int m = p[1];
      if (m == 0xFF) {
        ++p;
        continue;
      }
      p += 2;
      if (m == 0xD9) {
        eoi = true;
        break;
      }
      if (m == 0x01 || m == 0x00 || (m >= 0xD0 && m <= 0xD7)) continue;
      if (p + 2 > end) break;
      int n = ((p[0] << 8) | p[1]) - 2;
      if (n < 0 || p + 2 + n > end) return bad("truncated segment");
      const uint8_t* seg = p + 2;
      p += 2 + n;
      switch (m) {
        case 0xDB:
          if ((rc = parse_dqt(seg, n)) != RT_OK) return rc;
          break;
        case 0xC4:
          if ((rc = parse_dht(seg, n)) != RT_OK) return rc;
          break;
        case 0xC0:
        case 0xC1:
        case 0xC2:
          if ((rc = parse_sof(seg, n, m)) != RT_OK) return rc;
          break;
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          return bad("lossless, hierarchical and arithmetic-coded JPEG are not supported");
        case 0xDD:
          if (n < 2) return bad("short DRI");
          restart_interval = (seg[0] << 8) | seg[1];
          break;
        case 0xEE:
          if (n >= 12 && std::memcmp(seg, "Adobe", 5) == 0) adobe_transform = seg[11];
          break;
        case 0xDA:
          if ((rc = decode_scan(seg, n, p)) != RT_OK) return rc;
          seen_scan = true;
          break;
        default:
          break;  // APPn, COM, ...
      }
    }
    if (!have_frame || !seen_scan) return bad("no image data");
    for (int i = 0; i < ncomp; ++i)
      if (!qt_present[comp[i].tq]) return bad("missing quantisation table");
    reconstruct();
    uint8_t* out = (uint8_t*)std::malloc((size_t)width * height * 3);
    if (!out) return bad("out of memory");
    std::vector<uint8_t> pl[3];
    for (int i = 0; i < ncomp; ++i) upsample(comp[i], pl[i]);
    const size_t npix = (size_t)width * height;
    bool is_rgb = ncomp == 3 && (adobe_transform == 0 || (adobe_transform < 0 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B'));
    auto clamp8 = [](int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    if (ncomp == 1) {
      for (size_t i = 0; i < npix; ++i) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = pl[0][i];
    } else if (is_rgb) {
      for (size_t i = 0; i < npix; ++i) {
        out[3 * i] = pl[0][i];
        out[3 * i + 1] = pl[1][i];
        out[3 * i + 2] = pl[2][i];
      }
    } else {  // JFIF: 16.16 fixed point of 1.402, 0.344136, 0.714136, 1.772
      for (size_t i = 0; i < npix; ++i) {
        int y = pl[0][i], cb = pl[1][i] - 128, cr = pl[2][i] - 128;
        out[3 * i] = clamp8(y + ((91881 * cr + 32768) >> 16));
        out[3 * i + 1] = clamp8(y + ((-22554 * cb - 46802 * cr + 32768) >> 16));
        out[3 * i + 2] = clamp8(y + ((116130 * cb + 32768) >> 16));
      }
    }
    *rgb = out;
    *w = (uint32_t)width;
    *h = (uint32_t)height;
    return RT_OK;
  }
};

}  // namespace

int jpeg_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h, std::string& err) {
  Decoder d(bytes, len, err);
  return d.run(rgb, w, h);
}

}  // namespace rt
