// Image decoding for glTF textures: what `gltf::import_slice` hands to `rgba8_image`
// [ref crates/lib/src/loaders/gltf.rs:12-44,150-153].  The reference gets its pixels from the
// `image` crate behind `gltf` 1.4.1; nothing of that is available here, so this file carries
// a self-contained PNG reader (RFC 2083: all colour types, bit depths 1-16, Adam7) on top of
// its own inflate (RFC 1951), and a JPEG reader (ITU T.81 baseline / extended sequential
// Huffman, 8-bit, 1 or 3 components, any sampling factors, restart intervals; progressive,
// arithmetic, lossless and CMYK files are rejected).
//
// Channel expansion follows rgba8_image literally: a source with c < 4 channels is copied into
// the first c bytes of a zero-initialised RGBA texel (so grey images land in R only and
// RGB images get alpha 0) [ref gltf.rs:29-38].  16-bit sources keep their high byte (the
// reference says "16bits will break", gltf.rs:26).
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene.hpp"

namespace lp {
namespace {

// ------------------------------------------------------------------ inflate (RFC 1951)
struct BitReader {
  const uint8_t *p, *end;
  uint32_t buf = 0;
  int cnt = 0;
  uint32_t bits(int n) {
    while (cnt < n) {
      if (p >= end) throw std::runtime_error("inflate: out of input");
      buf |= (uint32_t)(*p++) << cnt;
      cnt += 8;
    }
    const uint32_t v = buf & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1u));
    buf >>= n;
    cnt -= n;
    return v;
  }
};

struct Huffman {
  uint16_t count[16] = {};
  std::vector<uint16_t> symbol;
  void build(const uint8_t *lengths, int n) {
    std::memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) count[lengths[i]]++;
    count[0] = 0;
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    symbol.assign(n, 0);
    for (int i = 0; i < n; ++i)
      if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
  }
  int decode(BitReader &br) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
      code |= (int)br.bits(1);
      const int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    throw std::runtime_error("inflate: bad Huffman code");
  }
};

void inflate_raw(const uint8_t *src, size_t n, std::vector<uint8_t> &out) {
  static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                     35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                   3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                     257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                     8193, 12289, 16385, 24577};
  static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                   7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  BitReader br{src, src + n};
  for (;;) {
    const uint32_t last = br.bits(1), type = br.bits(2);
    if (type == 0) {
      br.buf = 0;
      br.cnt = 0;
      if (br.end - br.p < 4) throw std::runtime_error("inflate: truncated stored block");
      const uint32_t len = br.p[0] | (br.p[1] << 8), nlen = br.p[2] | (br.p[3] << 8);
      br.p += 4;
      if ((len ^ 0xFFFFu) != nlen || (size_t)(br.end - br.p) < len)
        throw std::runtime_error("inflate: bad stored block");
      out.insert(out.end(), br.p, br.p + len);
      br.p += len;
    } else if (type == 1 || type == 2) {
      Huffman lit, dist;
      uint8_t lengths[320];
      if (type == 1) {
        int i = 0;
        for (; i < 144; ++i) lengths[i] = 8;
        for (; i < 256; ++i) lengths[i] = 9;
        for (; i < 280; ++i) lengths[i] = 7;
        for (; i < 288; ++i) lengths[i] = 8;
        lit.build(lengths, 288);
        for (i = 0; i < 30; ++i) lengths[i] = 5;
        dist.build(lengths, 30);
      } else {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5,
                                          11, 4, 12, 3, 13, 2, 14, 1, 15};
        const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1,
                  ncode = (int)br.bits(4) + 4;
        if (nlen > 286 || ndist > 30) throw std::runtime_error("inflate: bad counts");
        uint8_t cl[19] = {};
        for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
        Huffman clh;
        clh.build(cl, 19);
        int i = 0;
        while (i < nlen + ndist) {
          const int sym = clh.decode(br);
          if (sym < 16) {
            lengths[i++] = (uint8_t)sym;
          } else {
            int rep, val = 0;
            if (sym == 16) {
              if (i == 0) throw std::runtime_error("inflate: repeat without previous length");
              val = lengths[i - 1];
              rep = 3 + (int)br.bits(2);
            } else if (sym == 17) {
              rep = 3 + (int)br.bits(3);
            } else {
              rep = 11 + (int)br.bits(7);
            }
            if (i + rep > nlen + ndist) throw std::runtime_error("inflate: too many lengths");
            while (rep--) lengths[i++] = (uint8_t)val;
          }
        }
        lit.build(lengths, nlen);
        dist.build(lengths + nlen, ndist);
      }
      for (;;) {
        const int sym = lit.decode(br);
        if (sym < 256) {
          out.push_back((uint8_t)sym);
        } else if (sym == 256) {
          break;
        } else {
          const int ls = sym - 257;
          if (ls >= 29) throw std::runtime_error("inflate: bad length symbol");
          const size_t len = lbase[ls] + br.bits(lext[ls]);
          const int ds = dist.decode(br);
          if (ds >= 30) throw std::runtime_error("inflate: bad distance symbol");
          const size_t d = dbase[ds] + br.bits(dext[ds]);
          if (d > out.size()) throw std::runtime_error("inflate: distance too far back");
          const size_t start = out.size() - d;
          for (size_t k = 0; k < len; ++k) out.push_back(out[start + k]);
        }
      }
    } else {
      throw std::runtime_error("inflate: bad block type");
    }
    if (last) break;
  }
}

// ------------------------------------------------------------------ PNG
uint32_t be32(const uint8_t *p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undoes the per-scanline filters of one (sub-)image in place; returns bytes consumed
size_t png_unfilter(uint8_t *data, size_t avail, uint32_t w, uint32_t h, int bits_per_pixel,
                    std::vector<uint8_t> &rows) {
  const size_t stride = ((size_t)w * bits_per_pixel + 7) / 8;
  const int bpp = std::max(1, bits_per_pixel / 8);
  if (avail < (stride + 1) * h) throw std::runtime_error("png: not enough pixel data");
  rows.assign(stride * h, 0);
  std::vector<uint8_t> zero(stride, 0);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t *src = data + (stride + 1) * y;
    const int filter = src[0];
    ++src;
    uint8_t *cur = rows.data() + stride * y;
    const uint8_t *up = y ? cur - stride : zero.data();
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up[i],
                c = i >= (size_t)bpp ? up[i - bpp] : 0;
      int v = src[i];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: throw std::runtime_error("png: bad filter type");
      }
      cur[i] = (uint8_t)v;
    }
  }
  return (stride + 1) * h;
}

bool decode_png(const uint8_t *d, size_t n, Image &out) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (n < 8 || std::memcmp(d, sig, 8)) return false;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  size_t off = 8;
  bool have_ihdr = false;
  while (off + 12 <= n) {
    const uint32_t len = be32(d + off);
    const uint8_t *type = d + off + 4, *body = d + off + 8;
    if (off + 12 + (size_t)len > n) throw std::runtime_error("png: truncated chunk");
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len < 13) throw std::runtime_error("png: bad IHDR");
      w = be32(body);
      h = be32(body + 4);
      depth = body[8];
      ctype = body[9];
      interlace = body[12];
      if (body[10] || body[11]) throw std::runtime_error("png: unknown compression/filter");
      have_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!std::memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    off += 12 + (size_t)len;
  }
  // 16384 = the atlas layer limit (textures.cpp): a larger image could never be used, and the
  // header alone must not be able to ask for gigabytes
  if (!have_ihdr || !w || !h || w > 16384 || h > 16384) throw std::runtime_error("png: bad size");
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: throw std::runtime_error("png: bad colour type");
  }
  if (!(depth == 8 || depth == 16 || (channels == 1 && (depth == 1 || depth == 2 || depth == 4))) ||
      (ctype == 3 && depth == 16))
    throw std::runtime_error("png: bad bit depth");
  if (idat.size() < 2) throw std::runtime_error("png: no image data");
  std::vector<uint8_t> raw;
  raw.reserve(((size_t)w * channels * depth / 8 + 1) * h);
  inflate_raw(idat.data() + 2, idat.size() - 2, raw);  // zlib header skipped, Adler-32 ignored

  const int bits_pp = channels * depth;
  out.width = w;
  out.height = h;
  out.data.assign((size_t)w * h * 4, 0);
  // source channels after palette expansion (what the `image` crate reports to gltf)
  const int out_channels = ctype == 3 ? (trns.empty() ? 3 : 4) : channels;
  auto put = [&](const std::vector<uint8_t> &rows, size_t stride, uint32_t sx, uint32_t sy,
                 uint32_t dx, uint32_t dy) {
    uint8_t *dst = &out.data[((size_t)dy * w + dx) * 4];
    const uint8_t *row = rows.data() + stride * sy;
    uint32_t v[4] = {0, 0, 0, 0};
    if (depth < 8) {
      const uint32_t bit = sx * depth;
      const uint32_t s = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
      v[0] = ctype == 3 ? s : s * 255u / ((1u << depth) - 1u);
    } else {
      const int bytes = depth / 8;
      for (int c = 0; c < channels; ++c) v[c] = row[((size_t)sx * channels + c) * bytes];
    }
    if (ctype == 3) {
      const uint32_t i = v[0];
      if ((size_t)i * 3 + 3 > plte.size()) throw std::runtime_error("png: palette index out of range");
      dst[0] = plte[3 * i];
      dst[1] = plte[3 * i + 1];
      dst[2] = plte[3 * i + 2];
      if (out_channels == 4) dst[3] = i < trns.size() ? trns[i] : 255;
    } else {
      for (int c = 0; c < out_channels; ++c) dst[c] = (uint8_t)v[c];
    }
  };
  std::vector<uint8_t> rows;
  if (!interlace) {
    png_unfilter(raw.data(), raw.size(), w, h, bits_pp, rows);
    const size_t stride = ((size_t)w * bits_pp + 7) / 8;
    for (uint32_t y = 0; y < h; ++y)
      for (uint32_t x = 0; x < w; ++x) put(rows, stride, x, y, x, y);
  } else if (interlace == 1) {
    static const int x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t pos = 0;
    for (int p = 0; p < 7; ++p) {
      const uint32_t pw = (w + dx[p] - 1 - x0[p]) / dx[p], ph = (h + dy[p] - 1 - y0[p]) / dy[p];
      if (!pw || !ph || (uint32_t)x0[p] >= w || (uint32_t)y0[p] >= h) continue;
      pos += png_unfilter(raw.data() + pos, raw.size() - pos, pw, ph, bits_pp, rows);
      const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
      for (uint32_t y = 0; y < ph; ++y)
        for (uint32_t x = 0; x < pw; ++x) put(rows, stride, x, y, x0[p] + x * dx[p], y0[p] + y * dy[p]);
    }
  } else {
    throw std::runtime_error("png: bad interlace method");
  }
  return true;
}

// ------------------------------------------------------------------ JPEG (ITU T.81 sequential)
struct JpegHuff {
  bool present = false;
  uint8_t bits[17] = {};
  uint8_t vals[256] = {};
  int mincode[17], maxcode[18], valptr[17];
  void prepare() {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l];
      k += bits[l];
      maxcode[l] = bits[l] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7FFFFFFF;
  }
};

struct JpegComponent {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int dc_pred = 0;
  int blocks_w = 0, blocks_h = 0;  // allocated blocks (padded to whole MCUs)
  std::vector<uint8_t> plane;       // blocks_w*8 x blocks_h*8 samples
};

struct JpegBits {
  const uint8_t *p, *end;
  uint32_t buf = 0;
  int cnt = 0;
  int marker = 0;  // a marker met inside the entropy-coded segment
  void fill() {
    while (cnt <= 24) {
      int b = 0;
      if (!marker && p < end) {
        b = *p++;
        if (b == 0xFF) {
          int m = p < end ? *p : 0xD9;
          while (m == 0xFF && p + 1 < end) m = *++p;  // fill bytes
          if (m == 0) {
            ++p;
          } else {
            marker = m;
            ++p;
            b = 0;
          }
        }
      }
      buf |= (uint32_t)b << (24 - cnt);
      cnt += 8;
    }
  }
  int bit() {
    if (cnt < 1) fill();
    const int v = (int)(buf >> 31);
    buf <<= 1;
    --cnt;
    return v;
  }
  int receive(int n) {
    if (!n) return 0;
    if (cnt < n) fill();
    const int v = (int)(buf >> (32 - n));
    buf <<= n;
    cnt -= n;
    return v;
  }
  void reset() {
    buf = 0;
    cnt = 0;
    marker = 0;
  }
};

int jpeg_decode_symbol(JpegBits &br, const JpegHuff &h) {
  int code = 0;
  for (int l = 1; l <= 16; ++l) {
    code = (code << 1) | br.bit();
    if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l])
      return h.vals[h.valptr[l] + code - h.mincode[l]];
  }
  throw std::runtime_error("jpeg: bad Huffman code");
}

int jpeg_extend(int v, int t) { return (t && v < (1 << (t - 1))) ? v - (1 << t) + 1 : v; }

// separable 8x8 inverse DCT in double precision (T.81 A.3.3), level shift + clamp
void jpeg_idct(const int coef[64], const uint16_t q[64], uint8_t *dst, size_t stride) {
  static double c[8][8];
  static bool init = false;
  if (!init) {
    for (int x = 0; x < 8; ++x)
      for (int u = 0; u < 8; ++u)
        c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0);
    init = true;
  }
  double tmp[64], in[64];
  for (int i = 0; i < 64; ++i) in[i] = (double)coef[i] * q[i];
  for (int y = 0; y < 8; ++y)      // rows: v stays, u -> x
    for (int x = 0; x < 8; ++x) {
      double s = 0;
      for (int u = 0; u < 8; ++u) s += c[x][u] * in[y * 8 + u];
      tmp[y * 8 + x] = s;
    }
  for (int x = 0; x < 8; ++x)
    for (int y = 0; y < 8; ++y) {
      double s = 0;
      for (int v = 0; v < 8; ++v) s += c[y][v] * tmp[v * 8 + x];
      const long r = std::lround(s + 128.0);
      dst[y * stride + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

bool decode_jpeg(const uint8_t *d, size_t n, Image &out) {
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
  static const uint8_t zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  uint16_t qt[4][64] = {};
  JpegHuff hdc[4], hac[4];
  std::vector<JpegComponent> comps;
  uint32_t W = 0, H = 0;
  int restart_interval = 0, hmax = 1, vmax = 1;
  int adobe_transform = -1;
  bool decoded = false;
  size_t off = 2;
  while (off + 4 <= n && !decoded) {
    if (d[off] != 0xFF) {
      ++off;
      continue;
    }
    const int m = d[off + 1];
    if (m == 0xFF) {
      ++off;
      continue;
    }
    off += 2;
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) break;
    if (off + 2 > n) throw std::runtime_error("jpeg: truncated");
    const size_t len = ((size_t)d[off] << 8) | d[off + 1];
    if (len < 2 || off + len > n) throw std::runtime_error("jpeg: truncated segment");
    const uint8_t *s = d + off + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {  // DQT
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) throw std::runtime_error("jpeg: bad DQT");
        for (int k = 0; k < 64; ++k) {
          qt[tq][zigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
          i += pq ? 2 : 1;
        }
      }
    } else if (m == 0xC4) {  // DHT
      size_t i = 0;
      while (i + 17 <= sl) {
        const int tc = s[i] >> 4, th = s[i] & 15;
        if (tc > 1 || th > 3) throw std::runtime_error("jpeg: bad DHT");
        JpegHuff &h = tc ? hac[th] : hdc[th];
        int total = 0;
        for (int l = 1; l <= 16; ++l) total += (h.bits[l] = s[i + l]);
        i += 17;
        if (total > 256 || i + total > sl) throw std::runtime_error("jpeg: bad DHT");
        std::memcpy(h.vals, s + i, total);
        i += total;
        h.present = true;
        h.prepare();
      }
    } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1
      if (sl < 6 || s[0] != 8) throw std::runtime_error("jpeg: only 8-bit samples");
      H = (s[1] << 8) | s[2];
      W = (s[3] << 8) | s[4];
      const int nc = s[5];
      if ((nc != 1 && nc != 3) || sl < 6 + 3 * (size_t)nc)
        throw std::runtime_error("jpeg: unsupported component count");
      if (!W || !H || W > 16384 || H > 16384) throw std::runtime_error("jpeg: bad size");
      comps.resize(nc);
      for (int c = 0; c < nc; ++c) {
        comps[c].id = s[6 + 3 * c];
        comps[c].h = s[7 + 3 * c] >> 4;
        comps[c].v = s[7 + 3 * c] & 15;
        comps[c].tq = s[8 + 3 * c];
        if (comps[c].h < 1 || comps[c].h > 4 || comps[c].v < 1 || comps[c].v > 4 || comps[c].tq > 3)
          throw std::runtime_error("jpeg: bad component");
        hmax = std::max(hmax, comps[c].h);
        vmax = std::max(vmax, comps[c].v);
      }
    } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
      throw std::runtime_error("jpeg: progressive / lossless / arithmetic coding not supported");
    } else if (m == 0xDD) {
      if (sl >= 2) restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xEE) {
      if (sl >= 12 && !std::memcmp(s, "Adobe", 5)) adobe_transform = s[11];
    } else if (m == 0xDA) {  // SOS + entropy-coded data
      if (comps.empty()) throw std::runtime_error("jpeg: SOS before SOF");
      const int ns = s[0];
      if (ns != (int)comps.size() || sl < 1 + 2 * (size_t)ns + 3)
        throw std::runtime_error("jpeg: non-interleaved scans not supported");
      for (int k = 0; k < ns; ++k) {
        const int id = s[1 + 2 * k];
        JpegComponent *cp = nullptr;
        for (auto &c : comps)
          if (c.id == id) cp = &c;
        if (!cp) throw std::runtime_error("jpeg: unknown scan component");
        cp->td = s[2 + 2 * k] >> 4;
        cp->ta = s[2 + 2 * k] & 15;
        if (cp->td > 3 || cp->ta > 3 || !hdc[cp->td].present || !hac[cp->ta].present)
          throw std::runtime_error("jpeg: missing Huffman table");
      }
      if (comps.size() == 1) {  // a single-component scan is never interleaved (A.2.2)
        comps[0].h = comps[0].v = 1;
        hmax = vmax = 1;
      }
      const uint32_t mcu_w = 8 * hmax, mcu_h = 8 * vmax;
      const uint32_t mcus_x = (W + mcu_w - 1) / mcu_w, mcus_y = (H + mcu_h - 1) / mcu_h;
      for (auto &c : comps) {
        c.blocks_w = mcus_x * c.h;
        c.blocks_h = mcus_y * c.v;
        c.plane.assign((size_t)c.blocks_w * 8 * c.blocks_h * 8, 0);
        c.dc_pred = 0;
      }
      JpegBits br{d + off + len, d + n};
      int to_restart = restart_interval;
      for (uint32_t my = 0; my < mcus_y; ++my)
        for (uint32_t mx = 0; mx < mcus_x; ++mx) {
          if (restart_interval && to_restart == 0) {
            // byte-align, expect RSTn
            if (!br.marker) {
              br.cnt = 0;
              br.buf = 0;
              br.fill();
            }
            if (br.marker < 0xD0 || br.marker > 0xD7) throw std::runtime_error("jpeg: missing RST");
            br.reset();
            for (auto &c : comps) c.dc_pred = 0;
            to_restart = restart_interval;
          }
          for (auto &c : comps)
            for (int by = 0; by < c.v; ++by)
              for (int bx = 0; bx < c.h; ++bx) {
                int coef[64] = {};
                const int t = jpeg_decode_symbol(br, hdc[c.td]);
                if (t > 11) throw std::runtime_error("jpeg: bad DC size");
                c.dc_pred += jpeg_extend(br.receive(t), t);
                coef[0] = c.dc_pred;
                for (int k = 1; k < 64;) {
                  const int rs = jpeg_decode_symbol(br, hac[c.ta]);
                  const int r = rs >> 4, sz = rs & 15;
                  if (sz == 0) {
                    if (r == 15) {
                      k += 16;
                      continue;
                    }
                    break;
                  }
                  k += r;
                  if (k > 63) throw std::runtime_error("jpeg: AC index out of range");
                  coef[zigzag[k]] = jpeg_extend(br.receive(sz), sz);
                  ++k;
                }
                const size_t stride = (size_t)c.blocks_w * 8;
                uint8_t *dst = c.plane.data() + ((size_t)(my * c.v + by) * 8) * stride +
                               (size_t)(mx * c.h + bx) * 8;
                jpeg_idct(coef, qt[c.tq], dst, stride);
              }
          if (restart_interval) --to_restart;
        }
      decoded = true;
    }
    off += len;
  }
  if (!decoded) throw std::runtime_error("jpeg: no scan found");

  // ---- upsample to full resolution.  h2v1 / h2v2 use the triangle filter of libjpeg's
  // "fancy upsampling" (3/4 nearer + 1/4 farther, rounding as in jdsample.c semantics);
  // every other ratio replicates samples.
  out.width = W;
  out.height = H;
  out.data.assign((size_t)W * H * 4, 0);
  std::vector<std::vector<uint8_t>> full(comps.size());
  for (size_t ci = 0; ci < comps.size(); ++ci) {
    const JpegComponent &c = comps[ci];
    const size_t stride = (size_t)c.blocks_w * 8;
    const uint32_t cw = (W * c.h + hmax - 1) / hmax, ch = (H * c.v + vmax - 1) / vmax;
    std::vector<uint8_t> &f = full[ci];
    f.assign((size_t)W * H, 0);
    const int fx = hmax / c.h, fy = vmax / c.v;
    auto at = [&](long x, long y) -> int {
      x = x < 0 ? 0 : (x >= (long)cw ? (long)cw - 1 : x);
      y = y < 0 ? 0 : (y >= (long)ch ? (long)ch - 1 : y);
      return c.plane[(size_t)y * stride + x];
    };
    if (c.h == hmax && c.v == vmax) {
      for (uint32_t y = 0; y < H; ++y)
        std::memcpy(&f[(size_t)y * W], &c.plane[(size_t)y * stride], W);
    } else if (fx == 2 && fy == 1 && hmax % c.h == 0 && vmax % c.v == 0) {
      for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
          const long sx = x >> 1;
          const int near = at(sx, y);
          int v;
          if ((sx == 0 && !(x & 1)) || (sx == (long)cw - 1 && (x & 1)))
            v = near;
          else if (x & 1)
            v = (3 * near + at(sx + 1, y) + 2) >> 2;
          else
            v = (3 * near + at(sx - 1, y) + 1) >> 2;
          f[(size_t)y * W + x] = (uint8_t)v;
        }
    } else if (fx == 2 && fy == 2 && hmax % c.h == 0 && vmax % c.v == 0) {
      for (uint32_t y = 0; y < H; ++y) {
        const long sy = y >> 1, oy = (y & 1) ? sy + 1 : sy - 1;
        for (uint32_t x = 0; x < W; ++x) {
          const long sx = x >> 1;
          const int cur = 3 * at(sx, sy) + at(sx, oy);  // vertical pass, scaled by 4
          int v;
          if ((sx == 0 && !(x & 1)) || (sx == (long)cw - 1 && (x & 1))) {
            v = (cur * 4 + 8) >> 4;
          } else if (x & 1) {
            const int nxt = 3 * at(sx + 1, sy) + at(sx + 1, oy);
            v = (3 * cur + nxt + 7) >> 4;
          } else {
            const int prv = 3 * at(sx - 1, sy) + at(sx - 1, oy);
            v = (3 * cur + prv + 8) >> 4;
          }
          f[(size_t)y * W + x] = (uint8_t)v;
        }
      }
    } else {
      for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x)
          f[(size_t)y * W + x] = (uint8_t)at((long)((uint64_t)x * c.h / hmax), (long)((uint64_t)y * c.v / vmax));
    }
  }
  if (comps.size() == 1) {
    for (size_t i = 0; i < (size_t)W * H; ++i) out.data[4 * i] = full[0][i];  // R8 [ref gltf.rs:14]
  } else {
    const bool ycc = adobe_transform != 0;  // JFIF default; Adobe transform 0 = plain RGB
    for (size_t i = 0; i < (size_t)W * H; ++i) {
      if (!ycc) {
        out.data[4 * i] = full[0][i];
        out.data[4 * i + 1] = full[1][i];
        out.data[4 * i + 2] = full[2][i];
        continue;
      }
      const double Y = full[0][i], cb = full[1][i] - 128.0, cr = full[2][i] - 128.0;
      const long r = std::lround(Y + 1.402 * cr);
      const long g = std::lround(Y - 0.344136 * cb - 0.714136 * cr);
      const long b = std::lround(Y + 1.772 * cb);
      out.data[4 * i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
      out.data[4 * i + 1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
      out.data[4 * i + 2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
    }
  }
  return true;
}

}  // namespace

bool decode_image(const uint8_t *data, size_t size, Image &out, std::string &err) {
  try {
    if (decode_png(data, size, out)) return true;
    if (decode_jpeg(data, size, out)) return true;
    err = "unknown image format (PNG and JPEG are supported)";
  } catch (const std::exception &e) {
    err = e.what();
  }
  out = Image();
  return false;
}

}  // namespace lp
