// Scene ingest: loaders::load_gltf and loaders::load_binary_from_path
// [ref crates/lib/src/loaders/gltf.rs:46-156, loaders/binary.rs:6-70].
// The reference delegates parsing to the `gltf` 1.4.1 crate; this is a self-contained
// GLB / .gltf(JSON with embedded base64 buffers) reader covering what load_gltf consumes:
// POSITION / NORMAL / TEXCOORD_0 / indices accessors, pbrMetallicRoughness factors,
// node-local transforms (matrix or TRS), baseColorTexture / metallicRoughnessTexture, and the
// images (bufferView or data-URI PNG / JPEG, decoded by image_decode.cpp).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>

#include "scene.hpp"

namespace lp {
namespace {

// ------------------------------------------------------------------ minimal JSON DOM
struct JValue {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JValue> arr;
  std::vector<std::pair<std::string, JValue>> obj;

  const JValue *get(const char *key) const {
    if (kind != Obj) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  double number(const char *key, double dflt) const {
    const JValue *v = get(key);
    return (v && v->kind == Num) ? v->num : dflt;
  }
  // integers arrive as JSON numbers: anything that is not a finite value of at most 2^53 in
  // magnitude (NaN, 1e308, ...) cannot be converted without undefined behaviour and is
  // mapped to -1, which every caller rejects as an index, an offset or a count
  long integer(const char *key, long dflt) const {
    const double v = number(key, (double)dflt);
    if (!(v >= -9007199254740992.0 && v <= 9007199254740992.0)) return -1;
    return (long)v;
  }
  std::string string(const char *key, const char *dflt = "") const {
    const JValue *v = get(key);
    return (v && v->kind == Str) ? v->str : std::string(dflt);
  }
  size_t size() const { return kind == Arr ? arr.size() : 0; }
};

struct JParser {
  const char *p, *end;
  int depth = 0;
  [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("json: ") + what); }
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  JValue parse() {
    ws();
    JValue v = value();
    ws();
    return v;
  }
  JValue value() {
    if (++depth > 256) fail("nesting too deep");
    ws();
    if (p >= end) fail("unexpected end");
    JValue v;
    const char c = *p;
    if (c == '{') {
      v.kind = JValue::Obj;
      ++p;
      ws();
      if (p < end && *p == '}') {
        ++p;
      } else {
        for (;;) {
          ws();
          if (p >= end || *p != '"') fail("expected key");
          std::string key = str();
          ws();
          if (p >= end || *p != ':') fail("expected ':'");
          ++p;
          v.obj.emplace_back(std::move(key), value());
          ws();
          if (p < end && *p == ',') {
            ++p;
            continue;
          }
          if (p < end && *p == '}') {
            ++p;
            break;
          }
          fail("expected ',' or '}'");
        }
      }
    } else if (c == '[') {
      v.kind = JValue::Arr;
      ++p;
      ws();
      if (p < end && *p == ']') {
        ++p;
      } else {
        for (;;) {
          v.arr.push_back(value());
          ws();
          if (p < end && *p == ',') {
            ++p;
            continue;
          }
          if (p < end && *p == ']') {
            ++p;
            break;
          }
          fail("expected ',' or ']'");
        }
      }
    } else if (c == '"') {
      v.kind = JValue::Str;
      v.str = str();
    } else if (c == 't' && end - p >= 4 && !std::strncmp(p, "true", 4)) {
      v.kind = JValue::Bool;
      v.b = true;
      p += 4;
    } else if (c == 'f' && end - p >= 5 && !std::strncmp(p, "false", 5)) {
      v.kind = JValue::Bool;
      p += 5;
    } else if (c == 'n' && end - p >= 4 && !std::strncmp(p, "null", 4)) {
      p += 4;
    } else {
      v.kind = JValue::Num;
      std::string tmp;
      while (p < end && (std::strchr("+-0123456789.eE", *p) != nullptr)) tmp.push_back(*p++);
      if (tmp.empty()) fail("unexpected character");
      char *e = nullptr;
      v.num = std::strtod(tmp.c_str(), &e);
      if (!e || *e) fail("bad number");
    }
    --depth;
    return v;
  }
  std::string str() {
    std::string s;
    ++p;  // opening quote
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) fail("bad escape");
        switch (*p) {
          case 'n': s.push_back('\n'); break;
          case 't': s.push_back('\t'); break;
          case 'r': s.push_back('\r'); break;
          case 'b': s.push_back('\b'); break;
          case 'f': s.push_back('\f'); break;
          case 'u': {
            if (end - p < 5) fail("bad \\u escape");
            unsigned cp = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
            p += 4;
            if (cp < 0x80) {
              s.push_back((char)cp);
            } else if (cp < 0x800) {
              s.push_back((char)(0xC0 | (cp >> 6)));
              s.push_back((char)(0x80 | (cp & 0x3F)));
            } else {
              s.push_back((char)(0xE0 | (cp >> 12)));
              s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
              s.push_back((char)(0x80 | (cp & 0x3F)));
            }
            break;
          }
          default: s.push_back(*p);
        }
        ++p;
      } else {
        s.push_back(*p++);
      }
    }
    if (p >= end) fail("unterminated string");
    ++p;
    return s;
  }
};

std::vector<uint8_t> base64_decode(const char *s, size_t n) {
  std::vector<uint8_t> out;
  out.reserve(n * 3 / 4);
  uint32_t acc = 0;
  int bits = 0;
  for (size_t i = 0; i < n; ++i) {
    const char c = s[i];
    int v;
    if (c >= 'A' && c <= 'Z') v = c - 'A';
    else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
    else if (c >= '0' && c <= '9') v = c - '0' + 52;
    else if (c == '+' || c == '-') v = 62;
    else if (c == '/' || c == '_') v = 63;
    else continue;
    acc = (acc << 6) | (uint32_t)v;
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out.push_back((uint8_t)((acc >> bits) & 0xFF));
    }
  }
  return out;
}

// ------------------------------------------------------------------ accessors
struct Doc {
  JValue root;
  std::vector<std::vector<uint8_t>> buffers;
};

int type_components(const std::string &t) {
  if (t == "SCALAR") return 1;
  if (t == "VEC2") return 2;
  if (t == "VEC3") return 3;
  if (t == "VEC4") return 4;
  if (t == "MAT4") return 16;
  return 0;
}
int component_size(long ct) {
  switch (ct) {
    case 5120: case 5121: return 1;
    case 5122: case 5123: return 2;
    case 5125: case 5126: return 4;
  }
  return 0;
}

// [offset, offset + length) lies inside a buffer of `size` bytes (no overflow, no negatives)
bool span_ok(long offset, long length, size_t size) {
  return offset >= 0 && length >= 0 && (size_t)offset <= size &&
         (size_t)length <= size - (size_t)offset;
}

// Where accessor `a` reads: validates bufferView / buffer indices and that `count` elements
// of `elem` bytes, `stride` apart, lie inside the buffer -- with every quantity taken from
// the file checked for sign and overflow BEFORE anything is allocated or read.  Returns false
// for an invalid accessor; *bytes == nullptr (with true) for an accessor without a bufferView
// (all zeros; sparse accessors are not expanded).
constexpr size_t kMaxAccessorCount = (size_t)1 << 30;
bool accessor_view(const Doc &d, const JValue &a, size_t elem, size_t &count,
                   const uint8_t *&bytes, size_t &stride) {
  bytes = nullptr;
  stride = elem;
  const long n = a.integer("count", 0);
  if (n < 0 || (size_t)n > kMaxAccessorCount) return false;
  count = (size_t)n;
  const JValue *bvp = a.get("bufferView");
  if (!bvp) return true;
  const long bv_index = a.integer("bufferView", -1);
  const JValue *bvs = d.root.get("bufferViews");
  if (!bvs || bv_index < 0 || (size_t)bv_index >= bvs->size()) return false;
  const JValue &bv = bvs->arr[bv_index];
  const long buf = bv.integer("buffer", 0);
  if (buf < 0 || (size_t)buf >= d.buffers.size()) return false;
  const std::vector<uint8_t> &buffer = d.buffers[buf];
  const long bv_off = bv.integer("byteOffset", 0), acc_off = a.integer("byteOffset", 0);
  const long bv_stride = bv.integer("byteStride", 0);
  if (bv_off < 0 || acc_off < 0 || bv_stride < 0) return false;
  if (bv_stride) stride = (size_t)bv_stride;
  if (!span_ok(bv_off, acc_off, buffer.size())) return false;  // base = bv_off + acc_off
  const size_t base = (size_t)bv_off + (size_t)acc_off;
  if (count) {
    if (elem > buffer.size() - base) return false;
    const size_t room = buffer.size() - base - elem;  // bytes left for count - 1 strides
    if (stride == 0 || (count - 1) > room / stride) return false;
  }
  bytes = buffer.data() + base;
  return true;
}

// Reads accessor `index` as floats (out_comp components per element), following the
// `gltf` crate's `into_f32` normalisation rules for integer texcoords.
bool read_accessor_f32(const Doc &d, long index, int want_comp, std::vector<float> &out,
                       size_t &count) {
  const JValue *accs = d.root.get("accessors");
  if (!accs || index < 0 || (size_t)index >= accs->size()) return false;
  const JValue &a = accs->arr[index];
  const int comps = type_components(a.string("type"));
  const long ct = a.integer("componentType", 0);
  const int csz = component_size(ct);
  count = 0;
  if (comps != want_comp || !csz) return false;
  const uint8_t *bytes = nullptr;
  size_t stride = 0;
  if (!accessor_view(d, a, (size_t)comps * csz, count, bytes, stride)) return false;
  out.assign(count * comps, 0.f);
  if (!bytes) return true;  // all zeros (sparse accessors are not expanded)
  const bool normalized = a.get("normalized") && a.get("normalized")->b;
  for (size_t i = 0; i < count; ++i) {
    const uint8_t *src = bytes + i * stride;
    for (int c = 0; c < comps; ++c) {
      float v = 0.f;
      switch (ct) {
        case 5126: std::memcpy(&v, src + 4 * c, 4); break;
        case 5121: v = src[c] / (normalized || want_comp == 2 ? 255.f : 1.f); break;
        case 5120: v = std::max((int8_t)src[c] / 127.f, -1.f); break;
        case 5123: { uint16_t u; std::memcpy(&u, src + 2 * c, 2); v = u / (normalized || want_comp == 2 ? 65535.f : 1.f); break; }
        case 5122: { int16_t s; std::memcpy(&s, src + 2 * c, 2); v = std::max(s / 32767.f, -1.f); break; }
        case 5125: { uint32_t u; std::memcpy(&u, src + 4 * c, 4); v = (float)u; break; }
      }
      out[i * comps + c] = v;
    }
  }
  return true;
}

bool read_accessor_u32(const Doc &d, long index, std::vector<uint32_t> &out) {
  const JValue *accs = d.root.get("accessors");
  if (!accs || index < 0 || (size_t)index >= accs->size()) return false;
  const JValue &a = accs->arr[index];
  const long ct = a.integer("componentType", 0);
  const int csz = component_size(ct);
  if (type_components(a.string("type")) != 1 || !csz || ct == 5126) return false;
  size_t count = 0, stride = 0;
  const uint8_t *bytes = nullptr;
  if (!accessor_view(d, a, (size_t)csz, count, bytes, stride)) return false;
  out.assign(count, 0u);
  if (!bytes) return true;
  for (size_t i = 0; i < count; ++i) {
    const uint8_t *src = bytes + i * stride;
    switch (csz) {
      case 1: out[i] = src[0]; break;
      case 2: { uint16_t u; std::memcpy(&u, src, 2); out[i] = u; break; }
      default: std::memcpy(&out[i], src, 4);
    }
  }
  return true;
}

// gltf::scene::Transform::matrix(): column-major local matrix from `matrix` or T*R*S.
void node_matrix(const JValue &node, float m[16]) {
  const JValue *mat = node.get("matrix");
  if (mat && mat->size() == 16) {
    for (int i = 0; i < 16; ++i) m[i] = (float)mat->arr[i].num;
    return;
  }
  float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
  if (const JValue *v = node.get("translation"))
    for (size_t i = 0; i < 3 && i < v->size(); ++i) t[i] = (float)v->arr[i].num;
  if (const JValue *v = node.get("rotation"))
    for (size_t i = 0; i < 4 && i < v->size(); ++i) r[i] = (float)v->arr[i].num;
  if (const JValue *v = node.get("scale"))
    for (size_t i = 0; i < 3 && i < v->size(); ++i) s[i] = (float)v->arr[i].num;
  const float x = r[0], y = r[1], z = r[2], w = r[3];
  const float x2 = x + x, y2 = y + y, z2 = z + z;
  const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
  const float wx = w * x2, wy = w * y2, wz = w * z2;
  m[0] = (1 - (yy + zz)) * s[0]; m[1] = (xy + wz) * s[0]; m[2] = (xz - wy) * s[0]; m[3] = 0;
  m[4] = (xy - wz) * s[1]; m[5] = (1 - (xx + zz)) * s[1]; m[6] = (yz + wx) * s[1]; m[7] = 0;
  m[8] = (xz + wy) * s[2]; m[9] = (yz - wx) * s[2]; m[10] = (1 - (xx + yy)) * s[2]; m[11] = 0;
  m[12] = t[0]; m[13] = t[1]; m[14] = t[2]; m[15] = 1;
}

}  // namespace

lp_status load_gltf(const uint8_t *data, size_t size, Scene &scene, std::string &err) {
  Doc doc;
  DeferredBuildScope build_scope(scene);  // the meshes' trees are built together at the end
  try {
    if (!data || size < 4) throw std::runtime_error("empty input");
    const uint8_t *json_ptr = data;
    size_t json_len = size;
    std::vector<uint8_t> bin_chunk;
    bool has_bin = false;
    if (size >= 12 && !std::memcmp(data, "glTF", 4)) {
      uint32_t version, total;
      std::memcpy(&version, data + 4, 4);
      std::memcpy(&total, data + 8, 4);
      if (version != 2) throw std::runtime_error("unsupported GLB version");
      if (total > size) throw std::runtime_error("truncated GLB");
      size_t off = 12;
      json_ptr = nullptr;
      while (off + 8 <= total) {
        uint32_t clen, ctype;
        std::memcpy(&clen, data + off, 4);
        std::memcpy(&ctype, data + off + 4, 4);
        off += 8;
        if (off + clen > total) throw std::runtime_error("truncated GLB chunk");
        if (ctype == 0x4E4F534Au && !json_ptr) {  // "JSON"
          json_ptr = data + off;
          json_len = clen;
        } else if (ctype == 0x004E4942u && !has_bin) {  // "BIN\0"
          bin_chunk.assign(data + off, data + off + clen);
          has_bin = true;
        }
        off += (clen + 3u) & ~3u;
      }
      if (!json_ptr) throw std::runtime_error("GLB without JSON chunk");
    }
    JParser parser{(const char *)json_ptr, (const char *)json_ptr + json_len};
    doc.root = parser.parse();
    if (doc.root.kind != JValue::Obj) throw std::runtime_error("glTF root is not an object");

    if (const JValue *bufs = doc.root.get("buffers")) {
      for (size_t i = 0; i < bufs->size(); ++i) {
        const JValue &b = bufs->arr[i];
        const std::string uri = b.string("uri");
        if (uri.empty()) {
          if (i == 0 && has_bin) doc.buffers.push_back(bin_chunk);
          else throw std::runtime_error("buffer without uri and no BIN chunk");
        } else if (uri.rfind("data:", 0) == 0) {
          const size_t comma = uri.find(',');
          if (comma == std::string::npos) throw std::runtime_error("bad data uri");
          doc.buffers.push_back(base64_decode(uri.c_str() + comma + 1, uri.size() - comma - 1));
        } else {
          // gltf::import_slice cannot resolve external files either (gltf::Error::Io)
          err = "failed to load gltf";
          return LP_ERR_FILE_NOT_FOUND;
        }
      }
    }
  } catch (const std::exception &e) {
    err = e.what();
    return LP_ERR_FILE_NOT_FOUND;  // every import error maps to FileNotFound [ref gltf.rs:49-55]
  }

  // everything below only appends to the scene: a failure half way rolls it all back
  const Scene::Mark mark = scene.mark();
  try {
    // ---- meshes: one BLAS entry per primitive that has POSITION and a triangle mode
    const uint32_t bvh_offset = (uint32_t)scene.entries.size();
    (void)bvh_offset;
    std::vector<std::vector<long>> prim_entry;  // [mesh][primitive] -> entry index or -1
    const JValue *meshes = doc.root.get("meshes");
    for (size_t mi = 0; meshes && mi < meshes->size(); ++mi) {
      prim_entry.emplace_back();
      const JValue *prims = meshes->arr[mi].get("primitives");
      for (size_t pi = 0; prims && pi < prims->size(); ++pi) {
        const JValue &prim = prims->arr[pi];
        prim_entry.back().push_back(-1);
        const JValue *attrs = prim.get("attributes");
        if (!attrs || !attrs->get("POSITION")) continue;  // [ref gltf.rs:64-66]
        const long mode = prim.integer("mode", 4);
        if (mode != 4 && mode != 5 && mode != 6) continue;  // [ref gltf.rs:68-73]
        std::vector<float> pos, nrm, uv;
        size_t vcount = 0, ncount = 0, tcount = 0;
        if (!read_accessor_f32(doc, attrs->integer("POSITION", -1), 3, pos, vcount)) continue;
        const bool has_n = attrs->get("NORMAL") &&
                           read_accessor_f32(doc, attrs->integer("NORMAL", -1), 3, nrm, ncount) &&
                           ncount == vcount;
        const bool has_t =
            attrs->get("TEXCOORD_0") &&
            read_accessor_f32(doc, attrs->integer("TEXCOORD_0", -1), 2, uv, tcount) &&
            tcount == vcount;
        std::vector<uint32_t> idx;
        const bool indexed = prim.get("indices") && read_accessor_u32(doc, prim.integer("indices", -1), idx);
        if (!indexed) {
          idx.resize(vcount);
          for (size_t i = 0; i < vcount; ++i) idx[i] = (uint32_t)i;
        }
        // expand strips / fans into a triangle list (the reference forwards the raw
        // index list for those modes; documented deviation in DESIGN.md)
        std::vector<uint32_t> tris;
        if (mode == 4) {
          tris.swap(idx);
          tris.resize(tris.size() / 3 * 3);
        } else if (mode == 5) {
          for (size_t i = 2; i < idx.size(); ++i) {
            if (i & 1) { tris.push_back(idx[i - 1]); tris.push_back(idx[i - 2]); tris.push_back(idx[i]); }
            else { tris.push_back(idx[i - 2]); tris.push_back(idx[i - 1]); tris.push_back(idx[i]); }
          }
        } else {
          for (size_t i = 2; i < idx.size(); ++i) {
            tris.push_back(idx[0]); tris.push_back(idx[i - 1]); tris.push_back(idx[i]);
          }
        }
        const uint32_t entry =
            scene.add_bvh(pos.data(), 12, has_n ? nrm.data() : nullptr, 12,
                          has_t ? uv.data() : nullptr, 8, vcount, tris.data(), tris.size());
        prim_entry.back().back() = (long)entry;
      }
    }

    // ---- images [ref gltf.rs:150-153]: every glTF image becomes one Scene image, in order,
    // so image i of this file is scene image texture_offset + i.  An image that cannot be
    // decoded (external uri, progressive JPEG, ...) keeps its slot as a 1x1 white texel so
    // the indices of the others stay valid (gltf::import_slice would fail the whole load).
    const uint32_t texture_offset = (uint32_t)scene.images.size();
    const JValue *imgs = doc.root.get("images");
    for (size_t i = 0; imgs && i < imgs->size(); ++i) {
      const JValue &im = imgs->arr[i];
      Image decoded;
      std::string ierr;
      bool ok = false;
      const long bv_index = im.integer("bufferView", -1);
      const std::string uri = im.string("uri");
      if (bv_index >= 0) {
        const JValue *bvs = doc.root.get("bufferViews");
        if (bvs && (size_t)bv_index < bvs->size()) {
          const JValue &bv = bvs->arr[bv_index];
          const long buf = bv.integer("buffer", 0);
          const long boff = bv.integer("byteOffset", 0), blen = bv.integer("byteLength", 0);
          if (buf >= 0 && (size_t)buf < doc.buffers.size() &&
              span_ok(boff, blen, doc.buffers[buf].size()))
            ok = decode_image(doc.buffers[buf].data() + boff, (size_t)blen, decoded, ierr);
        }
      } else if (uri.rfind("data:", 0) == 0) {
        const size_t comma = uri.find(',');
        if (comma != std::string::npos) {
          const std::vector<uint8_t> bytes =
              base64_decode(uri.c_str() + comma + 1, uri.size() - comma - 1);
          ok = decode_image(bytes.data(), bytes.size(), decoded, ierr);
        }
      }
      if (!ok) {
        decoded.width = decoded.height = 1;
        decoded.data.assign(4, 255);
      }
      scene.images.push_back(std::move(decoded));
    }
    // texture -> image: the reference stores `texture_offset + texture().index()` and then
    // pushes one Scene image per glTF IMAGE [ref gltf.rs:117-124,150-153], which only
    // addresses the right pixels when texture i uses image i; this loader resolves
    // textures[i].source (documented deviation).
    const JValue *texs = doc.root.get("textures");
    auto texture_image = [&](const JValue *info) -> uint32_t {
      if (!info) return LP_INVALID_INDEX;
      const long ti = info->integer("index", -1);
      if (!texs || ti < 0 || (size_t)ti >= texs->size()) return LP_INVALID_INDEX;
      const long src = texs->arr[ti].integer("source", -1);
      if (!imgs || src < 0 || (size_t)src >= imgs->size()) return LP_INVALID_INDEX;
      return texture_offset + (uint32_t)src;
    };

    // ---- materials [ref gltf.rs:109-127]
    const uint32_t mat_offset = (uint32_t)scene.materials.size();
    const JValue *mats = doc.root.get("materials");
    for (size_t i = 0; mats && i < mats->size(); ++i) {
      lp_material m{};
      m.color[0] = m.color[1] = m.color[2] = m.color[3] = 1.f;
      m.roughness = 1.f;
      m.reflectivity = 1.f;  // glTF default metallicFactor = 1
      m.albedo_texture = LP_INVALID_INDEX;
      m.mra_texture = LP_INVALID_INDEX;
      if (const JValue *pbr = mats->arr[i].get("pbrMetallicRoughness")) {
        if (const JValue *c = pbr->get("baseColorFactor"))
          for (size_t k = 0; k < 4 && k < c->size(); ++k) m.color[k] = (float)c->arr[k].num;
        m.roughness = (float)pbr->number("roughnessFactor", 1.0);
        m.reflectivity = (float)pbr->number("metallicFactor", 1.0);
        m.albedo_texture = texture_image(pbr->get("baseColorTexture"));
        m.mra_texture = texture_image(pbr->get("metallicRoughnessTexture"));
      }
      scene.materials.push_back(m);
      scene.emission.push_back({0.f, 0.f, 0.f, 0.f});
    }

    // ---- nodes: LOCAL transform only, no hierarchy walk [ref gltf.rs:129-148]
    const JValue *nodes = doc.root.get("nodes");
    for (size_t ni = 0; nodes && ni < nodes->size(); ++ni) {
      const JValue &node = nodes->arr[ni];
      const long mesh_index = node.integer("mesh", -1);
      if (mesh_index < 0 || (size_t)mesh_index >= prim_entry.size()) continue;
      const size_t mi = (size_t)mesh_index;
      float m[16];
      node_matrix(node, m);
      const JValue *prims = meshes->arr[mi].get("primitives");
      for (size_t pi = 0; pi < prim_entry[mi].size(); ++pi) {
        if (prim_entry[mi][pi] < 0) continue;
        // missing (or unusable) material -> mat_offset + u32::MAX, wrapping [ref gltf.rs:137-144]
        const long mat_i = prims->arr[pi].integer("material", -1);
        const uint32_t material_index =
            (mat_i >= 0 && mat_i < 0xFFFFFFFFL) ? (uint32_t)mat_i : 0xFFFFFFFFu;
        uint32_t material = mat_offset + material_index;
        if (material >= scene.materials.size()) material = 0;
        scene.add_instance((uint32_t)prim_entry[mi][pi], m, material);
      }
    }
    build_scope.finish();
    scene.derived_dirty = true;  // images / materials were pushed straight onto the arrays
  } catch (const std::exception &e) {
    scene.rollback(mark);
    err = e.what();
    return LP_ERR_ACCEL_BUILD;
  }
  return LP_OK;
}

lp_status load_binary(const char *path, Scene &scene, std::string &err) {
  // u32 triangle count, then 3*count little-endian vec4 positions [ref binary.rs:6-31].
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    err = path ? path : "(null)";
    return LP_ERR_FILE_NOT_FOUND;
  }
  uint32_t tri_count = 0;
  f.read((char *)&tri_count, 4);
  if (!f) {
    err = std::string(path) + ": truncated";
    return LP_ERR_FILE_NOT_FOUND;
  }
  const size_t vcount = (size_t)tri_count * 3;
  // the count comes from the file: compare it with what the file holds BEFORE allocating
  const std::streampos here = f.tellg();
  f.seekg(0, std::ios::end);
  const std::streampos end = f.tellg();
  f.seekg(here);
  if (!f || end < here || (uint64_t)(end - here) < (uint64_t)vcount * 16u) {
    err = std::string(path) + ": truncated";
    return LP_ERR_FILE_NOT_FOUND;
  }
  std::vector<float> raw, nrm;
  try {
    raw.resize(vcount * 4);
    nrm.resize(vcount * 3);
  } catch (const std::exception &e) {
    err = e.what();
    return LP_ERR_ACCEL_BUILD;
  }
  f.read((char *)raw.data(), (std::streamsize)(raw.size() * 4));
  if (!f) {
    err = std::string(path) + ": truncated";
    return LP_ERR_FILE_NOT_FOUND;
  }
  // flat normals: cross(normalize(v0-v1), normalize(v0-v2)) [ref binary.rs:33-47]
  for (size_t i = 0; i < vcount; i += 3) {
    const float *a = &raw[4 * i], *b = &raw[4 * (i + 1)], *c = &raw[4 * (i + 2)];
    float e0[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    float e1[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]};
    const float l0 = std::sqrt(e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2]);
    const float l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    for (int k = 0; k < 3; ++k) {
      e0[k] /= l0;
      e1[k] /= l1;
    }
    const float n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2],
                        e0[0] * e1[1] - e0[1] * e1[0]};
    for (int v = 0; v < 3; ++v)
      for (int k = 0; k < 3; ++k) nrm[3 * (i + v) + k] = n[k];
  }
  const Scene::Mark mark = scene.mark();
  try {
    const uint32_t blas = scene.add_bvh(raw.data(), 16, nrm.data(), 12, nullptr, 0, vcount, nullptr, 0);
    const uint32_t material_index = (uint32_t)scene.materials.size();
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    scene.add_instance(blas, identity, material_index);  // [ref binary.rs:55-61]
    lp_material m{};  // white, fully rough dielectric [ref binary.rs:63-69]
    m.color[0] = m.color[1] = m.color[2] = m.color[3] = 1.f;
    m.roughness = 1.f;
    m.reflectivity = 0.f;
    m.albedo_texture = LP_INVALID_INDEX;
    m.mra_texture = LP_INVALID_INDEX;
    scene.materials.push_back(m);
    scene.emission.push_back({0.f, 0.f, 0.f, 0.f});
    scene.derived_dirty = true;
  } catch (const std::exception &e) {
    scene.rollback(mark);
    err = e.what();
    return LP_ERR_ACCEL_BUILD;
  }
  return LP_OK;
}

}  // namespace lp
