// Texture atlas and probe sampling tables: the host half of SceneGPU::new_from_scene's
// "build atlas and copy to GPU" [ref crates/lib/src/scene.rs:172-184] and of ProbeGPU::new
// [ref scene.rs:71-121].  The reference delegates packing to albedo_backend::Atlas2D
// (un-vendored); the contract kept here is the one its call site shows: one reserved block
// per Scene image, in image order, addressed by the image index a Material stores.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <stdexcept>

#include "scene.hpp"

namespace lp {

// Shelf packing into square layers.  Images are placed tallest first (ties: wider first,
// then lower index) on left-to-right shelves; a layer that cannot take the next image is
// closed and a new one opened.  Layer edge = the smallest power of two that holds the largest
// image (at least 64).  Deterministic, so every GPU of a replicated scene builds the same
// atlas.
void build_atlas(const std::vector<Image> &images, uint32_t max_layer_size, Atlas &out) {
  out = Atlas();
  if (images.empty()) return;
  uint32_t max_dim = 1;
  for (const Image &im : images) {
    if (!im.width || !im.height || im.data.size() < (size_t)im.width * im.height * 4)
      throw std::invalid_argument("scene image with inconsistent size");
    max_dim = std::max(max_dim, std::max(im.width, im.height));
  }
  if (max_dim > max_layer_size) throw std::invalid_argument("scene image larger than the atlas limit");
  uint32_t size = 64;
  while (size < max_dim) size <<= 1;
  std::vector<uint32_t> order(images.size());
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    if (images[a].height != images[b].height) return images[a].height > images[b].height;
    return images[a].width > images[b].width;
  });
  out.size = size;
  out.blocks.assign(images.size(), AtlasBlock{0, 0, 0, 0, 0});
  uint32_t layer = 0, shelf_y = 0, shelf_h = 0, cursor_x = 0;
  for (uint32_t i : order) {
    const Image &im = images[i];
    if (cursor_x + im.width > size) {  // next shelf
      shelf_y += shelf_h;
      shelf_h = 0;
      cursor_x = 0;
    }
    if (shelf_y + im.height > size) {  // next layer
      ++layer;
      shelf_y = shelf_h = cursor_x = 0;
    }
    out.blocks[i] = AtlasBlock{cursor_x, shelf_y, im.width, im.height, layer};
    cursor_x += im.width;
    shelf_h = std::max(shelf_h, im.height);
  }
  out.layers = layer + 1;
  out.texels.assign((size_t)out.layers * size * size * 4, 0);
  out.gpu_blocks.resize(images.size() * 4);
  for (size_t i = 0; i < images.size(); ++i) {
    const AtlasBlock &b = out.blocks[i];
    const Image &im = images[i];
    for (uint32_t y = 0; y < b.h; ++y)
      std::memcpy(&out.texels[(((size_t)b.layer * size + b.y + y) * size + b.x) * 4],
                  &im.data[(size_t)y * im.width * 4], (size_t)im.width * 4);
    out.gpu_blocks[4 * i + 0] = b.x | (b.y << 16);
    out.gpu_blocks[4 * i + 1] = b.w | (b.h << 16);
    out.gpu_blocks[4 * i + 2] = b.layer;
    out.gpu_blocks[4 * i + 3] = 0;
  }
}

// f(x, y) = luminance(texel) * sin(pi (y + 0.5) / h), accumulated in double;
// pmf = f / sum f; cdf_row[y] = sum_{y' <= y} rowsum / total; cdf_col[y][x] =
// sum_{x' <= x} f / rowsum(y) (all 1 for an empty row); the last entry of every CDF is
// exactly 1.  A black probe gets the uniform-over-solid-angle distribution (f = sin theta).
void build_probe_tables(const uint8_t *rgbe8, uint32_t w, uint32_t h, ProbeTables &out) {
  const size_t n = (size_t)w * h;
  std::vector<double> f(n);
  std::vector<double> rowsum(h, 0.0);
  double total = 0.0;
  for (int pass = 0; pass < 2 && !(total > 0.0); ++pass) {
    total = 0.0;
    for (uint32_t y = 0; y < h; ++y) {
      const double st = std::sin(M_PI * ((double)y + 0.5) / (double)h);
      double rs = 0.0;
      for (uint32_t x = 0; x < w; ++x) {
        const uint8_t *p = rgbe8 + 4 * ((size_t)y * w + x);
        double lum = 1.0;
        if (pass == 0) {
          if (p[3] == 0) {
            lum = 0.0;
          } else {
            const float s = std::ldexp(1.0f, (int)p[3] - (128 + 8));
            const float r = (float)p[0] * s, g = (float)p[1] * s, b = (float)p[2] * s;
            lum = 0.2126 * (double)r + 0.7152 * (double)g + 0.0722 * (double)b;
          }
        }
        const double v = lum * st;
        f[(size_t)y * w + x] = v;
        rs += v;
      }
      rowsum[y] = rs;
      total += rs;
    }
  }
  out.pmf.resize(n);
  out.cdf_row.resize(h);
  out.cdf_col.resize(n);
  double acc_rows = 0.0;
  for (uint32_t y = 0; y < h; ++y) {
    acc_rows += rowsum[y];
    out.cdf_row[y] = y + 1 == h ? 1.0f : (float)(acc_rows / total);
    double acc = 0.0;
    for (uint32_t x = 0; x < w; ++x) {
      const size_t i = (size_t)y * w + x;
      acc += f[i];
      out.pmf[i] = (float)(f[i] / total);
      out.cdf_col[i] = (x + 1 == w || !(rowsum[y] > 0.0)) ? 1.0f : (float)(acc / rowsum[y]);
    }
  }
}

}  // namespace lp
