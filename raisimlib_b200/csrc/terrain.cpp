// Height-map front end: a fractal terrain generator in the spirit of raisim::TerrainProperties /
// World::addHeightMap(centerX, centerY, terrainProperties) ([RECALL]; upstream's noise source is not public, so the
// heights cannot match upstream's -- the parameters and their meaning do).  Host code, runs once at set-up.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/rsb.h"

namespace {
inline uint32_t hash2(int x, int y, uint32_t seed) {
  uint32_t h = seed ^ (uint32_t(x) * 0x9E3779B1u) ^ (uint32_t(y) * 0x85EBCA77u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
inline double grad_dot(int ix, int iy, double dx, double dy, uint32_t seed) {
  const double a = (hash2(ix, iy, seed) & 0xFFFFFF) * (6.283185307179586 / 16777216.0);
  return std::cos(a) * dx + std::sin(a) * dy;
}
inline double fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
double perlin(double x, double y, uint32_t seed) {
  const int ix = int(std::floor(x)), iy = int(std::floor(y));
  const double fx = x - ix, fy = y - iy, u = fade(fx), v = fade(fy);
  const double n00 = grad_dot(ix, iy, fx, fy, seed), n10 = grad_dot(ix + 1, iy, fx - 1, fy, seed);
  const double n01 = grad_dot(ix, iy + 1, fx, fy - 1, seed), n11 = grad_dot(ix + 1, iy + 1, fx - 1, fy - 1, seed);
  return (n00 * (1 - u) + n10 * u) * (1 - v) + (n01 * (1 - u) + n11 * u) * v;
}
}  // namespace

extern "C" int rsb_terrain_generate(const rsb_terrain_properties* p, float* heights_out) {
  if (!p || !heights_out || p->x_samples < 2 || p->y_samples < 2 || p->fractal_octaves < 1) return RSB_ERR_INVALID;
  const double dx = p->x_size / (p->x_samples - 1), dy = p->y_size / (p->y_samples - 1);
  for (int iy = 0; iy < p->y_samples; iy++)
    for (int ix = 0; ix < p->x_samples; ix++) {
      double x = ix * dx * p->frequency, y = iy * dy * p->frequency, amp = 1.0, sum = 0.0;
      for (int o = 0; o < p->fractal_octaves; o++) {
        sum += amp * perlin(x, y, p->seed + 1013u * o);
        x *= p->fractal_lacunarity; y *= p->fractal_lacunarity; amp *= p->fractal_gain;
      }
      double h = p->z_scale * sum + p->height_offset;
      if (p->step_size > 0) h = std::floor(h / p->step_size) * p->step_size;     // terraced terrain
      heights_out[(size_t)iy * p->x_samples + ix] = float(h);
    }
  return RSB_OK;
}

// ---- height-map files: World::addHeightMap(raisimHeightMapFileName, centerX, centerY) and
//      World::addHeightMap(pngFileName, centerX, centerY, xSize, ySize, heightScale, heightOffset) ([RECALL]) ------------
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

extern "C" void rsb_internal_set_error(const char* msg);
namespace {
int hm_fail(const std::string& msg) { rsb_internal_set_error(msg.c_str()); return RSB_ERR_INVALID; }
uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

// text format: "xSamples ySamples xSize ySize" followed by xSamples * ySamples heights, x fastest (row = one y)
extern "C" int rsb_heightmap_read_text(const char* path, int* x_samples, int* y_samples, double* x_size, double* y_size, float* heights, int capacity) {
  if (!path || !x_samples || !y_samples || !x_size || !y_size) return hm_fail("rsb_heightmap_read_text: null argument");
  std::ifstream f(path);
  if (!f) return hm_fail(std::string("cannot open height map '") + path + "'");
  long xs = 0, ys = 0; double sx = 0, sy = 0;
  if (!(f >> xs >> ys >> sx >> sy) || xs < 2 || ys < 2 || !(sx > 0) || !(sy > 0) || xs * ys > (1L << 28))
    return hm_fail(std::string("height map '") + path + "': header must be 'xSamples ySamples xSize ySize'");
  *x_samples = int(xs); *y_samples = int(ys); *x_size = sx; *y_size = sy;
  if (!heights) return RSB_OK;                                   // size query
  if (capacity < xs * ys) return hm_fail("rsb_heightmap_read_text: output buffer too small");
  for (long i = 0; i < xs * ys; i++) {
    double h;
    if (!(f >> h)) return hm_fail(std::string("height map '") + path + "': " + std::to_string(i) + " of " + std::to_string(xs * ys) + " heights found");
    heights[i] = float(h);
  }
  return RSB_OK;
}

// 8- or 16-bit PNG (grey, grey+alpha, RGB, RGBA: the first channel is the height), non-interlaced.
// height = pixel value * height_scale + height_offset; image row r is y index r, column c is x index c.
extern "C" int rsb_heightmap_read_png(const char* path, double height_scale, double height_offset, int* x_samples, int* y_samples, float* heights,
                                      int capacity) {
  if (!path || !x_samples || !y_samples) return hm_fail("rsb_heightmap_read_png: null argument");
  std::ifstream f(path, std::ios::binary);
  if (!f) return hm_fail(std::string("cannot open height map '") + path + "'");
  std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) return hm_fail(std::string("'") + path + "' is not a PNG file");
  uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
  std::vector<unsigned char> idat;
  size_t pos = 8;
  bool have_ihdr = false;
  while (pos + 12 <= file.size()) {
    const uint32_t len = be32(&file[pos]);
    const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
    if (pos + 12 + (size_t)len > file.size()) return hm_fail(std::string("PNG '") + path + "' is truncated");
    const unsigned char* d = &file[pos + 8];
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) { w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; have_ihdr = true; }
    else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
    else if (!std::memcmp(type, "IEND", 4)) break;
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || w < 2 || h < 2 || (uint64_t)w * h > (1u << 28)) return hm_fail(std::string("PNG '") + path + "': bad or missing IHDR");
  int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!channels || (depth != 8 && depth != 16) || interlace)
    return hm_fail(std::string("PNG '") + path + "': only non-interlaced 8/16-bit grey, grey+alpha, RGB and RGBA images are supported");
  *x_samples = int(w); *y_samples = int(h);
  if (!heights) return RSB_OK;
  if ((uint64_t)capacity < (uint64_t)w * h) return hm_fail("rsb_heightmap_read_png: output buffer too small");
  const size_t bpp = size_t(channels) * (depth / 8), stride = bpp * w;
  std::vector<unsigned char> raw((stride + 1) * h);
  uLongf out_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size())
    return hm_fail(std::string("PNG '") + path + "': corrupt image data");
  std::vector<unsigned char> prev(stride, 0), cur(stride);
  for (uint32_t r = 0; r < h; r++) {
    const unsigned char* line = &raw[(stride + 1) * r];
    const int ft = line[0];
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
      int v = line[1 + i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) / 2; break;
        case 4: v += paeth(a, b, c); break;
        default: return hm_fail(std::string("PNG '") + path + "': unknown filter type");
      }
      cur[i] = (unsigned char)(v & 0xff);
    }
    for (uint32_t x = 0; x < w; x++) {
      const unsigned char* px = &cur[x * bpp];
      const double val = depth == 8 ? double(px[0]) : double((px[0] << 8) | px[1]);
      heights[(size_t)r * w + x] = float(val * height_scale + height_offset);
    }
    prev.swap(cur);
  }
  return RSB_OK;
}
