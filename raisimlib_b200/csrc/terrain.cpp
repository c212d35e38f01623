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
