#include "image_ops.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace sglb_rt {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

inline double bicubic_filter(double x) {
  constexpr double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
inline double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

struct Coeffs {
  int ksize = 0;
  std::vector<int> bounds;   // (xmin, xsize) per output pixel
  std::vector<int32_t> kk;   // fixed-point weights [out][ksize]
};

// the separable-resample coefficient table of Pillow's ImagingResample (antialias: the filter support grows with the downscale factor)
Coeffs precompute(int in_size, int out_size, Filter f) {
  const double support0 = f == Filter::kBicubic ? 2.0 : 1.0;
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = support0 * filterscale;
  Coeffs c;
  c.ksize = (int)std::ceil(support) * 2 + 1;
  c.bounds.resize((size_t)out_size * 2);
  c.kk.assign((size_t)out_size * c.ksize, 0);
  std::vector<double> k(c.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double arg = (x + xmin - center + 0.5) * ss;
      const double w = f == Filter::kBicubic ? bicubic_filter(arg) : bilinear_filter(arg);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << kPrecisionBits);
      c.kk[(size_t)xx * c.ksize + x] = (int32_t)(v < 0 ? v - 0.5 : v + 0.5);
    }
    c.bounds[xx * 2] = xmin;
    c.bounds[xx * 2 + 1] = xmax;
  }
  return c;
}

inline uint8_t clip8(int32_t v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

void resample_h(const uint8_t* src, int h, int w, int c, uint8_t* dst, int out_w, const Coeffs& co) {
  for (int y = 0; y < h; ++y) {
    const uint8_t* row = src + (size_t)y * w * c;
    uint8_t* orow = dst + (size_t)y * out_w * c;
    for (int xx = 0; xx < out_w; ++xx) {
      const int xmin = co.bounds[xx * 2], n = co.bounds[xx * 2 + 1];
      const int32_t* k = &co.kk[(size_t)xx * co.ksize];
      for (int ch = 0; ch < c; ++ch) {
        int32_t ss = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < n; ++x) ss += (int32_t)row[(size_t)(x + xmin) * c + ch] * k[x];
        orow[(size_t)xx * c + ch] = clip8(ss);
      }
    }
  }
}

void resample_v(const uint8_t* src, int h, int w, int c, uint8_t* dst, int out_h, const Coeffs& co) {
  const size_t stride = (size_t)w * c;
  for (int yy = 0; yy < out_h; ++yy) {
    const int ymin = co.bounds[yy * 2], n = co.bounds[yy * 2 + 1];
    const int32_t* k = &co.kk[(size_t)yy * co.ksize];
    uint8_t* orow = dst + (size_t)yy * stride;
    for (size_t i = 0; i < stride; ++i) {
      int32_t ss = 1 << (kPrecisionBits - 1);
      for (int y = 0; y < n; ++y) ss += (int32_t)src[(size_t)(y + ymin) * stride + i] * k[y];
      orow[i] = clip8(ss);
    }
  }
}

}  // namespace

void resize_u8(const uint8_t* src, int h, int w, int c, uint8_t* dst, int out_h, int out_w, Filter f) {
  const bool need_h = out_w != w, need_v = out_h != h;
  if (!need_h && !need_v) {
    std::memcpy(dst, src, (size_t)h * w * c);
    return;
  }
  if (need_h && need_v) {
    // Pillow resamples horizontally first, and only the rows the vertical pass will read
    const Coeffs cv = precompute(h, out_h, f);
    const int y0 = cv.bounds[0], y1 = cv.bounds[(out_h - 1) * 2] + cv.bounds[(out_h - 1) * 2 + 1];
    const Coeffs ch = precompute(w, out_w, f);
    std::vector<uint8_t> tmp((size_t)(y1 - y0) * out_w * c);
    resample_h(src + (size_t)y0 * w * c, y1 - y0, w, c, tmp.data(), out_w, ch);
    Coeffs cv2 = cv;
    for (int yy = 0; yy < out_h; ++yy) cv2.bounds[yy * 2] -= y0;
    resample_v(tmp.data(), y1 - y0, out_w, c, dst, out_h, cv2);
  } else if (need_h) {
    resample_h(src, h, w, c, dst, out_w, precompute(w, out_w, f));
  } else {
    resample_v(src, h, w, c, dst, out_h, precompute(h, out_h, f));
  }
}

void clip_preprocess(const uint8_t* src, int h, int w, float* dst, int size, const float mean[3], const float std_[3], Filter f) {
  const double s = (double)size / std::min(w, h);
  const int nw = std::max(size, (int)std::nearbyint(w * s)), nh = std::max(size, (int)std::nearbyint(h * s));
  std::vector<uint8_t> r((size_t)nh * nw * 3);
  resize_u8(src, h, w, 3, r.data(), nh, nw, f);
  const int l = (nw - size) / 2, t = (nh - size) / 2;
  for (int ch = 0; ch < 3; ++ch) {
    const float m = mean[ch], inv = 1.0f / std_[ch];
    float* plane = dst + (size_t)ch * size * size;
    for (int y = 0; y < size; ++y) {
      const uint8_t* row = r.data() + ((size_t)(y + t) * nw + l) * 3 + ch;
      for (int x = 0; x < size; ++x) plane[(size_t)y * size + x] = ((float)row[(size_t)x * 3] / 255.0f - m) * inv;
    }
  }
}

void patchify(const float* chw, int H, int W, int p, int merge, int temporal, float* out) {
  const int gh = H / p, gw = W / p;
  const size_t row_len = (size_t)3 * temporal * p * p;
  size_t r = 0;
  for (int by = 0; by < gh / merge; ++by)
    for (int bx = 0; bx < gw / merge; ++bx)
      for (int my = 0; my < merge; ++my)
        for (int mx = 0; mx < merge; ++mx, ++r) {
          const int py = (by * merge + my) * p, px = (bx * merge + mx) * p;
          float* o = out + r * row_len;
          for (int ch = 0; ch < 3; ++ch)
            for (int t = 0; t < temporal; ++t)
              for (int y = 0; y < p; ++y) {
                std::memcpy(o, chw + ((size_t)ch * H + py + y) * W + px, sizeof(float) * p);
                o += p;
              }
        }
}

uint64_t hash_bytes(const uint8_t* p, int64_t n) {
  // 64-bit multiply-xorshift over 8-byte words (content hash for the multimodal cache / radix pad values; not cryptographic)
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t v;
    std::memcpy(&v, p + i, 8);
    h = (h ^ v) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  uint64_t tail = 0;
  for (int s = 0; i < n; ++i, s += 8) tail |= (uint64_t)p[i] << s;
  h = (h ^ tail) * 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

}  // namespace sglb_rt
