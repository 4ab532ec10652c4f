// Native image preprocessing for multimodal inputs: PIL-exact antialiased bicubic/bilinear resize (8 bpc fixed point), centre crop,
// mean/std normalisation to planar float, ViT patchification, content hashing; batch entry points run on a small thread pool.
// Parity: reference rust/sglang-mm (image decode / PIL-exact resize / patchify / normalize / hash).
#pragma once
#include <cstdint>
#include <vector>

namespace sglb_rt {

enum class Filter : int { kBilinear = 2, kBicubic = 3 };

// src: HWC uint8 (C channels, 1..4), dst: out_h x out_w x C
void resize_u8(const uint8_t* src, int h, int w, int c, uint8_t* dst, int out_h, int out_w, Filter f);

// resize shortest side to `size` (PIL rounding of the long side), centre-crop size x size, (x/255 - mean)/std -> CHW float
void clip_preprocess(const uint8_t* src, int h, int w, float* dst_chw, int size, const float mean[3], const float std_[3], Filter f);

// CHW float [3, H, W] -> [gh*gw, 3*t*p*p] rows in (merge x merge) block order (Qwen2-VL layout; t copies of the frame)
void patchify(const float* chw, int H, int W, int patch, int merge, int temporal, float* out);

uint64_t hash_bytes(const uint8_t* p, int64_t n);

}  // namespace sglb_rt
