// Fused ReID stem (conv3x3 3->64 + ReLU + maxpool 3x3 s2) on tcgen05; see stem_pool.cu.
#pragma once
#include "common.cuh"

namespace aicam {

struct StemPool {
  __nv_bfloat16* w = nullptr;  // device, [10 chunks = 9 taps + zero][64][8] bf16
  float* bias = nullptr;       // device, [64]
};

int pack_stem_pool(const float* w_oihw, const float* bias, int cout, int cin, StemPool* out);
void free_stem_pool(StemPool* s);
// in: bf16 [batch][h][w][4] -> out: bf16 [batch][h][w][8] (zero upper channels)
int launch_nhwc4_to_nhwc8(const __nv_bfloat16* in, int batch, int h, int w, __nv_bfloat16* out, const int* n_dev,
                          cudaStream_t stream);
// in: bf16 NHWC8 [batch][h][w][8]; out: bf16 [batch][h/2][w/2][64], or with out_pad the zero-bordered
// [batch][h/2 + 2][w/2 + 2][64] (interior at (1, 1), border untouched)
int launch_stem_pool(const StemPool& sp, const __nv_bfloat16* in_nhwc8, int batch, int h, int w, const int* n_dev,
                     __nv_bfloat16* out, cudaStream_t stream, int out_pad = 0);

}  // namespace aicam
