// Fused ReID stem: Conv3x3(3 -> 64, stride 1, pad 1) + bias (BN folded) + ReLU + MaxPool(3, stride 2, pad 1)
// in one kernel, tcgen05 tensor cores, sm_100a.
//
// First two layers of the DeepSORT ReID net, which the reference runs inside
// its TensorRT engine (/root/reference/src/tracker/reid_model.py:115).  Unfused, the 128x64x64
// convolution output (1 MB per crop) is written to HBM and read back by the pool: 2.3 GB per 1 100
// crops, 1.9 ms of a 9.9 ms step.  Here it lives only in shared memory.
//
// Tile = 3 x 16 pooled pixels of one crop.  They need convolution rows 2*py0-1 .. 2*py0+5 (7 rows) and
// columns 2*px0-1 .. 2*px0+31 (33 columns); those are enumerated in a raster of width RW = 35 (two junk
// columns per row), 245 positions -> two 128-row accumulators.  The input patch (9 rows x 35 columns of
// NHWC8 pixels, 16 bytes each, out-of-image pixels zero-filled by the TMA unit) is loaded ONCE; the A
// operand of filter tap (dy, dx) is that patch read through a UMMA descriptor advanced by (dy*RW + dx)
// pixels.  A pixel is exactly one 16-byte K chunk (8 channels, 3 used), so two taps form one K = 16
// MMA: the descriptor's leading-dimension byte offset is the distance between the two taps' pixels.
// 9 taps + 1 zero chunk = 5 MMAs per accumulator.
//
// Epilogue (16 warps): TMEM -> bf16 -> a shared-memory convolution tile indexed by raster position
// (positions outside the crop are written as -inf); block barrier; 48 x 8 threads each reduce one pooled
// pixel x 8 channels (nine 16-byte reads), then add the bias and apply the ReLU - both commute with the
// max, so they run on 48 pooled pixels instead of 231 convolution pixels - and store 16 bytes.  The
// convolution tile is double-buffered: one block barrier per tile.
#include <cstring>
#include <vector>

#include "conv_tc.cuh"
#include "stem_pool.cuh"
#include "tc_ptx.cuh"

namespace aicam {

extern void count_launch();
bool profile_begin(cudaStream_t st, size_t* slot);
void profile_end(cudaStream_t st, size_t slot);

namespace {

using namespace ptx;

constexpr int COUT = 64;
constexpr int PH = 3, PW = 16;                 // pooled pixels per tile
constexpr int CR = 2 * PH + 1, CC = 2 * PW + 1;  // convolution rows / columns per tile: 7 x 33
constexpr int RW = CC + 2;                     // raster width: 35
constexpr int NPOS = CR * RW;                  // 245 raster positions, two 128-row accumulators
constexpr int BH = CR + 2;                     // input rows per patch: 9
constexpr int KCHUNKS = 10;                    // 9 taps + 1 zero chunk
constexpr uint32_t BOX_BYTES = BH * RW * 16;   // 5040
constexpr uint32_t STAGE_BYTES = 6144;         // >= (255 + 2*RW + 2 + 1) * 16: junk rows stay inside the stage
constexpr int STAGES = 4;
constexpr uint32_t W_BYTES = KCHUNKS * COUT * 16;  // 10 240
constexpr uint32_t PITCH = COUT * 2 + 16;      // convolution tile row pitch (bank-conflict-free 16-byte accesses)
constexpr int EPI_WARPS = 16;
constexpr int THREADS = (EPI_WARPS + 3) * 32;
constexpr uint32_t OFF_BIAS = 256, OFF_W = 1024, OFF_A = OFF_W + W_BYTES, OFF_CONV = OFF_A + STAGES * STAGE_BYTES;
constexpr uint32_t CONV_TILE_BYTES = 256 * PITCH;
constexpr uint32_t SMEM_BYTES = OFF_CONV + 4 * CONV_TILE_BYTES;  // two epilogue groups x double-buffered convolution tile

struct StemArgs {
  const __nv_bfloat16* wgt;
  const float* bias;
  __nv_bfloat16* out;
  int h, w, ph, pw;  // convolution size (= input size), pooled size
  int out_pad;       // extra rows / columns of a zero-bordered output image ([ph + out_pad][pw + out_pad][64]) ...
  int out_lo;        // ... and the offset of the interior in it (conv_tc.cuh: pad kinds)
  int tiles_y, tiles_x;
  int batch;
  const int* batch_dev;
};

__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 r = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&r);
}

__device__ __forceinline__ void mma_issue(bool leader, uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(leader))
      : "memory");
}
__device__ __forceinline__ void commit_if(bool leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
      "r"(static_cast<uint32_t>(leader))
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(THREADS, 1) reid_stem_pool_kernel(const StemArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_a_full = sbase, bar_a_empty = sbase + 32, bar_acc_full = sbase + 64, bar_acc_empty = sbase + 80;
  const uint32_t bar_w_full = sbase + 96;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 128);
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int batch = a.batch;
  if (a.batch_dev) batch = min(batch, __ldg(a.batch_dev));
  const int tiles_per_img = a.tiles_y * a.tiles_x;
  const int total_tiles = batch * tiles_per_img;
  if (static_cast<int>(blockIdx.x) >= total_tiles) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_a_full + 8 * s, 1);
      mbar_init(bar_a_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc_full + 8 * s, 1);
      mbar_init(bar_acc_empty + 8 * s, EPI_WARPS / 2);  // one group of 8 epilogue warps per accumulator buffer, one elected arrive per warp
    }
    mbar_init(bar_w_full, 1);
    mbar_init_fence();
  }
  if (warp == EPI_WARPS) tc_alloc(smem_u32(tmem_ptr_smem), 256);
  if (warp == EPI_WARPS + 1 && lane == 0) tma_prefetch_desc(&tmap);
  if (threadIdx.x < COUT) bias_s[threadIdx.x] = __ldg(a.bias + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < EPI_WARPS) {
    // ================================================================== epilogue + pooling, two groups of 8 warps
    // Group g owns the tiles with (it & 1) == g, i.e. always accumulator buffer g and convolution tile g: while one
    // group waits for TMEM or pools, the other is in the opposite phase - the phases of consecutive tiles overlap
    // instead of all 16 warps marching through them together (one named barrier per group, 256 threads).
    const int g = warp >> 3, lw = warp & 7;
    const int wq = lw & 3;                   // TMEM lane quarter (= warp % 4)
    const int j = lw >> 2;                   // which 128-row accumulator
    const int q = j * 128 + wq * 32 + lane;  // my raster position
    const int ry = q / RW, rx = q - ry * RW;
    const uint32_t my_row_off = static_cast<uint32_t>(q) * PITCH;
    const uint32_t taddr_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + g * 128 + j * COUT;
    // pooling role: item i < 384 is pooled pixel i / 8 of the tile, channels 8 (i % 8) .. + 7; thread gt takes gt, gt + 256
    const int gt = lw * 32 + lane;
    const int pg = gt & 7;
    float pbias[8];  // bias and ReLU commute with the max: applied to the 48 pooled pixels, not the 231 convolution pixels
#pragma unroll
    for (int i = 0; i < 8; ++i) pbias[i] = bias_s[pg * 8 + i];
    int it = g;
    for (int tile = blockIdx.x + g * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, it += 2) {
      // the group's convolution tile is double-buffered: a thread that is already writing tile it + 2 cannot disturb a
      // slower thread still pooling tile it; tile it + 4 reuses the buffer only after the barrier of tile it + 2
      uint8_t* conv_s = smem + OFF_CONV + (g * 2 + ((it >> 1) & 1)) * CONV_TILE_BYTES;
      const int n = tile / tiles_per_img;
      const int r2 = tile - n * tiles_per_img;
      const int ty = r2 / a.tiles_x, tx = r2 - ty * a.tiles_x;
      const int y = 2 * ty * PH - 1 + ry, x = 2 * tx * PW - 1 + rx;  // my convolution pixel
      const bool valid = q < NPOS && rx < CC && y >= 0 && y < a.h && x >= 0 && x < a.w;
      mbar_wait(bar_acc_full + 8 * g, (it >> 1) & 1);
      tc_fence_after();
      uint32_t v0[16], v1[16], v2[16], v3[16];
      tc_ld16_nowait(taddr_lane, v0);
      tc_ld16_nowait(taddr_lane + 16, v1);
      tc_ld16_nowait(taddr_lane + 32, v2);
      tc_ld16_nowait(taddr_lane + 48, v3);
      tc_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * g);  // accumulator read: the MMA warp may reuse the buffer
      if (q < NPOS) {
        // raw accumulators as bf16; positions outside the crop lose every max (-inf)
        uint8_t* my_row = conv_s + my_row_off;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint32_t(&v)[16] = h == 0 ? v0 : (h == 1 ? v1 : (h == 2 ? v2 : v3));
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = valid ? pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])) : 0xFF80FF80u;
          *reinterpret_cast<uint4*>(my_row + h * 32) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(my_row + h * 32 + 16) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      // the group's convolution tile is complete (and every thread of the group has finished pooling the previous one)
      asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(256) : "memory");
      // pooling: one item = TWO horizontally adjacent pooled pixels x 8 channels: their windows share a column, so 15
      // shared-memory reads serve two outputs (9 each before), and the 24 pairs x 8 channel groups = 192 items fit the
      // group's 256 threads in a single round
      if (gt < PH * (PW / 2) * 8) {
        const int pr = gt >> 3;
        const int ppy = pr / (PW / 2), ppx = (pr - ppy * (PW / 2)) * 2;
        const int py = ty * PH + ppy, px = tx * PW + ppx;
        if (py < a.ph && px < a.pw) {
          const uint8_t* base = conv_s + static_cast<size_t>((2 * ppy) * RW + 2 * ppx) * PITCH + pg * 16;
          uint4 c[5];  // column maxima of the 3 x 5 window
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            uint4 m = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(dx) * PITCH);
#pragma unroll
            for (int dy = 1; dy < 3; ++dy) {
              const uint4 u = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(dy * RW + dx) * PITCH);
              m.x = hmax2(m.x, u.x); m.y = hmax2(m.y, u.y); m.z = hmax2(m.z, u.z); m.w = hmax2(m.w, u.w);
            }
            c[dx] = m;
          }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            if (px + o >= a.pw) break;
            const uint4 m0 = c[2 * o], m1 = c[2 * o + 1], m2 = c[2 * o + 2];
            const uint32_t mw[4] = {hmax2(hmax2(m0.x, m1.x), m2.x), hmax2(hmax2(m0.y, m1.y), m2.y), hmax2(hmax2(m0.z, m1.z), m2.z),
                                    hmax2(hmax2(m0.w, m1.w), m2.w)};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ow[k] = pack_bf16x2(fmaxf(bf16_lo(mw[k]) + pbias[2 * k], 0.0f), fmaxf(bf16_hi(mw[k]) + pbias[2 * k + 1], 0.0f));
            *reinterpret_cast<uint4*>(a.out + ((static_cast<long long>(n) * (a.ph + a.out_pad) + py + a.out_lo) * (a.pw + a.out_pad) + px + o + a.out_lo) * COUT + pg * 8) =
                make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
      }
    }
  } else if (warp == EPI_WARPS) {
    // ================================================================== MMA issuer (converged warp, elected lane)
    const bool leader = elect_one();
    const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    // instruction descriptor: fp32 accumulate, bf16 A/B, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(COUT >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14);  // SBO = 128 B (8 rows x 16 B), descriptor version 1, no swizzle
    const uint32_t w_lo = ((sbase + OFF_W) >> 4) | ((COUT * 16u >> 4) << 16);  // B: chunk stride = 64 rows x 16 B
    mbar_wait(bar_w_full, 0);
    uint32_t sa = 0, pa = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(bar_acc_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
      mbar_wait(bar_a_full + 8 * sa, pa);
      tc_fence_after();
      const uint32_t a_base = (sbase + OFF_A + sa * STAGE_BYTES) >> 4;  // one raster position = one 16-byte unit
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          // taps 2i and 2i + 1 (tap 9 does not exist: its weights are zero)
          const int t0 = 2 * i, t1 = 2 * i + 1;
          const int s0 = (t0 / 3) * RW + t0 % 3;
          const int s1 = t1 < 9 ? (t1 / 3) * RW + t1 % 3 : s0 + 1;
          const uint32_t a_lo = (a_base + jj * 128 + s0) | (static_cast<uint32_t>(s1 - s0) << 16);  // LBO = tap distance
          mma_issue(leader, tmem0 + buf * 128 + jj * COUT, a_lo, hi, w_lo + i * (2 * COUT), hi, idesc, i != 0 ? 1u : 0u);
        }
      }
      commit_if(leader, bar_a_empty + 8 * sa);
      commit_if(leader, bar_acc_full + 8 * buf);
      if (++sa == STAGES) { sa = 0; pa ^= 1; }
    }
    tc_fence_before();
  } else if (warp == EPI_WARPS + 1) {
    // ================================================================== patch producer
    if (lane == 0) {
      uint32_t sa = 0, pa = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int r2 = tile - n * tiles_per_img;
        const int ty = r2 / a.tiles_x, tx = r2 - ty * a.tiles_x;
        mbar_wait(bar_a_empty + 8 * sa, pa);
        mbar_arrive_expect_tx(bar_a_full + 8 * sa, BOX_BYTES);
        // input rows 2*ty*PH - 2 .., columns 2*tx*PW - 2 ..: convolution origin (-1) and its own halo (-1)
        tma_load_4d(sbase + OFF_A + sa * STAGE_BYTES, &tmap, bar_a_full + 8 * sa, 0, 2 * tx * PW - 2, 2 * ty * PH - 2, n);
        if (++sa == STAGES) { sa = 0; pa ^= 1; }
      }
    }
  } else {
    // ================================================================== weights: resident for the CTA's life
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w_full, W_BYTES);
      bulk_g2s(sbase + OFF_W, a.wgt, W_BYTES, bar_w_full);
    }
  }
  __syncthreads();
  if (warp == EPI_WARPS) {
    tc_fence_after();
    tc_dealloc(tmem_base, 256);
  }
}

// NHWC4 -> NHWC8 (RGB + zero channels): one pixel = one 16-byte K chunk of the stem's implicit GEMM
__global__ void nhwc4_to_nhwc8_kernel(const uint2* __restrict__ in, uint4* __restrict__ out, long long pixels_per_img,
                                      int batch, const int* __restrict__ n_dev) {
  if (n_dev) batch = min(batch, __ldg(n_dev));
  const long long total = pixels_per_img * batch;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint2 v = __ldg(in + i);
    out[i] = make_uint4(v.x, v.y, 0u, 0u);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline uint16_t bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace

int pack_stem_pool(const float* w_oihw, const float* bias, int cout, int cin, StemPool* out) {
  if (cout != COUT || cin < 1 || cin > 8) return fail(AICAM_ERR_UNSUPPORTED, "stem_pool: only (<= 8) -> 64 channel 3x3 stems are fused");
  // [chunk = tap][cout][8 channels]; chunk 9 stays zero
  std::vector<uint16_t> packed(static_cast<size_t>(KCHUNKS) * COUT * 8, 0);
  for (int o = 0; o < cout; ++o)
    for (int t = 0; t < 9; ++t)
      for (int c = 0; c < cin; ++c)
        packed[(static_cast<size_t>(t) * COUT + o) * 8 + c] = bf16_bits(w_oihw[(static_cast<size_t>(o) * cin + c) * 9 + t]);
  std::vector<float> b(COUT, 0.0f);
  for (int o = 0; o < cout; ++o) b[o] = bias ? bias[o] : 0.0f;
  StemPool s;
  AICAM_CUDA_OK(cudaMalloc(&s.w, packed.size() * 2));
  AICAM_CUDA_OK(cudaMalloc(&s.bias, b.size() * 4));
  AICAM_CUDA_OK(cudaMemcpy(s.w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  AICAM_CUDA_OK(cudaMemcpy(s.bias, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  *out = s;
  return AICAM_OK;
}

void free_stem_pool(StemPool* s) {
  if (s->w) cudaFree(s->w);
  if (s->bias) cudaFree(s->bias);
  s->w = nullptr;
  s->bias = nullptr;
}

int launch_nhwc4_to_nhwc8(const __nv_bfloat16* in, int batch, int h, int w, __nv_bfloat16* out, const int* n_dev,
                          cudaStream_t stream) {
  if (batch <= 0) return AICAM_OK;
  const long long ppi = static_cast<long long>(h) * w;
  const long long total = ppi * batch;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 16));
  nhwc4_to_nhwc8_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint2*>(in), reinterpret_cast<uint4*>(out), ppi, batch, n_dev);
  count_launch();
  return last_launch("nhwc4_to_nhwc8_kernel");
}

int launch_stem_pool(const StemPool& sp, const __nv_bfloat16* in_nhwc8, int batch, int h, int w, const int* n_dev,
                     __nv_bfloat16* out, cudaStream_t stream, int out_pad) {
  if (batch <= 0) return AICAM_OK;
  if (encode_tiled() == nullptr) return fail(AICAM_ERR_CUDA, "stem_pool: cuTensorMapEncodeTiled is not available");
  if (h % 2 || w % 2) return fail(AICAM_ERR_INVALID_ARG, "stem_pool: even input sizes only");
  StemArgs a;
  a.wgt = sp.w; a.bias = sp.bias; a.out = out;
  a.h = h; a.w = w; a.ph = h / 2; a.pw = w / 2; a.out_pad = pad_ext(out_pad); a.out_lo = pad_lo(out_pad);
  a.tiles_y = cdiv(a.ph, PH); a.tiles_x = cdiv(a.pw, PW);
  a.batch = batch; a.batch_dev = n_dev;
  alignas(64) CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  const cuuint64_t dims[4] = {8, static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(batch)};
  const cuuint64_t strides[3] = {16, static_cast<cuuint64_t>(w) * 16, static_cast<cuuint64_t>(h) * w * 16};
  const cuuint32_t box[4] = {8, RW, BH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(in_nhwc8), dims, strides, box,
                                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(AICAM_ERR_CUDA, "stem_pool: cuTensorMapEncodeTiled failed with " + std::to_string(static_cast<int>(cr)));
  if (int rc = ensure_dynamic_smem(reid_stem_pool_kernel, SMEM_BYTES)) return rc;
  const int num_sms = current_num_sms();
  const long long tiles = static_cast<long long>(batch) * a.tiles_y * a.tiles_x;
  dim3 grid(static_cast<unsigned>(std::min<long long>(tiles, num_sms)));
  size_t slot = 0;
  const bool prof = profile_begin(stream, &slot);
  reid_stem_pool_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(a, tmap);
  if (prof) profile_end(stream, slot);
  count_launch();
  return last_launch("reid_stem_pool_kernel");
}

}  // namespace aicam
