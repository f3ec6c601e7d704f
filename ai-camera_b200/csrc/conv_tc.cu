// Implicit-GEMM convolution (+bias, SiLU/ReLU, residual) on tcgen05 tensor cores, sm_100a.
//
// This is the operator behind every Conv+BN+SiLU / C2f / SPPF / Detect layer of YOLOv8 and
// every BasicBlock of the DeepSORT ReID net - the work the reference hands to a TensorRT
// engine (/root/reference/src/trt_utils/trt_engine.py:191, called from
// src/detector/yolo_detector.py:97 and src/tracker/reid_model.py:115).
//
// GEMM view:  D[m][n] = sum_k A[m][k] * B[n][k]
//   m = output pixel (batch, oy, ox) flattened            tile M = 128 (one UMMA, cta_group::1)
//   n = output channel                                    tile N = n_tile (16..128, divides cout_pad)
//   k = (tap, input channel)                              staged 64 at a time (8 x 16-byte chunks)
// Operands are bf16, accumulation is fp32 in TMEM.  Both operands sit in shared memory in the
// canonical K-major NO-SWIZZLE UMMA layout: 8-row x 16-byte core matrices, rows of one chunk
// contiguous:  addr(row, chunk) = chunk * (ROWS*16) + row*16   (SBO = 128 B, LBO = ROWS*16 B).
//
// Warp roles (192 threads):
//   warps 0-3  im2col gather: thread t owns output pixel m0+t and cp.async's its 16-byte
//              channel chunks (zero-filled outside the image) straight into the UMMA layout;
//              afterwards the same warps run the epilogue (TMEM -> registers -> bias/act/
//              residual -> bf16/fp32 NHWC stores); warp w reads TMEM lanes 32w..32w+31.
//   warp 4     allocates TMEM, issues tcgen05.mma (one elected lane), commits to mbarriers.
//   warp 5     streams the packed weights with cp.async.bulk (1-D bulk copy, mbarrier tx).
// Pipeline: 3 shared-memory stages, full/empty mbarriers; several CTAs are resident per SM so
// one CTA's epilogue overlaps another's main loop.
#include "conv_tc.cuh"

#include <vector>

namespace aicam {

namespace {

constexpr int TILE_M = 128;
constexpr int STAGES = 3;
constexpr int CHUNKS_PER_STAGE = 8;                 // 8 x 16 B = 64 bf16 of K per stage
constexpr int A_STAGE_BYTES = TILE_M * 16 * CHUNKS_PER_STAGE;  // 16 KiB
constexpr int A_CHUNK_BYTES = TILE_M * 16;          // LBO of A
constexpr int SMEM_HEADER = 256;
constexpr int NUM_THREADS = 192;
constexpr int GATHER_LAG = 2;                       // cp.async groups kept in flight per thread

struct ConvKernelArgs {
  const __nv_bfloat16* in;
  long long in_img_stride;
  int in_cstride, in_coff;
  int h, w, ho, wo, howo, m_total;
  int ksize, stride, pad;
  int cin_chunks;   // cin_pad / 8 (0 in stem mode)
  int stem;         // 1: cin_pad == 4, a chunk is two taps of 4 channels
  int taps;
  int q, q_pad;     // real / padded number of 16-byte K chunks
  const __nv_bfloat16* wgt;
  const float* bias;
  int cout, cout_pad, n_tile;
  void* out;
  long long out_img_stride;
  int out_cstride, out_coff, out_f32;
  const __nv_bfloat16* res;
  long long res_img_stride;
  int res_cstride, res_coff, res_mode;
  int act;
  uint32_t idesc;
  uint32_t tmem_cols;
  const int* batch_dev;  // optional: images actually present (device), m_total is the capacity
};

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) {
      printf("aicam conv: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, no swizzle: start >> 4 | LBO >> 4 (bits 16..29) | SBO >> 4 (bits 32..45) | version 1 (bit 46)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return x * __frcp_rn(1.0f + __expf(-x));
  if (act == 2) return fmaxf(x, 0.0f);
  return x;
}

__global__ void __launch_bounds__(NUM_THREADS) conv_tc_kernel(const ConvKernelArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base;            // STAGES x 8 B
  const uint32_t bar_empty = smem_base + 64;      // STAGES x 8 B
  const uint32_t bar_tmem_full = smem_base + 128;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 136);
  const uint32_t smem_a = smem_base + SMEM_HEADER;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(a.n_tile) * 16u * CHUNKS_PER_STAGE;
  const uint32_t smem_b = smem_a + STAGES * A_STAGE_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TILE_M;
  const int n0 = blockIdx.y * a.n_tile;
  const int num_kb = (a.q_pad + CHUNKS_PER_STAGE - 1) / CHUNKS_PER_STAGE;
  int m_total = a.m_total;
  if (a.batch_dev) m_total = min(m_total, __ldg(a.batch_dev) * a.howo);
  if (m0 >= m_total) return;  // whole CTA: tile beyond the (device-side) batch

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, TILE_M + 1);  // 128 gather threads + the weight loader's expect_tx
      mbar_init(bar_empty + 8 * s, 1);          // one tcgen05.commit
    }
    mbar_init(bar_tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    // ------------------------------------------------------------------ im2col gather
    const int row = threadIdx.x;  // 0..127
    const int m = m0 + row;
    const bool valid = m < m_total;
    int n_img = 0, rem = 0, oy = 0, ox = 0;
    if (valid) {
      n_img = m / a.howo;
      rem = m - n_img * a.howo;
      oy = rem / a.wo;
      ox = rem - oy * a.wo;
    }
    const int iy0 = oy * a.stride - a.pad;
    const int ix0 = ox * a.stride - a.pad;
    const __nv_bfloat16* in_img = a.in + static_cast<long long>(n_img) * a.in_img_stride + a.in_coff;
    const uint32_t dst_row = row * 16;

    int tap = 0, tr = 0, tc = 0, c8 = 0;  // running (tap, channel chunk) of the next K chunk
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const int it = kb / STAGES;
      mbar_wait(bar_empty + 8 * s, (it & 1) ^ 1);
      const uint32_t dst_stage = smem_a + s * A_STAGE_BYTES + dst_row;
      const int nchunks = min(CHUNKS_PER_STAGE, a.q_pad - kb * CHUNKS_PER_STAGE);
      if (!a.stem) {
        for (int j = 0; j < nchunks; ++j) {
          const int qi = kb * CHUNKS_PER_STAGE + j;
          const int iy = iy0 + tr, ix = ix0 + tc;
          const bool ok = valid && qi < a.q && iy >= 0 && iy < a.h && ix >= 0 && ix < a.w;
          const __nv_bfloat16* src =
              ok ? in_img + (static_cast<long long>(iy) * a.w + ix) * a.in_cstride + c8 * 8 : a.in;
          cp_async_16(dst_stage + j * A_CHUNK_BYTES, src, ok ? 16u : 0u);
          if (++c8 == a.cin_chunks) {
            c8 = 0;
            ++tap;
            if (++tc == a.ksize) { tc = 0; ++tr; }
          }
        }
      } else {
        for (int j = 0; j < nchunks; ++j) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int iy = iy0 + tr, ix = ix0 + tc;
            const bool ok = valid && tap < a.taps && iy >= 0 && iy < a.h && ix >= 0 && ix < a.w;
            const __nv_bfloat16* src =
                ok ? in_img + (static_cast<long long>(iy) * a.w + ix) * a.in_cstride : a.in;
            cp_async_8(dst_stage + j * A_CHUNK_BYTES + half * 8, src, ok ? 8u : 0u);
            ++tap;
            if (++tc == a.ksize) { tc = 0; ++tr; }
          }
        }
      }
      cp_async_commit();
      if (kb >= GATHER_LAG) {
        cp_async_wait<GATHER_LAG>();
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * ((kb - GATHER_LAG) % STAGES));
      }
    }
    if (num_kb >= 2) {
      cp_async_wait<1>();
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * ((num_kb - 2) % STAGES));
    }
    cp_async_wait<0>();
    fence_proxy_async();
    mbar_arrive(bar_full + 8 * ((num_kb - 1) % STAGES));

    // ------------------------------------------------------------------ epilogue
    mbar_wait(bar_tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const long long out_pix = static_cast<long long>(n_img) * a.out_img_stride +
                              static_cast<long long>(rem) * a.out_cstride + a.out_coff + n0;
    const long long res_pix = static_cast<long long>(n_img) * a.res_img_stride +
                              static_cast<long long>(rem) * a.res_cstride + a.res_coff + n0;
    for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
      uint32_t v[16];
      tc_ld16(taddr_row + c0, v);
      if (!valid) continue;
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]) + __ldg(a.bias + n0 + c0 + i);
      const int cvalid = min(16, a.cout - (n0 + c0));  // <= 0 when the group is channel padding
      if (cvalid <= 0) continue;
      float r[16];
      if (a.res_mode != 0) {
        const __nv_bfloat16* rp = a.res + res_pix + c0;
        if (cvalid == 16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
          const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rp));
          const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) { r[2 * i] = bf16_lo(rw[i]); r[2 * i + 1] = bf16_hi(rw[i]); }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = i < cvalid ? __bfloat162float(rp[i]) : 0.0f;
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (a.res_mode == 2) x[i] += r[i];
        x[i] = apply_act(x[i], a.act);
        if (a.res_mode == 1) x[i] += r[i];
      }
      if (a.out_f32) {
        float* op = reinterpret_cast<float*>(a.out) + out_pix + c0;
        if (cvalid == 16 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(op)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        } else {
          for (int i = 0; i < cvalid; ++i) op[i] = x[i];
        }
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + out_pix + c0;
        if (cvalid == 16 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
          uint4 o0, o1;
          o0.x = pack_bf16x2(x[0], x[1]);   o0.y = pack_bf16x2(x[2], x[3]);
          o0.z = pack_bf16x2(x[4], x[5]);   o0.w = pack_bf16x2(x[6], x[7]);
          o1.x = pack_bf16x2(x[8], x[9]);   o1.y = pack_bf16x2(x[10], x[11]);
          o1.z = pack_bf16x2(x[12], x[13]); o1.w = pack_bf16x2(x[14], x[15]);
          reinterpret_cast<uint4*>(op)[0] = o0;
          reinterpret_cast<uint4*>(op)[1] = o1;
        } else {
          for (int i = 0; i < cvalid; ++i) op[i] = __float2bfloat16_rn(x[i]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t b_chunk_bytes = static_cast<uint32_t>(a.n_tile) * 16u;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const int it = kb / STAGES;
      mbar_wait(bar_full + 8 * s, it & 1);
      tc_fence_after();
      if (lane == 0) {
        const int nchunks = min(CHUNKS_PER_STAGE, a.q_pad - kb * CHUNKS_PER_STAGE);
        const uint32_t a_stage = smem_a + s * A_STAGE_BYTES;
        const uint32_t b_stage = smem_b + s * b_stage_bytes;
        for (int kk = 0; kk < nchunks / 2; ++kk) {
          const uint64_t da = make_smem_desc(a_stage + kk * 2 * A_CHUNK_BYTES, A_CHUNK_BYTES, 128);
          const uint64_t db = make_smem_desc(b_stage + kk * 2 * b_chunk_bytes, b_chunk_bytes, 128);
          tc_mma_bf16(tmem_base, da, db, a.idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        tc_commit(bar_empty + 8 * s);
        if (kb == num_kb - 1) tc_commit(bar_tmem_full);
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      const uint32_t b_chunk_bytes = static_cast<uint32_t>(a.n_tile) * 16u;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const int it = kb / STAGES;
        mbar_wait(bar_empty + 8 * s, (it & 1) ^ 1);
        const int nchunks = min(CHUNKS_PER_STAGE, a.q_pad - kb * CHUNKS_PER_STAGE);
        const uint32_t bar = bar_full + 8 * s;
        const uint32_t dst = smem_b + s * b_stage_bytes;
        mbar_arrive_expect_tx(bar, nchunks * b_chunk_bytes);
        const __nv_bfloat16* src = a.wgt + (static_cast<long long>(kb) * CHUNKS_PER_STAGE * a.cout_pad + n0) * 8;
        if (a.n_tile == a.cout_pad) {
          bulk_g2s(dst, src, nchunks * b_chunk_bytes, bar);
        } else {
          for (int j = 0; j < nchunks; ++j)
            bulk_g2s(dst + j * b_chunk_bytes, src + static_cast<long long>(j) * a.cout_pad * 8, b_chunk_bytes, bar);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

int pick_n_tile(int cout_pad) {
  for (int nt = 128; nt >= 16; nt -= 16)
    if (cout_pad % nt == 0) return nt;
  return 16;
}

}  // namespace

extern void count_launch();
bool profile_begin(cudaStream_t st, size_t* slot);
void profile_end(cudaStream_t st, size_t slot);

int pack_conv_weights(const float* w, const float* bias, int cout, int cin, int ksize, int stride,
                      PackedConv* out) {
  if (cout <= 0 || cin <= 0 || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2))
    return fail(AICAM_ERR_INVALID_ARG, "pack_conv_weights: unsupported convolution shape");
  const int taps = ksize * ksize;
  const bool stem = cin <= 4;
  const int cin_pad = stem ? 4 : (cin + 7) / 8 * 8;
  const int k_total = taps * cin_pad;
  const int q = (k_total + 7) / 8;
  const int q_pad = (q + 1) / 2 * 2;
  const int cout_pad = (cout + 15) / 16 * 16;
  std::vector<uint16_t> packed(static_cast<size_t>(q_pad) * cout_pad * 8, 0);
  for (int o = 0; o < cout; ++o)
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < cin; ++c) {
        const int k = t * cin_pad + c;
        const float v = w[(static_cast<size_t>(o) * cin + c) * taps + t];
        packed[(static_cast<size_t>(k / 8) * cout_pad + o) * 8 + (k % 8)] = f32_to_bf16_bits(v);
      }
  std::vector<float> b(cout_pad, 0.0f);
  for (int o = 0; o < cout; ++o) b[o] = bias ? bias[o] : 0.0f;
  PackedConv p;
  p.cin = cin; p.cin_pad = cin_pad; p.cout = cout; p.ksize = ksize; p.stride = stride; p.q = q; p.q_pad = q_pad;
  AICAM_CUDA_OK(cudaMalloc(&p.w, packed.size() * 2));
  AICAM_CUDA_OK(cudaMalloc(&p.bias, b.size() * 4));
  AICAM_CUDA_OK(cudaMemcpy(p.w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  AICAM_CUDA_OK(cudaMemcpy(p.bias, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  *out = p;
  return AICAM_OK;
}

void free_packed_conv(PackedConv* p) {
  if (p->w) cudaFree(p->w);
  if (p->bias) cudaFree(p->bias);
  p->w = nullptr;
  p->bias = nullptr;
}

int launch_conv(const PackedConv& pc, const ConvLaunch& L, cudaStream_t stream) {
  ConvKernelArgs a;
  a.in = L.in; a.in_img_stride = L.in_img_stride; a.in_cstride = L.in_cstride; a.in_coff = L.in_coff;
  a.h = L.h; a.w = L.w; a.ho = L.ho; a.wo = L.wo; a.howo = L.ho * L.wo;
  a.m_total = L.batch * a.howo;
  a.ksize = pc.ksize; a.stride = pc.stride; a.pad = pc.ksize / 2;
  a.stem = pc.cin_pad == 4 ? 1 : 0;
  a.cin_chunks = a.stem ? 0 : pc.cin_pad / 8;
  a.taps = pc.ksize * pc.ksize;
  a.q = pc.q; a.q_pad = pc.q_pad;
  a.wgt = pc.w; a.bias = pc.bias;
  a.cout = pc.cout; a.cout_pad = (pc.cout + 15) / 16 * 16;
  a.n_tile = pick_n_tile(a.cout_pad);
  a.out = L.out; a.out_img_stride = L.out_img_stride; a.out_cstride = L.out_cstride; a.out_coff = L.out_coff;
  a.out_f32 = L.out_f32;
  a.res = L.res; a.res_img_stride = L.res_img_stride; a.res_cstride = L.res_cstride; a.res_coff = L.res_coff;
  a.res_mode = L.res ? L.res_mode : 0;
  a.act = L.act;
  a.batch_dev = L.batch_dev;
  // instruction descriptor: fp32 accumulate, bf16 A/B, both K-major, N = n_tile, M = 128
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a.n_tile >> 3) << 17) |
            (static_cast<uint32_t>(TILE_M >> 4) << 24);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(a.n_tile)) cols <<= 1;
  a.tmem_cols = cols;
  if (a.m_total <= 0) return AICAM_OK;
  if (!a.stem && (L.in_cstride % 8 != 0 || L.in_coff % 8 != 0))
    return fail(AICAM_ERR_INVALID_ARG, "launch_conv: input channel stride/offset must be multiples of 8");
  if (a.stem && (L.in_cstride != 4 || L.in_coff != 0))
    return fail(AICAM_ERR_INVALID_ARG, "launch_conv: stem input must be NHWC4");
  const size_t smem = SMEM_HEADER + STAGES * (A_STAGE_BYTES + static_cast<size_t>(a.n_tile) * 16 * CHUNKS_PER_STAGE);
  static bool attr_set = false;
  if (!attr_set) {
    AICAM_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    attr_set = true;
  }
  dim3 grid(cdiv(a.m_total, TILE_M), a.cout_pad / a.n_tile);
  size_t slot = 0;
  const bool prof = profile_begin(stream, &slot);
  conv_tc_kernel<<<grid, NUM_THREADS, smem, stream>>>(a);
  if (prof) profile_end(stream, slot);
  count_launch();
  return last_launch("conv_tc_kernel");
}

}  // namespace aicam
