// Training-side kernels of the trunk (reference: autograd through modelling/backbones/resnet.py:67-87,122-133
// in train mode -- torch.nn.Conv2d weight gradients, BatchNorm2d batch statistics and their backward).
//
// 1. conv2d weight gradient as a tcgen05 GEMM over the pixel dimension:
//      dW[co][tap][ci] = sum_pixels dy[pixel][co] * x[pixel + tap][ci]
//    Both operands are the SAME TMA boxes the forward uses ([128 pixels][64 channels], 128-byte rows,
//    SWIZZLE_128B) -- read by the tensor core as MN-major operands (instruction-descriptor bits 15/16):
//    rows are the K (pixel) dimension, the 64 channels of a row the M / N dimension.  The reduction is long
//    (N*Ho*Wo pixels) and the output small, so the pixel range is split across CTAs; fp32 partial tiles are
//    reduced in a fixed order by a second kernel (deterministic, no atomics).
#include <cuda_fp16.h>
#include <stdint.h>

#include <algorithm>

#include "common.h"
#include "umma.cuh"

namespace ctl {

struct ConvTapW {
  int map;  // which activation tensor map (stride-2 parity view)
  int dh, dw;
};

static constexpr int WG_THREADS = 192;            // TMA warp, MMA warp, 4 epilogue warps
static constexpr int WG_BOX_BYTES = 128 * 64 * 2;  // one [128 px][64 ch] box
static constexpr int WG_STAGES = 3;
static constexpr int WG_STAGE_BYTES = 4 * WG_BOX_BYTES;  // dy: 2 boxes (128 cout), x: up to 2 boxes (128 cin)
static constexpr size_t WG_SMEM = 1024 + WG_STAGES * WG_STAGE_BYTES + 256;

struct WgradParams {
  CUtensorMap x_map[4];
  CUtensorMap dy_map;
  ConvTapW taps[9];
  int n_taps, cin, cout;
  int TW, TH, tiles_w, tiles_h, m_tiles;
  int bnw;         // cin per work item: 64 or 128
  int cin_chunks;  // cin / bnw
  int cout_tiles;  // ceil(cout / 128)
  int n_items;     // cout_tiles * n_taps * cin_chunks
  int splits;      // pixel-range splits
  int cout_pad;    // cout_tiles * 128
  float* part;     // [splits][cout_pad][n_taps * cin]
};

// MN-major SWIZZLE_128B operand: 128-byte rows = 64 M/N elements, consecutive rows = consecutive K; `lbo` = byte
// distance between 64-element M/N blocks, 8-row K groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + WG_STAGES * WG_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * WG_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * WG_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * WG_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.dy_map);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.x_map[i]);
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n_units = p.n_items * p.splits;
  const int xboxes = p.bnw / 64;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  // unit -> (work item, pixel-tile range)
  auto unit_range = [&](int u, int& item, int& t0, int& t1) {
    item = u / p.splits;
    const int s = u - item * p.splits;
    t0 = (int)((long long)p.m_tiles * s / p.splits);
    t1 = (int)((long long)p.m_tiles * (s + 1) / p.splits);
  };
  auto item_coords = [&](int item, int& ct, int& tap, int& chunk) {
    ct = item / (p.n_taps * p.cin_chunks);
    const int r = item - ct * (p.n_taps * p.cin_chunks);
    tap = r / p.cin_chunks;
    chunk = r - tap * p.cin_chunks;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        int item, t0, t1, ct, tapi, chunk;
        unit_range(u, item, t0, t1);
        item_coords(item, ct, tapi, chunk);
        const ConvTapW tap = p.taps[tapi];
        int img = t0 / tiles_per_img;
        int tr = t0 - img * tiles_per_img;
        int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
        for (int t = t0; t < t1; ++t) {
          const int h0 = th * p.TH, w0 = tw * p.TW;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t dst = smem_base + stage * WG_STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), (2 + xboxes) * WG_BOX_BYTES);
          tma_load_4d(dst, &p.dy_map, full_bar(stage), ct * 128, w0, h0, img);
          tma_load_4d(dst + WG_BOX_BYTES, &p.dy_map, full_bar(stage), ct * 128 + 64, w0, h0, img);  // OOB channels -> 0
          for (int j = 0; j < xboxes; ++j)
            tma_load_4d(dst + (2 + j) * WG_BOX_BYTES, &p.x_map[tap.map], full_bar(stage), chunk * p.bnw + 64 * j,
                        w0 + tap.dw, h0 + tap.dh, img);
          if (++stage == WG_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
          if (++tw == p.tiles_w) {
            tw = 0;
            if (++th == p.tiles_h) {
              th = 0;
              ++img;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, (uint32_t)p.bnw) | (1u << 15) | (1u << 16);  // A and B MN-major
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        int item, t0, t1;
        unit_range(u, item, t0, t1);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * 128;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a0 = smem_base + stage * WG_STAGE_BYTES, b0 = a0 + 2 * WG_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 16 pixels (rows) per MMA
            const uint64_t da = make_sw128_mnmajor_desc(a0 + k * 2048, WG_BOX_BYTES);
            const uint64_t db = make_sw128_mnmajor_desc(b0 + k * 2048, WG_BOX_BYTES);
            umma_f16(acc, da, db, idesc, (t > t0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == WG_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) {
          as = 0;
          aphase ^= 1u;
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // cout inside the tile == TMEM lane
    const int ktot = p.n_taps * p.cin;
    int as = 0;
    uint32_t aphase = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      int item, t0, t1, ct, tapi, chunk;
      unit_range(u, item, t0, t1);
      item_coords(item, ct, tapi, chunk);
      const int split = u - item * p.splits;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      float* dst = p.part + ((size_t)split * p.cout_pad + ct * 128 + row) * ktot + tapi * p.cin + chunk * p.bnw;
      const uint32_t taddr = tmem_base + as * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int c = 0; c < p.bnw; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + c + 4 * q) =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                          __uint_as_float(r[4 * q + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// dW[row][col] = sum over splits (fixed order) of the fp32 partial tiles; rows >= cout are padding
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int splits, int cout_pad,
                                                           int cout, int ktot, float* __restrict__ dw) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n4 = (size_t)cout * ktot / 4;
  const size_t stride4 = (size_t)cout_pad * ktot / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(part)[i];
    for (int s = 1; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(part)[i + s * stride4];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    reinterpret_cast<float4*>(dw)[i] = acc;
  }
}

static int wgrad_plan(int n, int h, int w, int cin, int cout, int ksize, int stride, WgradParams* p) {
  const int pad = ksize == 3 ? 1 : 0;
  const int Ho = (h + 2 * pad - ksize) / stride + 1, Wo = (w + 2 * pad - ksize) / stride + 1;
  pick_tile(Ho, Wo, &p->TH, &p->TW);
  p->tiles_h = (Ho + p->TH - 1) / p->TH;
  p->tiles_w = (Wo + p->TW - 1) / p->TW;
  p->m_tiles = n * p->tiles_h * p->tiles_w;
  p->n_taps = ksize * ksize;
  p->cin = cin;
  p->cout = cout;
  p->bnw = cin % 128 == 0 ? 128 : 64;
  p->cin_chunks = cin / p->bnw;
  p->cout_tiles = (cout + 127) / 128;
  p->cout_pad = p->cout_tiles * 128;
  p->n_items = p->cout_tiles * p->n_taps * p->cin_chunks;
  const int want = (2 * sm_count() + p->n_items - 1) / p->n_items;
  p->splits = std::max(1, std::min(want, p->m_tiles));
  return Ho * 65536 + Wo;
}

}  // namespace ctl

using namespace ctl;

extern "C" {

size_t ctl_conv2d_wgrad_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize,
                                        int32_t stride) {
  if (n < 1 || h < 1 || w < 1 || cin % 64 != 0 || cout % 64 != 0 || (ksize != 1 && ksize != 3) ||
      (stride != 1 && stride != 2))
    return 0;
  WgradParams p = {};
  wgrad_plan(n, h, w, cin, cout, ksize, stride, &p);
  return (size_t)p.splits * p.cout_pad * p.n_taps * cin * sizeof(float) + 256;
}

int ctl_conv2d_wgrad_nhwc_f16(const void* x, int32_t n, int32_t h, int32_t w, int32_t cin, const void* dy, int32_t cout,
                              int32_t ksize, int32_t stride, void* workspace, size_t workspace_bytes, float* dw,
                              ctl_stream_t stream) {
  CTL_CHECK_ARG(x && dy && dw && workspace, "null pointer");
  CTL_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "bad activation shape");
  CTL_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0, "Cin=%d and Cout=%d must be multiples of 64", cin, cout);
  CTL_CHECK_ARG((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "only 1x1 / 3x3, stride 1 / 2");
  CTL_CHECK_ARG(stride == 1 || (h % 2 == 0 && w % 2 == 0), "stride 2 needs even H, W (got %dx%d)", h, w);
  int rc = ctl_device_check();
  if (rc) return rc;
  WgradParams p = {};
  const int hw = wgrad_plan(n, h, w, cin, cout, ksize, stride, &p);
  const int Ho = hw >> 16, Wo = hw & 65535;
  const size_t need = (size_t)p.splits * p.cout_pad * p.n_taps * cin * sizeof(float);
  CTL_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  CTL_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dw) & 15u) == 0,
                "workspace and dw must be 16-byte aligned");
  p.part = static_cast<float*>(workspace);
  const int pad = ksize == 3 ? 1 : 0;
  const __half* xb = static_cast<const __half*>(x);
  const uint32_t abox[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
  if (stride == 1) {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strd[4] = {2, (uint64_t)cin * 2, (uint64_t)w * cin * 2, (uint64_t)h * w * cin * 2};
    for (int i = 0; i < 4; ++i)
      if ((rc = encode_tensor_map(&p.x_map[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4, xb, dims, strd, abox,
                                  CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) p.taps[r * ksize + s] = ConvTapW{0, r - pad, s - pad};
  } else {
    // parity views (as the forward): view (ph, pw) holds input pixels (2i + ph, 2j + pw)
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)(w / 2), (uint64_t)(h / 2), (uint64_t)n};
    const uint64_t strd[4] = {2, (uint64_t)cin * 4, (uint64_t)w * cin * 4, (uint64_t)h * w * cin * 2};
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        if ((rc = encode_tensor_map(&p.x_map[ph * 2 + pw], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4,
                                    xb + ((size_t)ph * w + pw) * cin, dims, strd, abox, CU_TENSOR_MAP_SWIZZLE_128B)))
          return rc;
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) {
        const int ar = r - pad, as = s - pad;
        const int ph = ((ar % 2) + 2) % 2, pw = ((as % 2) + 2) % 2;
        p.taps[r * ksize + s] = ConvTapW{ph * 2 + pw, (ar - ph) / 2, (as - pw) / 2};
      }
  }
  {
    const uint64_t odims[4] = {(uint64_t)cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)n};
    const uint64_t ostr[4] = {2, (uint64_t)cout * 2, (uint64_t)Wo * cout * 2, (uint64_t)Ho * Wo * cout * 2};
    if ((rc = encode_tensor_map(&p.dy_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4, dy, odims, ostr, abox,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    CTL_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = std::min(p.n_items * p.splits, sm_count());
  CTL_CUDA(launch_k(conv_wgrad_kernel, dim3(grid), dim3(WG_THREADS), WG_SMEM, st, p));
  const size_t n4 = (size_t)cout * p.n_taps * cin / 4;
  const int rgrid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)sm_count() * 8);
  CTL_CUDA(launch_k(wgrad_reduce_kernel, dim3(rgrid), dim3(256), 0, st, (const float*)p.part, p.splits, p.cout_pad, (int)cout,
                    p.n_taps * (int)cin, dw));
  return 0;
}

}  // extern "C"
