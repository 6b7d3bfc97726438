// Per-identity centroid mean: a segmented (CSR) row reduction, HBM-bound.
// Replaces modelling/bases.py:92-95 (_calculate_centroids), the tensor part of
// modelling/bases.py:179-262 and inference/inference_utils.py:147-159.
//
// Layout: x [n, d] fp32 row-major.  One CTA column-slab per segment: thread t owns 4
// consecutive columns (one 16-byte load per row, fully coalesced across the warp) and adds the
// segment's rows in index order, so every output element has ONE deterministic summation
// order (row order, fp32) followed by one IEEE division by the row count -- the same
// `sum / length` the reference computes.
// Algorithmic bytes: 4*d per gathered row read + 4*d per centroid written.
#include "common.h"

namespace ctl {

__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ x, int d,
                                                           const long long* __restrict__ indptr,
                                                           const long long* __restrict__ indices,
                                                           float* __restrict__ out) {
  const long long seg = blockIdx.x;
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (col >= d) return;
  const long long beg = indptr[seg], end = indptr[seg + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  long long j = beg;
  // two rows in flight per thread to cover HBM latency without changing the add order
  for (; j + 1 < end; j += 2) {
    const long long r0 = indices ? indices[j] : j;
    const long long r1 = indices ? indices[j + 1] : j + 1;
    const float4 a = *reinterpret_cast<const float4*>(x + r0 * d + col);
    const float4 b = *reinterpret_cast<const float4*>(x + r1 * d + col);
    acc.x = __fadd_rn(__fadd_rn(acc.x, a.x), b.x);
    acc.y = __fadd_rn(__fadd_rn(acc.y, a.y), b.y);
    acc.z = __fadd_rn(__fadd_rn(acc.z, a.z), b.z);
    acc.w = __fadd_rn(__fadd_rn(acc.w, a.w), b.w);
  }
  if (j < end) {
    const long long r0 = indices ? indices[j] : j;
    const float4 a = *reinterpret_cast<const float4*>(x + r0 * d + col);
    acc.x = __fadd_rn(acc.x, a.x);
    acc.y = __fadd_rn(acc.y, a.y);
    acc.z = __fadd_rn(acc.z, a.z);
    acc.w = __fadd_rn(acc.w, a.w);
  }
  const float cnt = (float)(end - beg);
  float4 o;
  if (end > beg) {
    o.x = __fdiv_rn(acc.x, cnt);
    o.y = __fdiv_rn(acc.y, cnt);
    o.z = __fdiv_rn(acc.z, cnt);
    o.w = __fdiv_rn(acc.w, cnt);
  } else {
    o = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  *reinterpret_cast<float4*>(out + seg * d + col) = o;
}

}  // namespace ctl

extern "C" int ctl_segment_mean(const float* x, int64_t n, int32_t d, const int64_t* indptr, const int64_t* indices,
                                int64_t n_seg, float* out, ctl_stream_t stream) {
  using namespace ctl;
  CTL_CHECK_ARG(x && indptr && out, "null pointer");
  CTL_CHECK_ARG(n > 0 && n_seg > 0 && d > 0 && d % 4 == 0, "bad shape n=%lld n_seg=%lld d=%d (d %% 4 == 0)",
                (long long)n, (long long)n_seg, d);
  int rc = ctl_device_check();
  if (rc) return rc;
  dim3 grid((unsigned)n_seg, (unsigned)((d / 4 + 255) / 256));
  segment_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, d, reinterpret_cast<const long long*>(indptr),
                                                            reinterpret_cast<const long long*>(indices), out);
  CTL_LAUNCH_CHECK();
  return 0;
}
