// Frequency-domain input stage of the reference: what Augmenter.fft_preprocess
// (src/data_augmenter/Augmenter.py:141-158) does after torch.fft.fft -- view_as_real, permute(0,1,4,2,3), reshape to
// [b, 2c, i, s] -- fused with PhaseShiftAugmenter's rotation (src/data_augmenter/PhaseShiftAugmenter.py:35-57): one pass
// over the spectrum instead of seven elementwise ATen passes (clone, abs, angle, cos, sin, 2 mul, stack, permute).
// HBM-bound: 8 bytes read + 8 bytes written per complex bin, 16-byte vectors on both sides.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fb {

// in:  interleaved == 1: complex64 [bc][plane] as (re, im) pairs (cuFFT output);  0: planar fp32 [bc][2][plane]
// out: planar fp32 [bc][2][plane]  (channel 2c = real part, 2c + 1 = imaginary part)
// (re', im') = (re cos - im sin, re sin + im cos); cos = 1, sin = 0 is the plain layout change.
__global__ void __launch_bounds__(256) spectrum_rotate_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              long long n_bc, int plane, int interleaved, float cs,
                                                              float sn) {
  const int vec_per_plane = plane >> 2;                           // plane % 4 == 0 (checked by the host)
  const long long total = n_bc * vec_per_plane;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const long long bc = v / vec_per_plane;
    const int p4 = (int)(v - bc * vec_per_plane) * 4;
    float re[4], im[4];
    if (interleaved) {
      const float4* src = reinterpret_cast<const float4*>(in + (bc * plane + p4) * 2);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      re[0] = a.x; im[0] = a.y; re[1] = a.z; im[1] = a.w; re[2] = b.x; im[2] = b.y; re[3] = b.z; im[3] = b.w;
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(in + (bc * 2) * plane + p4));
      const float4 b = __ldg(reinterpret_cast<const float4*>(in + (bc * 2 + 1) * plane + p4));
      re[0] = a.x; re[1] = a.y; re[2] = a.z; re[3] = a.w; im[0] = b.x; im[1] = b.y; im[2] = b.z; im[3] = b.w;
    }
    float4 o_re, o_im;
    o_re.x = fmaf(re[0], cs, -im[0] * sn); o_im.x = fmaf(re[0], sn, im[0] * cs);
    o_re.y = fmaf(re[1], cs, -im[1] * sn); o_im.y = fmaf(re[1], sn, im[1] * cs);
    o_re.z = fmaf(re[2], cs, -im[2] * sn); o_im.z = fmaf(re[2], sn, im[2] * cs);
    o_re.w = fmaf(re[3], cs, -im[3] * sn); o_im.w = fmaf(re[3], sn, im[3] * cs);
    *reinterpret_cast<float4*>(out + (bc * 2) * plane + p4) = o_re;
    *reinterpret_cast<float4*>(out + (bc * 2 + 1) * plane + p4) = o_im;
  }
}

}  // namespace fb
