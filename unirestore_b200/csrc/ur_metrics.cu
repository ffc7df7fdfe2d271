// ur_image_metrics: PSNR / SSIM partial sums on the GPU right after the hot path.
//
// The reference's validate loop copies every restored image to the host and calls skimage per image
// (src/core/base/eval_image_restoration.py:71,255-313: 8-bit quantisation, SKPSNR, SKSSIM with win_size 7, uniform
// window, sample covariance, channel_axis 0).  Here one kernel per batch produces, per image,
//   out[b][0] = sum over all (c, y, x) of (t - p)^2
//   out[b][1] = sum over channels and interior pixels (3 px border cropped) of the SSIM map S
// in fp64; the host turns them into PSNR = 10 log10(1 / mse) and SSIM = sum / (C (H-6) (W-6)).
#include "ur_common.cuh"
#include "ur_host.h"

namespace ur {

__device__ __forceinline__ float quant8(float v) { return fminf(fmaxf(rintf(v * 255.0f), 0.0f), 255.0f) / 255.0f; }

__global__ void image_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, int C, int H,
                                     int W, int quantize, double data_range, double* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, c = blockIdx.y;
  const long long plane = static_cast<long long>(H) * W;
  const float* P = pred + (static_cast<long long>(b) * C + c) * plane;
  const float* T = target + (static_cast<long long>(b) * C + c) * plane;
  const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
  double se = 0.0, ss = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W), y = static_cast<int>(i / W);
    const float p0 = quantize ? quant8(P[i]) : P[i];
    const double d = static_cast<double>(T[i]) - static_cast<double>(p0);
    se += d * d;
    if (x >= 3 && x < W - 3 && y >= 3 && y < H - 3) {
      double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
      for (int dy = -3; dy <= 3; ++dy) {
        const float* pr = P + static_cast<long long>(y + dy) * W + x;
        const float* tr = T + static_cast<long long>(y + dy) * W + x;
#pragma unroll
        for (int dx = -3; dx <= 3; ++dx) {
          const double a = quantize ? quant8(pr[dx]) : pr[dx];
          const double t = tr[dx];
          sx += a;
          sy += t;
          sxx += a * a;
          syy += t * t;
          sxy += a * t;
        }
      }
      const double n = 49.0, cov = 49.0 / 48.0;
      const double ux = sx / n, uy = sy / n;
      const double vx = cov * (sxx / n - ux * ux), vy = cov * (syy / n - uy * uy), vxy = cov * (sxy / n - ux * uy);
      ss += ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 2 * b, se);
    atomicAdd(out + 2 * b + 1, ss);
  }
}

}  // namespace ur

using namespace ur;

extern "C" int ur_image_metrics(const float* pred, const float* target, int batch, int channels, int h, int w,
                                int quantize_pred, float data_range, double* out, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!pred || !target || !out || batch <= 0 || channels <= 0 || h < 7 || w < 7)
    return set_error(UR_ERR_ARG, "ur_image_metrics: bad arguments (images must be at least 7x7)");
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 2 * batch, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "ur_image_metrics memset");
  const long long plane = static_cast<long long>(h) * w;
  int gx = static_cast<int>((plane + 255) / 256);
  const int cap = (8 * num_sms()) / (batch * channels) + 1;
  if (gx > cap) gx = cap;
  e = launch_kernel(image_metrics_kernel, dim3(gx, channels, batch), dim3(256), 0, stream, pred, target, channels, h, w,
                    quantize_pred, static_cast<double>(data_range), out);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "ur_image_metrics launch");
}
