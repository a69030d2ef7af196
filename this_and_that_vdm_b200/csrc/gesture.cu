// Gesture rasteriser: the [F, 3, H, W] "this / that" condition image of
// data_loader/video_this_that_dataset.py:28-130 (get_thisthat_sam; duplicated in app.py:282-328).
//
// Reference, per gesture point: 255-filled float32 image of the ORIGINAL frame size with a 21 x 21 square around the
// point (first point [0,0,255], later points [0,255,0], BGR), cv2.filter2D with the normalised 99 x 99 isotropic
// Gaussian (sigma 10, BORDER_REFLECT_101), cv2.resize INTER_CUBIC to (W, H), optional left-right flip, / 255, written
// into frame `frame_idx` of a zero tensor — 99 x 99 taps per pixel of the original frame on the CPU.
//
// Here the same function is evaluated in closed form. The image is 255 - (255 - colour) * rowmask(r) * colmask(c), the
// kernel is a product k1(dy) k1(dx), filtering / reflection / bicubic interpolation are all separable and linear, so
//     out(c, y, x) = 1 - (1 - colour_c / 255) * Py(y) * Px(x)
// with Py = cubic_resize_1d(reflect101_filter_1d(rowmask, k1)) (H numbers) and Px likewise (W numbers): one tiny
// kernel builds the 1-D profiles of every point (fp64 sums, cv2's float32 cubic coefficients), a second one writes the
// whole output tensor once at HBM speed (frames without a point are zeros, as in the reference).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kGsMaxPoints = TTVDM_GESTURE_MAX_POINTS;
constexpr int kGsRadius = 10;   // square "diameter" of the reference: [-10, 10]
constexpr int kGsTaps = 49;     // 99-tap Gaussian, sigma 10

struct GsPoints {
  int n;
  int frame[kGsMaxPoints], v[kGsMaxPoints], h[kGsMaxPoints];
};

__device__ __forceinline__ int gs_reflect101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  int m = i % period;
  if (m < 0) m += period;
  return m >= n ? period - m : m;
}

// one CTA per point: scratch[p][0..H) = Py, scratch[p][H..H+W) = Px
__global__ void __launch_bounds__(256)
gesture_profiles_kernel(const GsPoints pts, int org_h, int org_w, int H, int W, int dilate, float* __restrict__ scratch) {
  extern __shared__ double gs_smem[];
  double* R = gs_smem;          // [max(org_h, org_w)] blurred 1-D mask of the current axis
  __shared__ double k1[2 * kGsTaps + 1];
  __shared__ double ksum;
  const int p = blockIdx.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int d = -kGsTaps; d <= kGsTaps; ++d) s += exp(-0.5 * (double)(d * d) / 100.0);
    ksum = s;
  }
  __syncthreads();
  for (int d = threadIdx.x; d <= 2 * kGsTaps; d += blockDim.x)
    k1[d] = exp(-0.5 * (double)((d - kGsTaps) * (d - kGsTaps)) / 100.0) / ksum;
  __syncthreads();
#pragma unroll 1
  for (int axis = 0; axis < 2; ++axis) {
    const int n_src = axis == 0 ? org_h : org_w;
    const int n_dst = axis == 0 ? H : W;
    const int centre = axis == 0 ? pts.v[p] : pts.h[p];
    const int lo = max(centre - kGsRadius, 0), hi = min(centre + kGsRadius, n_src - 1);  // the clipped square
    for (int r = threadIdx.x; r < n_src; r += blockDim.x) {
      double acc;
      if (dilate) {
        acc = 0.0;
        for (int d = -kGsTaps; d <= kGsTaps; ++d) {
          const int q = gs_reflect101(r + d, n_src);
          if (q >= lo && q <= hi) acc += k1[d + kGsTaps];
        }
      } else {
        acc = (r >= lo && r <= hi) ? 1.0 : 0.0;
      }
      R[r] = acc;
    }
    __syncthreads();
    // cv2.resize INTER_CUBIC: half-pixel centres, A = -0.75, float32 coefficients, replicated border
    const double scale = (double)n_src / (double)n_dst;
    float* dst = scratch + (size_t)p * (H + W) + (axis == 0 ? 0 : H);
    for (int y = threadIdx.x; y < n_dst; y += blockDim.x) {
      const float f = (float)((y + 0.5) * scale - 0.5);
      const int s = (int)floorf(f);
      const float x = f - (float)s;
      const float A = -0.75f;
      float c[4];
      c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
      c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
      c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
      c[3] = 1.f - c[0] - c[1] - c[2];
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) acc += (double)c[a] * R[min(max(s - 1 + a, 0), n_src - 1)];
      dst[y] = (float)acc;
    }
    __syncthreads();
  }
}

// out[f][c][y][x]; one thread per 4 consecutive x (W % 4 == 0) or per element
template <int kVec>
__global__ void __launch_bounds__(256)
gesture_fill_kernel(const GsPoints pts, int F, int H, int W, int flip, const float* __restrict__ scratch,
                    float* __restrict__ out) {
  const size_t total = (size_t)F * 3 * H * (W / kVec);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xv = (int)(i % (W / kVec));
    size_t rest = i / (W / kVec);
    const int y = (int)(rest % H);
    rest /= H;
    const int c = (int)(rest % 3);
    const int f = (int)(rest / 3);
    int p = -1;
    for (int q = 0; q < pts.n; ++q)
      if (pts.frame[q] == f) p = q;  // later points overwrite earlier ones on the same frame
    float v[kVec];
#pragma unroll
    for (int k = 0; k < kVec; ++k) v[k] = 0.f;
    if (p >= 0) {
      // BGR colours of the reference: point 0 is [0, 0, 255], the others [0, 255, 0]; a channel at 255 stays 255
      const bool keep = (p == 0) ? (c == 2) : (c == 1);
      const float* Py = scratch + (size_t)p * (H + W);
      const float* Px = Py + H;
      const float py = Py[y];
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const int x = xv * kVec + k;
        v[k] = keep ? 1.0f : 1.0f - py * Px[flip ? W - 1 - x : x];
      }
    }
    float* o = out + (((size_t)f * 3 + c) * H + y) * W + xv * kVec;
    if (kVec == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else o[0] = v[0];
  }
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_gesture_raster(const ttvdm_gesture_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->out || !p->scratch) return fail(TTVDM_ERR_SHAPE, "gesture_raster: null");
  if (p->n_points < 0 || p->n_points > kGsMaxPoints)
    return fail(TTVDM_ERR_SHAPE, "gesture_raster: n_points=%d (0..%d)", p->n_points, kGsMaxPoints);
  if (p->org_h <= 0 || p->org_w <= 0 || p->H <= 0 || p->W <= 0 || p->F <= 0)
    return fail(TTVDM_ERR_SHAPE, "gesture_raster: empty image");
  if (p->org_h > 16384 || p->org_w > 16384) return fail(TTVDM_ERR_SHAPE, "gesture_raster: original frame larger than 16384");
  GsPoints pts;
  pts.n = p->n_points;
  for (int i = 0; i < kGsMaxPoints; ++i) {
    pts.frame[i] = i < p->n_points ? p->frame_idx[i] : -1;
    pts.v[i] = i < p->n_points ? p->vertical[i] : 0;
    pts.h[i] = i < p->n_points ? p->horizontal[i] : 0;
    if (i < p->n_points && (p->frame_idx[i] < 0 || p->frame_idx[i] >= p->F))
      return fail(TTVDM_ERR_SHAPE, "gesture_raster: frame_idx[%d]=%d outside [0, %d)", i, p->frame_idx[i], p->F);
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* scratch = static_cast<float*>(p->scratch);
  if (p->n_points > 0) {
    const int n_max = p->org_h > p->org_w ? p->org_h : p->org_w;
    const size_t smem = (size_t)n_max * sizeof(double);
    static size_t smem_set = 0;
    if (smem > smem_set) {
      cudaError_t e = cudaFuncSetAttribute(gesture_profiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "gesture_raster: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      smem_set = smem;
    }
    gesture_profiles_kernel<<<p->n_points, 256, smem, stream>>>(pts, p->org_h, p->org_w, p->H, p->W, p->dilate ? 1 : 0, scratch);
    TTVDM_CHECK_LAUNCH("gesture_profiles_kernel");
  }
  const bool vec = (p->W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p->out) & 15) == 0);
  const size_t total = (size_t)p->F * 3 * p->H * (vec ? p->W / 4 : p->W);
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)g_num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (vec) gesture_fill_kernel<4><<<(int)blocks, 256, 0, stream>>>(pts, p->F, p->H, p->W, p->flip ? 1 : 0, scratch, static_cast<float*>(p->out));
  else gesture_fill_kernel<1><<<(int)blocks, 256, 0, stream>>>(pts, p->F, p->H, p->W, p->flip ? 1 : 0, scratch, static_cast<float*>(p->out));
  TTVDM_CHECK_LAUNCH("gesture_fill_kernel");
  return 0;
}
