// Throughput of the softmax inner loop variants on one B200 SM (8 warps per CTA = 2 per SMSP, like attn_flash's softmax
// warps; one CTA per SM). Reports cycles per 128-element row-slice per warp and elements/clk/SM for:
//   0  f32 path            : FFMA(scale) + MUFU.EX2(f32) + FADD(sum) + F2FP.BF16 pack per pair   (the round-1 kernel)
//   1  f32 path, f32x2     : FFMA2(scale pair) + 2 MUFU.EX2 + FADD2(sum pair) + pack
//   2  f16x2 path          : FFMA2 + cvt.f16x2 + ex2.approx.f16x2 + (cvt f16->f32 x2 + FADD2 sum) + pack bf16x2
//   3  bf16x2 path         : FFMA2 + cvt.bf16x2 + ex2.approx.ftz.bf16x2 + (unpack + FADD2)
//   4  polynomial f32x2    : Cody-Waite + degree-3 Horner in fma.rn.f32x2, exponent splice on the ALU pipe
//   5  raw MUFU.EX2 f32    6  raw ex2.f16x2    7  raw ex2.bf16x2     (pipe rates, nothing else in the loop)
//   8  mix: 5/8 MUFU f32 + 3/8 polynomial (f32x2 everywhere)
// Also checks accuracy of variants 2 and 4 against exp2f.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t cvt_bf2(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float h_lo(uint32_t v) { float f; asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, l;}" : "=f"(f) : "r"(v)); return f; }
__device__ __forceinline__ float h_hi(uint32_t v) { float f; asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, h;}" : "=f"(f) : "r"(v)); return f; }

// exp2 of a pair on the FMA/ALU pipes: x <= 0 (clamped to >= -126). n = round(x) via the 1.5*2^23 magic constant,
// f = x - n in [-0.5, 0.5], p(f) ~ 2^f (degree-3 minimax, rel err ~1e-4 — below bf16 rounding of P), result = p * 2^n by adding n
// to the exponent field.
__device__ __forceinline__ void poly_exp2_pair(uint64_t x2, float& r0, float& r1) {
  const uint64_t magic = pk(12582912.f, 12582912.f);
  const uint64_t one = pk(1.f, 1.f), mone = pk(-1.f, -1.f);
  uint64_t t = add2(x2, magic);                   // low mantissa bits of t = round(x)
  uint64_t n = add2(t, pk(-12582912.f, -12582912.f));
  uint64_t f = fma2(n, mone, x2);                 // x - n
  uint64_t p = fma2(f, pk(0.05550411f, 0.05550411f), pk(0.24022651f, 0.24022651f));
  p = fma2(p, f, pk(0.69314718f, 0.69314718f));
  p = fma2(p, f, one);
  float p0, p1, t0, t1;
  upk(p, p0, p1);
  upk(t, t0, t1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench(const float* __restrict__ in, uint32_t* __restrict__ out, float* __restrict__ sums,
                                               long long* __restrict__ cyc, int iters, float c2, float ms) {
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = in[(threadIdx.x * 128 + i) & 4095];
  float s0 = 0.f, s1 = 0.f;
  uint64_t s2 = pk(0.f, 0.f);
  uint32_t acc = 0;
  const uint64_t c22 = pk(c2, c2), ms2 = pk(-ms, -ms);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 128; i += 2) {
      if (MODE == 0) {
        const float p0 = ex2f(fmaf(v[i], c2, -ms)), p1 = ex2f(fmaf(v[i + 1], c2, -ms));
        s0 += p0; s1 += p1;
        acc ^= cvt_bf2(p0, p1);
      } else if (MODE == 1) {
        float x0, x1;
        upk(fma2(pk(v[i], v[i + 1]), c22, ms2), x0, x1);
        const float p0 = ex2f(x0), p1 = ex2f(x1);
        s2 = add2(s2, pk(p0, p1));
        acc ^= cvt_bf2(p0, p1);
      } else if (MODE == 2) {
        float x0, x1;
        upk(fma2(pk(v[i], v[i + 1]), c22, ms2), x0, x1);
        const uint32_t e = ex2_h2(cvt_h2(x0, x1));
        const float p0 = h_lo(e), p1 = h_hi(e);
        s2 = add2(s2, pk(p0, p1));
        acc ^= cvt_bf2(p0, p1);
      } else if (MODE == 3) {
        float x0, x1;
        upk(fma2(pk(v[i], v[i + 1]), c22, ms2), x0, x1);
        const uint32_t e = ex2_bf2(cvt_bf2(x0, x1));
        const float p0 = __uint_as_float(e << 16), p1 = __uint_as_float(e & 0xffff0000u);
        s2 = add2(s2, pk(p0, p1));
        acc ^= e;
      } else if (MODE == 4) {
        float p0, p1;
        poly_exp2_pair(fma2(pk(v[i], v[i + 1]), c22, ms2), p0, p1);
        s2 = add2(s2, pk(p0, p1));
        acc ^= cvt_bf2(p0, p1);
      } else if (MODE == 5) {
        s0 += ex2f(v[i]); s1 += ex2f(v[i + 1]);
      } else if (MODE == 6) {
        acc ^= ex2_h2(__float_as_uint(v[i]) ^ acc);
        acc ^= ex2_h2(__float_as_uint(v[i + 1]));
      } else if (MODE == 7) {
        acc ^= ex2_bf2(__float_as_uint(v[i]));
        acc ^= ex2_bf2(__float_as_uint(v[i + 1]));
      } else if (MODE == 8) {
        float p0, p1;
        const uint64_t x2 = fma2(pk(v[i], v[i + 1]), c22, ms2);
        if ((i & 15) < 10) { float x0, x1; upk(x2, x0, x1); p0 = ex2f(x0); p1 = ex2f(x1); }
        else poly_exp2_pair(x2, p0, p1);
        s2 = add2(s2, pk(p0, p1));
        acc ^= cvt_bf2(p0, p1);
      }
    }
    // keep the inputs moving so nothing is hoisted
#pragma unroll
    for (int i = 0; i < 128; i += 32) v[i] = __uint_as_float(__float_as_uint(v[i]) ^ (acc & 1));
  }
  const long long t1 = clock64();
  float a, b;
  upk(s2, a, b);
  sums[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + a + b;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void accuracy(float* out) {
  // max relative error of the f16x2 and polynomial paths over x in [-20, 0]
  float e_h = 0.f, e_p = 0.f, e_b = 0.f;
  for (int i = threadIdx.x; i < 200000; i += blockDim.x) {
    const float x = -20.f * i / 200000.f;
    const float ref = exp2f(x);
    const uint32_t e = ex2_h2(cvt_h2(x, x));
    float p0, p1;
    poly_exp2_pair(pk(x, x), p0, p1);
    const uint32_t eb = ex2_bf2(cvt_bf2(x, x));
    if (ref > 1e-4f) {
      e_h = fmaxf(e_h, fabsf(h_lo(e) - ref) / ref);
      e_b = fmaxf(e_b, fabsf(__uint_as_float(eb << 16) - ref) / ref);
    }
    e_p = fmaxf(e_p, fabsf(p0 - ref) / ref);
  }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(e_h));
  atomicMax(reinterpret_cast<int*>(out) + 1, __float_as_int(e_p));
  atomicMax(reinterpret_cast<int*>(out) + 2, __float_as_int(e_b));
}

template <int MODE>
static void run(const char* name, const float* in, uint32_t* out, float* sums, long long* cyc, int n_sm) {
  const int iters = 200;
  bench<MODE><<<n_sm, 256>>>(in, out, sums, cyc, 10, 0.18f, 1.0f);
  bench<MODE><<<n_sm, 256>>>(in, out, sums, cyc, iters, 0.18f, 1.0f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s CUDA error %s\n", name, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, cyc, n_sm * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < n_sm; ++i) avg += (double)h[i];
  avg /= n_sm;
  const double per_row = avg / iters;  // cycles for 8 warps x 128 elements each (per thread)
  printf("%-28s %8.1f cycles per 128-element row (8 warps/SM)  -> %6.2f elements/clk/SM\n", name, per_row, 256.0 * 128.0 / per_row);
}

int main() {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  float* in; uint32_t* out; float* sums; long long* cyc; float* acc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, n_sm * 256 * 4); cudaMalloc(&sums, n_sm * 256 * 4); cudaMalloc(&cyc, 256 * 8);
  cudaMalloc(&acc, 16); cudaMemset(acc, 0, 16);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -(float)(i % 97) * 0.37f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("0 f32 (round-1 loop)", in, out, sums, cyc, n_sm);
  run<1>("1 f32 + f32x2 scale/sum", in, out, sums, cyc, n_sm);
  run<2>("2 f16x2 ex2", in, out, sums, cyc, n_sm);
  run<3>("3 bf16x2 ex2", in, out, sums, cyc, n_sm);
  run<4>("4 polynomial f32x2", in, out, sums, cyc, n_sm);
  run<5>("5 raw MUFU.EX2 f32", in, out, sums, cyc, n_sm);
  run<6>("6 raw ex2.f16x2 (per 2 elem)", in, out, sums, cyc, n_sm);
  run<7>("7 raw ex2.bf16x2 (per 2 elem)", in, out, sums, cyc, n_sm);
  run<8>("8 mix 5/8 MUFU + 3/8 poly", in, out, sums, cyc, n_sm);
  accuracy<<<1, 256>>>(acc);
  float ha[4];
  cudaMemcpy(ha, acc, 16, cudaMemcpyDeviceToHost);
  printf("max rel err vs exp2f on [-20,0]: f16x2 (p>1e-4) %.3e   polynomial %.3e   bf16x2 (p>1e-4) %.3e\n", ha[0], ha[1], ha[2]);
  return 0;
}
