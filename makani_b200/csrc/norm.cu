// Pointwise tail of the SFNO block around the spectral filter (reference: torch.nn.InstanceNorm2d(affine, eps 1e-6) + nn.GELU as
// built at makani/models/networks/sfnonet.py:618-620 and applied at :385-406, bias + GELU of the 1x1-convolution MLP / encoder / decoder,
// makani/models/common/layers.py:537-760).  At 721 x 1440 x 384 one activation is 0.8 GB in bf16: these layers are pure HBM traffic, and PyTorch's
// instance norm (batch-norm kernels with one block per channel) needs 15.6 ms of a 77 ms model step for them.
//
//   instance norm (+ GELU), rows = (b, c), n = H * W contiguous elements per row, every row split over `splits` CTAs:
//     forward : stats   partial (sum, sum of squares) of x - x[row start] per (row, split)      read x
//               apply   y = [gelu]((x - mean) * rstd * gamma[c] + beta[c])                       read x, write y
//     backward: reduce  partial S1 = sum g, S2 = sum g * xhat   (g = dy, or dy * gelu'(z))      read x, dy
//               apply   dx = rstd * gamma[c] * (g - S1 / n - xhat * S2 / n)                      read x, dy, write dx
//     dgamma[c] = sum_b S2, dbeta[c] = sum_b S1 are formed by the caller from the per-row sums (tiny).
//   bias + GELU: y = gelu(x + bias[c]);  dx = dy * gelu'(x + bias[c]), per-(row, split) partial sums of dx for dbias.
// All arithmetic in fp32, activations float or bf16, 16-byte vector accesses when the row length allows it.
#include "common.cuh"

namespace b200sht {

constexpr int kNormThreads = 256;
constexpr int kNormMaxSplits = 64;

__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678118654752440f)) + z * 0.39894228040143267794f * __expf(-0.5f * z * z);
}

template <typename T> __device__ __forceinline__ float ldf(const T* p, long long i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, long long i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, long long i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

// 16-byte packets: 4 floats or 8 bf16
template <typename T> struct Pack;
template <> struct Pack<float> {
  static constexpr int kN = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return i == 0 ? raw.x : i == 1 ? raw.y : i == 2 ? raw.z : raw.w; }
  __device__ __forceinline__ void set(int i, float v) { if (i == 0) raw.x = v; else if (i == 1) raw.y = v; else if (i == 2) raw.z = v; else raw.w = v; }
};
template <> struct Pack<__nv_bfloat16> {
  static constexpr int kN = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = (i >> 1) == 0 ? raw.x : (i >> 1) == 1 ? raw.y : (i >> 1) == 2 ? raw.z : raw.w;
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(int i, float v) {
    const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    uint32_t* w = (i >> 1) == 0 ? &raw.x : (i >> 1) == 1 ? &raw.y : (i >> 1) == 2 ? &raw.z : &raw.w;
    *w = (i & 1) ? ((*w & 0x0000ffffu) | (h << 16)) : ((*w & 0xffff0000u) | h);
  }
};

struct NormArgs {
  const void* x;
  const void* dy;
  void* out;           // y (forward) / dx (backward)
  const float* gamma;  // [C] or null (1)
  const float* beta;   // [C] or null (0): instance-norm shift, or the bias of bias + GELU
  const float* stats;  // [rows][2] mean, rstd
  const float* sums;   // [rows][2] S1, S2 (backward apply)
  float* partial;      // [rows][splits][2]
  long long n;         // elements per row
  long long chunk;     // elements per split (a multiple of 8)
  int rows, C, splits, gelu;
};

// two block-wide sums (blockDim.x == kNormThreads); result valid in thread 0
__device__ __forceinline__ void block_sum2(float& a, float& b) {
  __shared__ float sa[kNormThreads / 32], sb[kNormThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < kNormThreads / 32 ? sa[l] : 0.f;
    b = l < kNormThreads / 32 ? sb[l] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  }
}

// What a kernel does with one element.  MODE 0: forward statistics, 1: forward apply, 2: backward reduce, 3: backward apply,
// 4: bias + GELU forward, 5: bias + GELU backward (also accumulates sum dx)
template <typename T, int MODE, bool VEC>
__global__ void __launch_bounds__(kNormThreads) norm_kernel(const NormArgs a) {
  const int r = blockIdx.y, s = blockIdx.x;
  const long long beg = (long long)s * a.chunk, end = beg + a.chunk < a.n ? beg + a.chunk : a.n;
  const T* x = static_cast<const T*>(a.x) + (size_t)r * a.n;
  const T* dy = static_cast<const T*>(a.dy) + (size_t)r * a.n;
  T* out = static_cast<T*>(a.out) + (size_t)r * a.n;
  const int c = r % a.C;
  const float gamma = a.gamma ? a.gamma[c] : 1.f, beta = a.beta ? a.beta[c] : 0.f;
  float mean = 0.f, rstd = 1.f, pivot = 0.f, m1 = 0.f, m2 = 0.f;
  if (MODE == 0) pivot = ldf<T>(x, 0);
  if (MODE == 1 || MODE == 2 || MODE == 3) { mean = a.stats[2 * r]; rstd = a.stats[2 * r + 1]; }
  if (MODE == 3) { m1 = a.sums[2 * r] / (float)a.n; m2 = a.sums[2 * r + 1] / (float)a.n; }
  const float gs = gamma * rstd;
  float acc0 = 0.f, acc1 = 0.f;

  auto element = [&](float xv, float dv, float& ov) {
    if (MODE == 0) {
      const float d = xv - pivot;
      acc0 += d; acc1 = fmaf(d, d, acc1);
    } else if (MODE == 1) {
      const float z = fmaf((xv - mean) * rstd, gamma, beta);
      ov = a.gelu ? gelu_f(z) : z;
    } else if (MODE == 2 || MODE == 3) {
      const float xh = (xv - mean) * rstd;
      const float g = a.gelu ? dv * gelu_grad_f(fmaf(xh, gamma, beta)) : dv;
      if (MODE == 2) { acc0 += g; acc1 = fmaf(g, xh, acc1); }
      else ov = gs * (g - m1 - xh * m2);
    } else if (MODE == 4) {
      ov = gelu_f(xv + beta);
    } else {
      ov = dv * gelu_grad_f(xv + beta);
      acc0 += ov;
    }
  };
  constexpr bool kNeedDy = (MODE == 2 || MODE == 3 || MODE == 5), kWrites = (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 5);
  if (VEC) {
    constexpr int kN = Pack<T>::kN;
    constexpr long long kStride = (long long)kNormThreads * kN;
    long long i = beg + (long long)threadIdx.x * kN;   // beg, n multiples of kN: whole packets
    // two packets per tensor in flight per thread (first measurement of the one-packet loop: 45-50 % of HBM, latency bound at 5 CTAs per SM)
    for (; i + kStride < end; i += 2 * kStride) {
      Pack<T> px0, px1, pd0, pd1, po0, po1;
      px0.load(x + i);
      px1.load(x + i + kStride);
      if (kNeedDy) { pd0.load(dy + i); pd1.load(dy + i + kStride); }
#pragma unroll
      for (int j = 0; j < kN; ++j) {
        float ov = 0.f;
        element(px0.get(j), kNeedDy ? pd0.get(j) : 0.f, ov);
        if (kWrites) po0.set(j, ov);
      }
#pragma unroll
      for (int j = 0; j < kN; ++j) {
        float ov = 0.f;
        element(px1.get(j), kNeedDy ? pd1.get(j) : 0.f, ov);
        if (kWrites) po1.set(j, ov);
      }
      if (kWrites) { po0.store(out + i); po1.store(out + i + kStride); }
    }
    for (; i < end; i += kStride) {
      Pack<T> px, pd, po;
      px.load(x + i);
      if (kNeedDy) pd.load(dy + i);
#pragma unroll
      for (int j = 0; j < kN; ++j) {
        float ov = 0.f;
        element(px.get(j), kNeedDy ? pd.get(j) : 0.f, ov);
        if (kWrites) po.set(j, ov);
      }
      if (kWrites) po.store(out + i);
    }
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += kNormThreads) {
      float ov = 0.f;
      element(ldf<T>(x, i), kNeedDy ? ldf<T>(dy, i) : 0.f, ov);
      if (kWrites) stf<T>(out, i, ov);
    }
  }
  if (MODE == 0 || MODE == 2 || MODE == 5) {
    block_sum2(acc0, acc1);
    if (threadIdx.x == 0) {
      float* p = a.partial + ((size_t)r * a.splits + s) * 2;
      p[0] = acc0; p[1] = acc1;
    }
  }
}

// per row: combine the split partials.  what 0: (sum d, sum d^2) about the pivot -> (mean, rstd); 1: plain sums (S1, S2)
template <typename T>
__global__ void norm_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, const void* x, int rows, int splits, long long n, float eps, int what) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < splits; ++s) { s0 += partial[((size_t)r * splits + s) * 2]; s1 += partial[((size_t)r * splits + s) * 2 + 1]; }
  if (what == 0) {
    const double pivot = (double)ldf<T>(static_cast<const T*>(x) + (size_t)r * n, 0);
    const double md = s0 / (double)n;
    double var = s1 / (double)n - md * md;
    if (var < 0.0) var = 0.0;
    out[2 * r] = (float)(pivot + md);
    out[2 * r + 1] = (float)(1.0 / sqrt(var + (double)eps));
  } else {
    out[2 * r] = (float)s0;
    out[2 * r + 1] = (float)s1;
  }
}

int norm_splits(int rows, long long n) {
  long long s = (8LL * 148 + rows - 1) / rows;          // >= 8 CTAs per SM's worth of blocks
  const long long by_len = n / 2048 > 0 ? n / 2048 : 1;  // but not less than 2048 elements per CTA
  if (s > by_len) s = by_len;
  if (s > kNormMaxSplits) s = kNormMaxSplits;
  return s < 1 ? 1 : (int)s;
}

template <typename T, int MODE>
static int launch_mode(const NormArgs& a, cudaStream_t st) {
  const dim3 grid(a.splits, a.rows);
  const bool vec = (a.n % Pack<T>::kN == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) && (a.dy == nullptr || (reinterpret_cast<uintptr_t>(a.dy) & 15) == 0) &&
                   (a.out == nullptr || (reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
  if (vec) norm_kernel<T, MODE, true><<<grid, kNormThreads, 0, st>>>(a);
  else norm_kernel<T, MODE, false><<<grid, kNormThreads, 0, st>>>(a);
  B200_CHECK_LAUNCH();
  return 0;
}
template <int MODE>
static int launch_dtype(int dtype, const NormArgs& a, cudaStream_t st) {
  if (dtype == B200SHT_BF16) return launch_mode<__nv_bfloat16, MODE>(a, st);
  return launch_mode<float, MODE>(a, st);
}
static int finalize(int dtype, const float* partial, float* out, const void* x, int rows, int splits, long long n, float eps, int what, cudaStream_t st) {
  const int blocks = (rows + 127) / 128;
  if (dtype == B200SHT_BF16) norm_finalize_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(partial, out, x, rows, splits, n, eps, what);
  else norm_finalize_kernel<float><<<blocks, 128, 0, st>>>(partial, out, x, rows, splits, n, eps, what);
  B200_CHECK_LAUNCH();
  return 0;
}

static int fill_args(NormArgs* a, int B, int C, long long hw) {
  B200_REQUIRE(B > 0 && C > 0 && hw > 0 && (long long)B * C <= 65535, "norm: bad shape (B %d, C %d, H*W %lld; B*C must be <= 65535)", B, C, hw);
  memset(a, 0, sizeof(*a));
  a->rows = B * C; a->C = C; a->n = hw;
  a->splits = norm_splits(a->rows, hw);
  long long chunk = (hw + a->splits - 1) / a->splits;
  a->chunk = (chunk + 7) / 8 * 8;
  return 0;
}

int instance_norm_forward(const void* x, void* y, const float* gamma, const float* beta, float* stats, float* ws, int dtype, int B, int C, long long hw, float eps,
                          int gelu, cudaStream_t st) {
  NormArgs a;
  int rc = fill_args(&a, B, C, hw);
  if (rc) return rc;
  a.x = x; a.partial = ws;
  rc = launch_dtype<0>(dtype, a, st);
  if (!rc) rc = finalize(dtype, ws, stats, x, a.rows, a.splits, hw, eps, 0, st);
  a.out = y; a.gamma = gamma; a.beta = beta; a.stats = stats; a.gelu = gelu;
  if (!rc) rc = launch_dtype<1>(dtype, a, st);
  return rc;
}

int instance_norm_backward(const void* x, const void* dy, void* dx, const float* gamma, const float* beta, const float* stats, float* sums, float* ws, int dtype,
                           int B, int C, long long hw, int gelu, cudaStream_t st) {
  NormArgs a;
  int rc = fill_args(&a, B, C, hw);
  if (rc) return rc;
  a.x = x; a.dy = dy; a.gamma = gamma; a.beta = beta; a.stats = stats; a.partial = ws; a.gelu = gelu;
  rc = launch_dtype<2>(dtype, a, st);
  if (!rc) rc = finalize(dtype, ws, sums, x, a.rows, a.splits, hw, 0.f, 1, st);
  a.out = dx; a.sums = sums;
  if (!rc) rc = launch_dtype<3>(dtype, a, st);
  return rc;
}

int bias_gelu_forward(const void* x, const float* bias, void* y, int dtype, int B, int C, long long hw, cudaStream_t st) {
  NormArgs a;
  int rc = fill_args(&a, B, C, hw);
  if (rc) return rc;
  a.x = x; a.out = y; a.beta = bias;
  return launch_dtype<4>(dtype, a, st);
}

int bias_gelu_backward(const void* x, const float* bias, const void* dy, void* dx, float* row_sums, float* ws, int dtype, int B, int C, long long hw, cudaStream_t st) {
  NormArgs a;
  int rc = fill_args(&a, B, C, hw);
  if (rc) return rc;
  a.x = x; a.dy = dy; a.out = dx; a.beta = bias; a.partial = ws;
  rc = launch_dtype<5>(dtype, a, st);
  if (!rc && row_sums) rc = finalize(dtype, ws, row_sums, x, a.rows, a.splits, hw, 0.f, 1, st);
  return rc;
}

}  // namespace b200sht
