// Associated-Legendre stage of the SHT: table precompute, fp32 CUDA-core contractions (strict mode) and the
// packed <-> torch layout converters.
//
// Replaces, in torch-harmonics (third-party, pinned 887006c6..., not vendored in /root/reference):
//   legendre precompute (`_precompute_legpoly`)  -> build_table_kernel   (SURVEY.md Appendix A recurrence)
//   RealSHT.forward einsum  "...km,mlk->...lm"   -> legendre_analysis_simt_kernel
//   InverseRealSHT.forward  "...lm,mlk->...km"   -> legendre_synthesis_simt_kernel
// Reference call sites: /root/reference/makani/models/common/spectral_convolution.py:239-253.
#include "common.cuh"
#include <cmath>
#include <vector>

namespace b200sht {

// ------------------------------------------------------------------------------------------ table build
// Orthonormal associated Legendre functions for fixed order m at x = cos(theta): writes P[m][l] for l < lmax
// with stride `stride` between degrees.  fp64 recurrence, fp32 storage.
HD void legendre_column(int m, int lmax, double x, int csphase, float* out, size_t stride) {
  const double sgn = (csphase && (m & 1)) ? -1.0 : 1.0;
  double pmm = 0.28209479177387814347;  // 1/sqrt(4 pi)
  const double s2 = (1.0 + x) * (1.0 - x);
  for (int l = 1; l <= m; ++l) pmm *= sqrt((2.0 * l + 1.0) * s2 / (2.0 * l));
  for (int l = 0; l < m && l < lmax; ++l) out[(size_t)l * stride] = 0.f;
  if (m >= lmax) return;
  out[(size_t)m * stride] = (float)(sgn * pmm);
  if (m + 1 >= lmax) return;
  double p2 = pmm;                                   // P[m][l-2]
  double p1 = sqrt(2.0 * (m + 1) + 1.0) * x * pmm;   // P[m][l-1]
  out[(size_t)(m + 1) * stride] = (float)(sgn * p1);
  for (int l = m + 2; l < lmax; ++l) {
    const double a = sqrt((2.0 * l - 1.0) / (double)(l - m) * (2.0 * l + 1.0) / (double)(l + m));
    const double b = sqrt((double)(l + m - 1) / (double)(l - m) * (2.0 * l + 1.0) / (2.0 * l - 3.0) *
                          (double)(l - m - 1) / (double)(l + m));
    const double p0 = x * a * p1 - b * p2;
    out[(size_t)l * stride] = (float)(sgn * p0);
    p2 = p1;
    p1 = p0;
  }
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = tf32_rn(src[i]);
}

// 3 x TF32: an fp32 operand v enters a kind::tf32 MMA as trunc(v) (the tensor core ignores the 13 low mantissa bits); the residual
// v - trunc(v) is exact in fp32 and is fed to a second MMA, rounded to nearest TF32 here (its own truncation would add a 2^-21 bias).
__global__ void tf32_residual_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = src[i];
  float4 r;
  r.x = tf32_rn(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
  r.y = tf32_rn(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
  r.z = tf32_rn(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
  r.w = tf32_rn(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
  dst[i] = r;
}
// n floats, n % 4 == 0, both pointers 16-byte aligned
int tf32_residual(const float* src, float* dst, size_t n, cudaStream_t st) {
  const size_t n4 = n / 4;
  if (n4 == 0) return 0;
  tf32_residual_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), n4);
  B200_CHECK_LAUNCH();
  return 0;
}
__global__ void table_lo_kernel(const float* __restrict__ full, const float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lo[i] = tf32_rn(full[i] - hi[i]);
}
int table_residual(const float* full, const float* hi, float* lo, size_t n, cudaStream_t st) {
  table_lo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(full, hi, lo, n);
  B200_CHECK_LAUNCH();
  return 0;
}

int round_table_tf32(const float* src, float* dst, size_t n, cudaStream_t st) {
  round_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n);
  B200_CHECK_LAUNCH();
  return 0;
}

__global__ void build_table_kernel(float* __restrict__ table, const double* __restrict__ cost, int nlat, int kp, int lmax,
                                   int mmax, int csphase, int m0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (k >= kp) return;
  float* out = table + (size_t)m * lmax * kp + k;
  if (k >= nlat) {
    for (int l = 0; l < lmax; ++l) out[(size_t)l * kp] = 0.f;
    return;
  }
  legendre_column(m0 + m, lmax, cost[k], csphase, out, kp);
}

int build_table(Plan* pl, const double* d_cost, cudaStream_t st) {
  dim3 grid(ceil_div(pl->kp, 128), pl->mmax);
  build_table_kernel<<<grid, 128, 0, st>>>(pl->d_table, d_cost, pl->nlat, pl->kp, pl->lmax, pl->mmax, pl->csphase, pl->m0);
  B200_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------- packed index helpers
// jp: flattened padded (p, b, cp) index of a packed spec row;  returns the flattened (p, b, c) latspec row or -1.
__device__ __forceinline__ int jp_to_j(int jp, int B, int C, int cp) {
  const int c = jp % cp;
  const int pb = jp / cp;
  return (c < C) ? pb * C + c : -1;
}

// --------------------------------------------------------------------------------- SIMT contractions
constexpr int TS = 64;   // tile edge
constexpr int TK = 16;   // k slab

// spec[l][m][jp] = sum_k P[m][l][k] * X[m][j][k]      grid: (jp tiles, l tiles, m)
__global__ void __launch_bounds__(256) legendre_analysis_simt_kernel(const float* __restrict__ P, const float* __restrict__ X,
                                                                     float* __restrict__ spec, int L, int M, int kp, int B,
                                                                     int C, int cp, int m0) {
  __shared__ float As[TK][TS + 4];
  __shared__ float Bs[TK][TS + 4];
  const int m = blockIdx.z;
  const int l0 = lstart(m0 + m) + blockIdx.y * TS;
  if (l0 >= L) return;
  const int JP = 2 * B * cp, J = 2 * B * C;
  const int jp0 = blockIdx.x * TS;
  const int t = threadIdx.x;
  const int lrow = t >> 2, kq = (t & 3) * 4;
  const int ty = t >> 4, tx = t & 15;

  const int la = l0 + lrow;
  const float* arow = (la < L) ? P + ((size_t)m * L + la) * kp : nullptr;
  const int jb = (jp0 + lrow < JP) ? jp_to_j(jp0 + lrow, B, C, cp) : -1;
  const float* brow = (jb >= 0) ? X + ((size_t)m * J + jb) * kp : nullptr;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < kp; k0 += TK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (k0 + kq < kp) {  // kp is a multiple of 8, kq of 4: a float4 is entirely in or out
      if (arow) a = __ldg(reinterpret_cast<const float4*>(arow + k0 + kq));
      if (brow) b = __ldg(reinterpret_cast<const float4*>(brow + k0 + kq));
    }
    __syncthreads();
    As[kq + 0][lrow] = a.x; As[kq + 1][lrow] = a.y; As[kq + 2][lrow] = a.z; As[kq + 3][lrow] = a.w;
    Bs[kq + 0][lrow] = b.x; Bs[kq + 1][lrow] = b.y; Bs[kq + 2][lrow] = b.z; Bs[kq + 3][lrow] = b.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = l0 + ty * 4 + i;
    if (l >= L) continue;
    float* orow = spec + ((size_t)l * M + m) * JP;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jp = jp0 + tx * 4 + j;
      if (jp < JP) orow[jp] = acc[i][j];
    }
  }
}

// Z[m][j][k] = sum_{l >= lstart(m)} P[m][l][k] * spec[l][m][jp]     grid: (jp tiles, k tiles, m)
__global__ void __launch_bounds__(256) legendre_synthesis_simt_kernel(const float* __restrict__ P, const float* __restrict__ spec,
                                                                      float* __restrict__ Z, int L, int M, int kp, int B, int C,
                                                                      int cp, int m0) {
  __shared__ float As[TK][TS + 4];  // [l][k]
  __shared__ float Bs[TK][TS + 4];  // [l][jp]
  const int m = blockIdx.z;
  const int JP = 2 * B * cp, J = 2 * B * C;
  const int k0 = blockIdx.y * TS;
  const int jp0 = blockIdx.x * TS;
  const int t = threadIdx.x;
  const int lrow = t >> 4, cq = (t & 15) * 4;  // 16 l-rows x 16 float4 columns
  const int ty = t >> 4, tx = t & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int l0 = lstart(m0 + m); l0 < L; l0 += TK) {
    const int l = l0 + lrow;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (l < L) {
      if (k0 + cq < kp) a = __ldg(reinterpret_cast<const float4*>(P + ((size_t)m * L + l) * kp + k0 + cq));
      if (jp0 + cq < JP) b = __ldg(reinterpret_cast<const float4*>(spec + ((size_t)l * M + m) * JP + jp0 + cq));  // JP % 8 == 0
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lrow][cq]) = a;
    *reinterpret_cast<float4*>(&Bs[lrow][cq]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  // rows i -> k (contiguous in Z), cols j -> jp
  const int k = k0 + ty * 4;
  if (k >= kp) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int jp = jp0 + tx * 4 + j;
    if (jp >= JP) continue;
    const int jj = jp_to_j(jp, B, C, cp);
    if (jj < 0) continue;
    *reinterpret_cast<float4*>(Z + ((size_t)m * J + jj) * kp + k) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
  }
}

int legendre_analysis_simt(const Plan* pl, const float* X, float* spec, int B, int C, cudaStream_t st) {
  const int cp = round_up(C, 4);
  const int JP = 2 * B * cp;
  dim3 grid(ceil_div(JP, TS), ceil_div(pl->lmax, TS), pl->mmax);
  legendre_analysis_simt_kernel<<<grid, 256, 0, st>>>(pl->d_table, X, spec, pl->lmax, pl->mmax, pl->kp, B, C, cp, pl->m0);
  B200_CHECK_LAUNCH();
  return 0;
}

int legendre_synthesis_simt(const Plan* pl, const float* spec, float* Z, int B, int C, cudaStream_t st) {
  const int cp = round_up(C, 4);
  const int JP = 2 * B * cp;
  dim3 grid(ceil_div(JP, TS), ceil_div(pl->kp, TS), pl->mmax);
  legendre_synthesis_simt_kernel<<<grid, 256, 0, st>>>(pl->d_table, spec, Z, pl->lmax, pl->mmax, pl->kp, B, C, cp, pl->m0);
  B200_CHECK_LAUNCH();
  return 0;
}

// -------------------------------------------------------------------------------- pack / unpack
// spec [L][M][2][B][cp]  <->  coeffs complex64 [B*C][L][M].   block: 32 m x 32 c tile of one (l, b)
__global__ void __launch_bounds__(256) spec_unpack_kernel(const float* __restrict__ spec, float2* __restrict__ coeffs, int L, int M,
                                                          int B, int C, int cp, int mo, int dense) {
  __shared__ float tile[2][32][33];
  const int l = blockIdx.z / B, b = blockIdx.z % B;
  const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  const int JP = 2 * B * cp;
  for (int mm = ty; mm < 32; mm += 8) {
    const int m = m0 + mm, c = c0 + tx;
    float re = 0.f, im = 0.f;
    if (m < M && c < C && (dense || l >= mo + m)) {  // exact zeros for l < m (global order mo + m)
      const float* row = spec + ((size_t)l * M + m) * JP;
      re = row[(0 * B + b) * cp + c];
      im = row[(1 * B + b) * cp + c];
    }
    tile[0][mm][tx] = re;
    tile[1][mm][tx] = im;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc, m = m0 + tx;
    if (c < C && m < M) coeffs[((size_t)(b * C + c) * L + l) * M + m] = make_float2(tile[0][tx][cc], tile[1][tx][cc]);
  }
}

__global__ void __launch_bounds__(256) spec_pack_kernel(const float2* __restrict__ coeffs, float* __restrict__ spec, int L, int M, int B,
                                                        int C, int cp, int mo, int dense) {
  __shared__ float tile[2][32][33];
  const int l = blockIdx.z / B, b = blockIdx.z % B;
  const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;  // c0 runs over cp
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int JP = 2 * B * cp;
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc, m = m0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (c < C && m < M) v = coeffs[((size_t)(b * C + c) * L + l) * M + m];
    tile[0][tx][cc] = v.x;
    tile[1][tx][cc] = v.y;
  }
  __syncthreads();
  for (int mm = ty; mm < 32; mm += 8) {
    const int m = m0 + mm, c = c0 + tx;
    if (m < M && c < cp && (dense || l >= lstart(mo + m))) {
      float* row = spec + ((size_t)l * M + m) * JP;
      row[(0 * B + b) * cp + c] = tile[0][mm][tx];
      row[(1 * B + b) * cp + c] = tile[1][mm][tx];
    }
  }
}

int spec_unpack(const Plan* pl, const float* spec, void* coeffs, int B, int C, cudaStream_t st) {
  const int cp = round_up(C, 4);
  dim3 grid(ceil_div(pl->mmax, 32), ceil_div(C, 32), pl->lmax * B);
  B200_REQUIRE(grid.z <= 65535, "spec_unpack: lmax*B=%u exceeds grid limit", grid.z);
  spec_unpack_kernel<<<grid, 256, 0, st>>>(spec, static_cast<float2*>(coeffs), pl->lmax, pl->mmax, B, C, cp, pl->m0, pl->dense);
  B200_CHECK_LAUNCH();
  return 0;
}

int spec_pack(const Plan* pl, const void* coeffs, float* spec, int B, int C, cudaStream_t st) {
  const int cp = round_up(C, 4);
  dim3 grid(ceil_div(pl->mmax, 32), ceil_div(cp, 32), pl->lmax * B);
  B200_REQUIRE(grid.z <= 65535, "spec_pack: lmax*B=%u exceeds grid limit", grid.z);
  spec_pack_kernel<<<grid, 256, 0, st>>>(static_cast<const float2*>(coeffs), spec, pl->lmax, pl->mmax, B, C, cp, pl->m0, pl->dense);
  B200_CHECK_LAUNCH();
  return 0;
}

// latspec [M][2][R][kp]  <->  complex64 [R][nlat][M]   (the layout the distributed transposes exchange).  block: 32 k x 32 m of one row r
__global__ void __launch_bounds__(256) latspec_convert_kernel(float* __restrict__ lat, float2* __restrict__ coeffs, int M, int R, int nlat, int kp,
                                                              int to_packed) {
  __shared__ float tile[2][32][33];
  const int r = blockIdx.z;
  const int k0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!to_packed) {
    for (int mm = ty; mm < 32; mm += 8) {
      const int m = m0 + mm, k = k0 + tx;
      float re = 0.f, im = 0.f;
      if (m < M && k < nlat) {
        re = lat[(((size_t)m * 2 + 0) * R + r) * kp + k];
        im = lat[(((size_t)m * 2 + 1) * R + r) * kp + k];
      }
      tile[0][mm][tx] = re;
      tile[1][mm][tx] = im;
    }
    __syncthreads();
    for (int kk = ty; kk < 32; kk += 8) {
      const int k = k0 + kk, m = m0 + tx;
      if (k < nlat && m < M) coeffs[((size_t)r * nlat + k) * M + m] = make_float2(tile[0][tx][kk], tile[1][tx][kk]);
    }
  } else {
    for (int kk = ty; kk < 32; kk += 8) {
      const int k = k0 + kk, m = m0 + tx;
      float2 v = make_float2(0.f, 0.f);
      if (k < nlat && m < M) v = coeffs[((size_t)r * nlat + k) * M + m];
      tile[0][tx][kk] = v.x;
      tile[1][tx][kk] = v.y;
    }
    __syncthreads();
    for (int mm = ty; mm < 32; mm += 8) {
      const int m = m0 + mm, k = k0 + tx;
      if (m < M && k < kp) {  // the k padding is written as zeros
        lat[(((size_t)m * 2 + 0) * R + r) * kp + k] = tile[0][mm][tx];
        lat[(((size_t)m * 2 + 1) * R + r) * kp + k] = tile[1][mm][tx];
      }
    }
  }
}

int latspec_convert(const Plan* pl, float* lat, void* coeffs, int B, int C, int to_packed, cudaStream_t st) {
  dim3 grid(ceil_div(pl->kp, 32), ceil_div(pl->mmax, 32), B * C);
  B200_REQUIRE(grid.z <= 65535, "latspec_convert: B*C=%u exceeds grid limit", grid.z);
  latspec_convert_kernel<<<grid, 256, 0, st>>>(lat, static_cast<float2*>(coeffs), pl->mmax, B * C, pl->nlat, pl->kp, to_packed);
  B200_CHECK_LAUNCH();
  return 0;
}

// bias gradient: gbias[c] = sum_{b,k} latspec[m=0][re][b*C+c][k]
__global__ void bias_grad_kernel(const float* __restrict__ X, float* __restrict__ gbias, int B, int C, int kp, int nlat) {
  const int c = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* row = X + ((size_t)(b * C + c)) * kp;
    for (int k = threadIdx.x; k < nlat; k += blockDim.x) s += row[k];
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) gbias[c] = v;
  }
}

int bias_grad(const Plan* pl, const float* X, float* gbias, int B, int C, cudaStream_t st) {
  bias_grad_kernel<<<C, 256, 0, st>>>(X, gbias, B, C, pl->kp, pl->nlat);
  B200_CHECK_LAUNCH();
  return 0;
}

}  // namespace b200sht

// Debug: the table recurrence on the host (same code as the device kernel) for CPU tests.
extern "C" int b200sht_debug_table_host(int nlat, int lmax, int mmax, const double* cost, int csphase, float* table /*[mmax][lmax][nlat]*/) {
  for (int m = 0; m < mmax; ++m)
    for (int k = 0; k < nlat; ++k)
      b200sht::legendre_column(m, lmax, cost[k], csphase, table + (size_t)m * lmax * nlat + k, nlat);
  return 0;
}
