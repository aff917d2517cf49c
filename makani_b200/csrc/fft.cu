// Longitude FFT stage of the SHT (replaces torch.fft.rfft / irfft inside torch_harmonics.RealSHT /
// InverseRealSHT; reference call sites /root/reference/makani/models/common/spectral_convolution.py:239,253 and
// the FFT twin /root/reference/makani/mpu/fft.py:62,109).
//
// One CTA transforms KC = 2*PAIRS consecutive latitude rows of one (batch, channel) image: two real rows are packed
// into one complex sequence, transformed with a mixed-radix Stockham FFT (radices up to 16 kept in registers, shared
// memory only for the exchange between stages), split back into the two half spectra, truncated to mmax, scaled and
// written in the "latspec" layout [m][re/im][row r][k] so that the KC results of one (m, re/im) form one contiguous
// 32-byte sector and the Legendre GEMM reads K-major operands straight from it.
//
// Two kernel families share the butterflies:
//   *_ct  : radix plan fixed at compile time (lengths 64 ... 2880 listed in CT_PLANS): all index arithmetic folds to
//           constants, first stage fused with the global load, last stage of the inverse fused with the store,
//           shared-memory indices skewed by i / R0 to spread the stride-R0 writes of the first stage over the banks.
//   *_rt  : any length whose prime factors are <= 13 (runtime plan).
//
// The stage / butterfly code is __host__ __device__ so that the same arithmetic is unit-tested on the CPU
// (b200sht_debug_fft_host) without a GPU.
#include "common.cuh"
#include "fft_roots.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace b200sht {

// ------------------------------------------------------------------------------------------------ plan
static const int kRadices[] = {16, 15, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};

static void plan_search(int n, int max_r, int* cur, int depth, int* best, int* best_len, int* best_sum) {
  if (n == 1) {
    int sum = 0;
    for (int i = 0; i < depth; ++i) sum += cur[i];
    if (depth < *best_len || (depth == *best_len && sum < *best_sum)) {
      *best_len = depth; *best_sum = sum;
      for (int i = 0; i < depth; ++i) best[i] = cur[i];
    }
    return;
  }
  if (depth >= 12 || depth + 1 > *best_len) return;
  for (int r : kRadices) {
    if (r > max_r || n % r) continue;
    cur[depth] = r;
    plan_search(n / r, r, cur, depth + 1, best, best_len, best_sum);
  }
}

// fewest stages, then smallest radix sum (balanced stages); radices in non-increasing order
bool make_fft_plan(int N, FftPlan* p) {
  p->N = N;
  p->nstages = 0;
  if (N < 2) return false;
  int n = N;
  for (int f : {2, 3, 5, 7, 11, 13})
    while (n % f == 0) n /= f;
  if (n != 1) return false;
  int cur[20], best[20], best_len = 13, best_sum = 1 << 30;
  plan_search(N, 16, cur, 0, best, &best_len, &best_sum);
  if (best_len > 12) return false;
  p->nstages = best_len;
  for (int i = 0; i < best_len; ++i) p->radix[i] = best[i];
  return true;
}

// ------------------------------------------------------------------------------------------ butterflies
HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

template <int R>
struct Butterfly;

template <>
struct Butterfly<2> {
  HD static void run(float2* v, const float2*, int) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

template <>
struct Butterfly<4> {
  HD static void run(float2* v, const float2*, int) { dft4(v[0], v[1], v[2], v[3]); }
};

template <>
struct Butterfly<8> {
  HD static void run(float2* v, const float2*, int) {
    const float h = 0.70710678118654752440f;
    float2 b0 = cadd(v[0], v[4]), b4 = csub(v[0], v[4]);
    float2 b1 = cadd(v[1], v[5]), b5 = csub(v[1], v[5]);
    float2 b2 = cadd(v[2], v[6]), b6 = csub(v[2], v[6]);
    float2 b3 = cadd(v[3], v[7]), b7 = csub(v[3], v[7]);
    // twiddles W8^1 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    b5 = make_float2(h * (b5.x + b5.y), h * (b5.y - b5.x));
    b6 = cmul_mi(b6);
    b7 = make_float2(h * (b7.y - b7.x), -h * (b7.x + b7.y));
    dft4(b0, b1, b2, b3);  // even outputs X[0], X[2], X[4], X[6]
    dft4(b4, b5, b6, b7);  // odd outputs  X[1], X[3], X[5], X[7]
    v[0] = b0; v[2] = b1; v[4] = b2; v[6] = b3;
    v[1] = b4; v[3] = b5; v[5] = b6; v[7] = b7;
  }
};

template <>
struct Butterfly<3> {
  HD static void run(float2* v, const float2*, int) {
    const float s = 0.86602540378443864676f;
    float2 t = cadd(v[1], v[2]), u = csub(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
    float2 w = make_float2(s * u.y, -s * u.x);  // -i * s * u
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, w);
    v[2] = csub(m, w);
  }
};

template <>
struct Butterfly<5> {
  HD static void run(float2* v, const float2*, int) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    float2 in1 = cmul_mi(n1), in2 = cmul_mi(n2);  // -i n
    v[1] = cadd(m1, in1);
    v[4] = csub(m1, in1);
    v[2] = cadd(m2, in2);
    v[3] = csub(m2, in2);
  }
};

// generic O(R^2) butterfly for the rare odd primes (twiddles from the length-N table; R | N)
template <int R>
struct Butterfly {
  HD static void run(float2* v, const float2* tw, int N) {
    float2 o[R];
    const int step = N / R;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float2 acc = v[0];
#pragma unroll
      for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(v[r], tw[((r * q) % R) * step]));
      o[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = o[q];
  }
};

// Cooley-Tukey composite in registers: R = R1 * R2, input index n = R2 n1 + n2, output index k = k1 + R1 k2.
template <int R1, int R2>
struct Composite {
  HD static void run(float2* v) {
    constexpr int R = R1 * R2;
    float2 t[R];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
      float2 u[R1];
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) u[n1] = v[R2 * n1 + n2];
      Butterfly<R1>::run(u, nullptr, 0);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) t[n2 * R1 + k1] = (n2 * k1 == 0) ? u[k1] : cmul(u[k1], unit_root<R>(n2 * k1));
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      float2 u[R2];
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) u[n2] = t[n2 * R1 + k1];
      Butterfly<R2>::run(u, nullptr, 0);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = u[k2];
    }
  }
};
template <> struct Butterfly<6> { HD static void run(float2* v, const float2*, int) { Composite<2, 3>::run(v); } };
template <> struct Butterfly<9> { HD static void run(float2* v, const float2*, int) { Composite<3, 3>::run(v); } };
template <> struct Butterfly<10> { HD static void run(float2* v, const float2*, int) { Composite<2, 5>::run(v); } };
template <> struct Butterfly<12> { HD static void run(float2* v, const float2*, int) { Composite<4, 3>::run(v); } };
template <> struct Butterfly<15> { HD static void run(float2* v, const float2*, int) { Composite<3, 5>::run(v); } };
template <> struct Butterfly<16> { HD static void run(float2* v, const float2*, int) { Composite<4, 4>::run(v); } };

// One Stockham butterfly (index j of N/R) of a stage with sub-transform length Ns:  in -> out  (runtime plan / host)
template <int R>
HD void stage_butterfly(const float2* in, float2* out, const float2* tw, int N, int Ns, int j) {
  const int k = j % Ns;
  const int stride = N / R;
  const int tstep = k * (N / (Ns * R));
  float2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float2 a = in[j + r * stride];
    if (r > 0 && k > 0) a = cmul(a, tw[r * tstep]);
    v[r] = a;
  }
  Butterfly<R>::run(v, tw, N);
  const int j0 = (j - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[r];
}

#define B200_RADIX_SWITCH(R, CALL)      \
  switch (R) {                          \
    case 2: { CALL(2); } break;         \
    case 3: { CALL(3); } break;         \
    case 4: { CALL(4); } break;         \
    case 5: { CALL(5); } break;         \
    case 6: { CALL(6); } break;         \
    case 7: { CALL(7); } break;         \
    case 8: { CALL(8); } break;         \
    case 9: { CALL(9); } break;         \
    case 10: { CALL(10); } break;       \
    case 11: { CALL(11); } break;       \
    case 12: { CALL(12); } break;       \
    case 13: { CALL(13); } break;       \
    case 15: { CALL(15); } break;       \
    default: { CALL(16); } break;       \
  }

// split the FFT of z = a + i b (a, b real rows) into the half spectra of a and b at mode m
HD void split_pair(float2 Z, float2 Zm /* = FFT(z)[(N-m)%N] */, float2& A, float2& Bq) {
  A = make_float2(0.5f * (Z.x + Zm.x), 0.5f * (Z.y - Zm.y));
  Bq = make_float2(0.5f * (Z.y + Zm.y), -0.5f * (Z.x - Zm.x));
}

struct FftParams {
  FftPlan fp;
  int nlat, nlon, mmax, kp;
  int R;            // B*C image rows
  int C;            // channels (bias index = r % C)
  int scale_mode;
  int round_tf32;   // analysis output feeds a tcgen05 kind::tf32 GEMM: round to nearest TF32 here
  const float2* twiddle;
  const float* rowscale;
  const float* bias;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float finish_analysis(const FftParams& prm, float v) { return prm.round_tf32 ? tf32_rn(v) : v; }

__device__ __forceinline__ float mode_scale_analysis(const FftParams& prm, int m, int k) {
  if (k >= prm.nlat) return 0.f;  // rows in the k padding are written as exact zeros
  if (prm.scale_mode == 0) return prm.rowscale[k];
  return (m == 0 || 2 * m == prm.nlon) ? 1.f : 2.f;
}

// =========================================================================================== compile-time plans
// Real rows of even length N are transformed through ONE complex FFT of length H = N/2 each (z[n] = x[2n] + i x[2n+1], loaded
// as one 4- or 8-byte word), followed by the split  X[m] = E[m] + W_N^m O[m],  E = (Z[m] + conj Z[H-m])/2,
// O = (Z[m] - conj Z[H-m])/(2i).  A CTA owns ROWS = 8 consecutive latitude rows; the threads form GROUPS groups of TPG threads,
// a group owns RPT = ROWS/GROUPS rows and one thread owns a butterfly index of those rows (twiddles and skewed indices are
// computed once per index).
template <int R0>
__host__ __device__ constexpr int skew(int i) { return i + i / R0; }

template <int H, int R0>
__host__ __device__ constexpr int ct_bufstride() { return (skew<R0>(H) + 2) | 1; }  // odd stride: rows land on different banks

// stage of a compile-time plan: smem (skewed) -> smem (skewed), rows row0 .. row0 + RPT - 1
template <int H, int R, int Ns, int R0, int TPG, int RPT>
__device__ __forceinline__ void ct_stage(const float2* in, float2* out, const float2* tws /* [R][Ns]: W^(r k H/(Ns R)) */, int bufstride, int t, int row0) {
  constexpr int NB = H / R;
  for (int j = t; j < NB; j += TPG) {
    const int k = j % Ns;
    const int j0 = (j - k) * R + k;
    float2 w[R];
    int si[R], di[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      si[r] = skew<R0>(j + r * NB);
      di[r] = skew<R0>(j0 + r * Ns);
      if (Ns > 1 && r > 0) w[r] = tws[r * Ns + k];   // consecutive threads -> consecutive k: conflict-free
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const float2* src = in + (row0 + q) * bufstride;
      float2* dst = out + (row0 + q) * bufstride;
      float2 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float2 a = src[si[r]];
        if (Ns > 1 && r > 0) a = cmul(a, w[r]);
        v[r] = a;
      }
      Butterfly<R>::run(v, nullptr, H);
#pragma unroll
      for (int r = 0; r < R; ++r) dst[di[r]] = v[r];
    }
  }
}

// per-stage twiddle tables in shared memory: tw1 [R1][R0] for the second stage (Ns = R0), tw2 [R2][R0*R1] for the third (Ns = R0*R1).
// W_H^e = W_N^(2e) comes from the plan's length-N table.
template <int R0, int R1, int R2>
__device__ __forceinline__ void ct_build_twiddles(float2* tw1, float2* tw2, const float2* __restrict__ twN, int nthreads) {
  constexpr int H = R0 * R1 * R2;
  for (int i = threadIdx.x; i < R1 * R0; i += nthreads) {
    const int r = i / R0, k = i - r * R0;
    tw1[i] = twN[2 * (r * k * (H / (R0 * R1)))];
  }
  if (R2 > 1)
    for (int i = threadIdx.x; i < R2 * R0 * R1; i += nthreads) {
      const int r = i / (R0 * R1), k = i - r * (R0 * R1);
      tw2[i] = twN[2 * (r * k)];
    }
}
template <int R0, int R1, int R2>
__host__ __device__ constexpr int ct_tw_elems() { return R1 * R0 + (R2 > 1 ? R2 * R0 * R1 : 0); }

__device__ __forceinline__ float2 ld_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16* p) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
  return make_float2(__bfloat162float(v.x), __bfloat162float(v.y));
}
__device__ __forceinline__ void st_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st_pair(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// raw element pair as loaded from global memory (converted to float2 only when stage 0 consumes it)
template <typename T> struct RawPair;
template <> struct RawPair<float> {
  float2 v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float2*>(p)); }
  __device__ __forceinline__ void zero() { v = make_float2(0.f, 0.f); }
  __device__ __forceinline__ float2 get() const { return v; }
};
template <> struct RawPair<__nv_bfloat16> {
  unsigned int v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = __ldg(reinterpret_cast<const unsigned int*>(p)); }
  __device__ __forceinline__ void zero() { v = 0u; }
  __device__ __forceinline__ float2 get() const { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
};

// x [R][nlat][nlon] -> latspec [mmax][2][R][kp]        plan (R0, R1, R2) for H = nlon / 2, R2 == 1 for two stages.
// Persistent CTAs walk the (row group, image) tiles; the stage-0 operands of the NEXT tile are loaded into registers right after
// stage 0 of the current one, so the HBM latency is hidden behind stages 1, 2 and the store pass.
template <typename T, int ROWS, int GROUPS, int TPG, int R0, int R1, int R2, int MINB>
__global__ void __launch_bounds__(GROUPS * TPG, MINB) fft_analysis_ct_kernel(const T* __restrict__ x, float* __restrict__ X, const FftParams prm) {
  constexpr int H = R0 * R1 * R2, N = 2 * H;
  constexpr int BS = ct_bufstride<H, R0>();
  constexpr int THREADS = GROUPS * TPG, RPT = ROWS / GROUPS;
  constexpr int NB0 = H / R0;
  static_assert(ROWS % GROUPS == 0 && ROWS % 4 == 0, "row grouping");
  static_assert(NB0 <= TPG, "one stage-0 butterfly index per thread (register prefetch)");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw1 = reinterpret_cast<float2*>(smem_raw);
  float2* tw2 = tw1 + R1 * R0;
  float2* b0 = tw1 + ct_tw_elems<R0, R1, R2>();
  float2* b1 = b0 + ROWS * BS;
  const int grp = threadIdx.x / TPG, t = threadIdx.x - grp * TPG;
  const int row0 = grp * RPT;
  const int ntx = (prm.kp + ROWS - 1) / ROWS;
  const int ntiles = ntx * prm.R;
  ct_build_twiddles<R0, R1, R2>(tw1, tw2, prm.twiddle, THREADS);

  RawPair<T> raw[RPT][R0];
  auto load_tile = [&](int tile) {
    const int k0 = (tile % ntx) * ROWS, r = tile / ntx;
    const T* base = x + ((size_t)r * prm.nlat + k0) * N;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = row0 + q;
      const bool valid = (t < NB0) && (k0 + row) < prm.nlat;
      const T* rp = base + (size_t)row * N + 2 * t;
#pragma unroll
      for (int rr = 0; rr < R0; ++rr) {
        if (valid) raw[q][rr].load(rp + 2 * rr * NB0);
        else raw[q][rr].zero();
      }
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) load_tile(tile);
  __syncthreads();   // twiddle tables

  for (; tile < ntiles; tile += gridDim.x) {
    const int k0 = (tile % ntx) * ROWS, r = tile / ntx;
    // ---- stage 0 from the prefetched registers
    if (t < NB0) {
      int di[R0];
#pragma unroll
      for (int rr = 0; rr < R0; ++rr) di[rr] = skew<R0>(t * R0 + rr);
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        float2 v[R0];
#pragma unroll
        for (int rr = 0; rr < R0; ++rr) v[rr] = raw[q][rr].get();
        Butterfly<R0>::run(v, nullptr, H);
        float2* dst = b0 + (row0 + q) * BS;
#pragma unroll
        for (int rr = 0; rr < R0; ++rr) dst[di[rr]] = v[rr];
      }
    }
    if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);   // in flight until the next iteration
    __syncthreads();
    ct_stage<H, R1, R0, R0, TPG, RPT>(b0, b1, tw1, BS, t, row0);
    __syncthreads();
    const float2* res = b1;
    if (R2 > 1) {
      ct_stage<H, (R2 > 1 ? R2 : 2), R0 * R1, R0, TPG, RPT>(b1, b0, tw2, BS, t, row0);
      __syncthreads();
      res = b0;
    }
    // ---- split + truncate + scale + store: a thread owns one quad of 4 rows and walks the orders m; two 16-byte stores per (m, quad)
    {
      constexpr int QUADS = ROWS / 4;
      static_assert(THREADS % QUADS == 0, "quad ownership");
      const int qd = threadIdx.x % QUADS;
      float rsc[4];   // per-row factor: quadrature weight (SHT forward) or 1 (adjoint of irfft), 0 in the latitude padding
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + qd * 4 + i;
        rsc[i] = (k < prm.nlat) ? (prm.scale_mode == 0 ? prm.rowscale[k] : 1.f) : 0.f;
      }
      const float2* rb = res + (qd * 4) * BS;
      const int kq = k0 + qd * 4;
      float* xbase = X + (size_t)r * prm.kp + kq;
      const size_t mstride = (size_t)2 * prm.R * prm.kp, pstride = (size_t)prm.R * prm.kp;
      const bool rnd = prm.round_tf32 != 0;
      for (int m = threadIdx.x / QUADS; m < prm.mmax; m += THREADS / QUADS) {
        const float2 wm = __ldg(prm.twiddle + m);                 // W_N^m
        const int im = skew<R0>(m == H ? 0 : m), ic = skew<R0>((m == 0 || m == H) ? 0 : H - m);
        const float msc = (prm.scale_mode == 1 && !(m == 0 || 2 * m == N)) ? 2.f : 1.f;
        float re[4], imv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 Z = rb[i * BS + im], Zc = rb[i * BS + ic];
          const float2 E = make_float2(0.5f * (Z.x + Zc.x), 0.5f * (Z.y - Zc.y));
          const float2 Od = make_float2(0.5f * (Z.y + Zc.y), -0.5f * (Z.x - Zc.x));   // (Z - conj Zc) / (2i)
          const float2 WO = cmul(wm, Od);
          const float sc = rsc[i] * msc;
          const float a = (E.x + WO.x) * sc, b = (E.y + WO.y) * sc;
          re[i] = rnd ? tf32_rn(a) : a;
          imv[i] = rnd ? tf32_rn(b) : b;
        }
        if (kq < prm.kp) {
          float* dst = xbase + (size_t)m * mstride;
          *reinterpret_cast<float4*>(dst) = make_float4(re[0], re[1], re[2], re[3]);
          *reinterpret_cast<float4*>(dst + pstride) = make_float4(imv[0], imv[1], imv[2], imv[3]);
        }
      }
    }
    __syncthreads();   // b0 / b1 are reused by the next tile
  }
}

// latspec [mmax][2][R][kp] -> y [R][nlat][nlon]   (persistent, with register prefetch of the next tile's spectrum)
template <typename T, int ROWS, int GROUPS, int TPG, int R0, int R1, int R2, int MINB>
__global__ void __launch_bounds__(GROUPS * TPG, MINB) fft_synthesis_ct_kernel(const float* __restrict__ Zs, T* __restrict__ y, const FftParams prm) {
  constexpr int H = R0 * R1 * R2, N = 2 * H;
  constexpr int BS = ct_bufstride<H, R0>();
  constexpr int THREADS = GROUPS * TPG, RPT = ROWS / GROUPS;
  constexpr int RL = (R2 > 1) ? R2 : R1;       // radix of the last stage (fused with the store)
  constexpr int NsL = H / RL;
  constexpr int QUADS = ROWS / 4;
  constexpr int NITEMS = (H / 2 + 1) * QUADS;
  constexpr int IPT = (NITEMS + THREADS - 1) / THREADS;   // spectrum-build items per thread
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw1 = reinterpret_cast<float2*>(smem_raw);
  float2* tw2 = tw1 + R1 * R0;
  float2* b0 = tw1 + ct_tw_elems<R0, R1, R2>();
  float2* b1 = b0 + ROWS * BS;
  const int grp = threadIdx.x / TPG, t = threadIdx.x - grp * TPG;
  const int row0 = grp * RPT;
  const int mmax = prm.mmax;
  const int ntx = (prm.kp + ROWS - 1) / ROWS;
  const int ntiles = ntx * prm.R;
  ct_build_twiddles<R0, R1, R2>(tw1, tw2, prm.twiddle, THREADS);
  const float2* twL = (R2 > 1) ? tw2 : tw1;   // table of the last stage: [RL][NsL]

  // item = (q in [0, H/2], quad of 4 rows): X[q] and X[H-q] of 4 rows (re, im) = four 16-byte loads
  float4 pa_r[IPT], pa_i[IPT], pb_r[IPT], pb_i[IPT];
  auto load_tile = [&](int tile) {
    const int k0 = (tile % ntx) * ROWS, r = tile / ntx;
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
      const int e = threadIdx.x + it * THREADS;
      const int qd = e % QUADS, q = e / QUADS, q2 = H - q;
      const int k = k0 + qd * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      pa_r[it] = z; pa_i[it] = z; pb_r[it] = z; pb_i[it] = z;
      if (e < NITEMS && k < prm.kp) {
        if (q < mmax) {
          const float* src = Zs + (((size_t)q * 2) * prm.R + r) * prm.kp + k;
          pa_r[it] = __ldg(reinterpret_cast<const float4*>(src));
          pa_i[it] = __ldg(reinterpret_cast<const float4*>(src + (size_t)prm.R * prm.kp));
        }
        if (q2 < mmax) {
          const float* src = Zs + (((size_t)q2 * 2) * prm.R + r) * prm.kp + k;
          pb_r[it] = __ldg(reinterpret_cast<const float4*>(src));
          pb_i[it] = __ldg(reinterpret_cast<const float4*>(src + (size_t)prm.R * prm.kp));
        }
      }
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) load_tile(tile);
  __syncthreads();

  for (; tile < ntiles; tile += gridDim.x) {
    const int k0 = (tile % ntx) * ROWS, r = tile / ntx;
    // ---- build Z'[q] = (X[q] + conj X[H-q]) + i (X[q] - conj X[H-q]) W_N^-q for q in [0, H), stored swapped (im, re)
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
      const int e = threadIdx.x + it * THREADS;
      if (e >= NITEMS) continue;
      const int qd = e % QUADS, q = e / QUADS;
      const int q2 = H - q;                                    // partner index (q2 == H for q == 0)
      const int k = k0 + qd * 4;
      const float ar[4] = {pa_r[it].x, pa_r[it].y, pa_r[it].z, pa_r[it].w}, ai[4] = {pa_i[it].x, pa_i[it].y, pa_i[it].z, pa_i[it].w};
      const float br[4] = {pb_r[it].x, pb_r[it].y, pb_r[it].z, pb_r[it].w}, bi[4] = {pb_i[it].x, pb_i[it].y, pb_i[it].z, pb_i[it].w};
      const bool a_self = (q == 0), b_self = (q2 == H);        // DC and Nyquist: imaginary part ignored, no halving
      const float ha = (prm.scale_mode == 1 && !a_self) ? 0.5f : 1.f;
      const float hb = (prm.scale_mode == 1 && !b_self) ? 0.5f : 1.f;
      const float2 wq = __ldg(prm.twiddle + q);                 // W_N^q ; W_N^-q = conj
      const float2 wq2 = __ldg(prm.twiddle + q2);               // q2 <= H < N
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = qd * 4 + i;
        const bool valid = (k + i) < prm.nlat;
        const float Ar = valid ? ar[i] * ha : 0.f, Ai = (valid && !a_self) ? ai[i] * ha : 0.f;   // A = X[q]
        const float Br = valid ? br[i] * hb : 0.f, Bi = (valid && !b_self) ? bi[i] * hb : 0.f;   // B = X[H-q]
        {
          const float sr = Ar + Br, si2 = Ai - Bi;              // A + conj B
          const float dr = Ar - Br, dii = Ai + Bi;              // A - conj B
          const float tr = dr * wq.x + dii * wq.y, ti = dii * wq.x - dr * wq.y;   // (A - conj B) * conj(wq)
          b0[row * BS + skew<R0>(q)] = make_float2(si2 + tr, sr - ti);            // Z'[q] = s + i t, stored (im, re)
        }
        if (q != 0 && q2 != q) {
          const float sr = Br + Ar, si2 = Bi - Ai;
          const float dr = Br - Ar, dii = Bi + Ai;
          const float tr = dr * wq2.x + dii * wq2.y, ti = dii * wq2.x - dr * wq2.y;
          b0[row * BS + skew<R0>(q2)] = make_float2(si2 + tr, sr - ti);
        }
      }
    }
    if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);   // in flight until the next iteration
    __syncthreads();
    ct_stage<H, R0, 1, R0, TPG, RPT>(b0, b1, nullptr, BS, t, row0);
    __syncthreads();
    const float2* src = b1;
    if (R2 > 1) {
      ct_stage<H, R1, R0, R0, TPG, RPT>(b1, b0, tw1, BS, t, row0);
      __syncthreads();
      src = b0;
    }
    // ---- last stage fused with the store: butterfly j yields z[e], e = j + rr * NsL, (x[2e], x[2e+1]) = (Im, Re) of the swapped result
    {
      T* base = y + ((size_t)r * prm.nlat + k0) * N;
      const float bias = prm.bias ? prm.bias[r % prm.C] : 0.f;
      for (int j = t; j < NsL; j += TPG) {
        float2 w[RL];
        int si[RL];
#pragma unroll
        for (int rr = 0; rr < RL; ++rr) {
          si[rr] = skew<R0>(j + rr * NsL);
          if (rr > 0) w[rr] = twL[rr * NsL + j];   // k = j
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          const int row = row0 + q;
          const float2* sp = src + row * BS;
          float2 v[RL];
#pragma unroll
          for (int rr = 0; rr < RL; ++rr) {
            float2 a = sp[si[rr]];
            if (rr > 0) a = cmul(a, w[rr]);
            v[rr] = a;
          }
          Butterfly<RL>::run(v, nullptr, H);
          if (k0 + row < prm.nlat) {
            const float sc = (prm.scale_mode == 1) ? prm.rowscale[k0 + row] : 1.f;
            T* rp = base + (size_t)row * N + 2 * j;
#pragma unroll
            for (int rr = 0; rr < RL; ++rr) st_pair(rp + 2 * rr * NsL, v[rr].y * sc + bias, v[rr].x * sc + bias);
          }
        }
      }
    }
    __syncthreads();   // b0 / b1 are reused by the next tile
  }
}

// Hermitian-extend the truncated half spectra of rows a = k0+2q, b = a+1 into V = Za + i Zb, stored with real/imag swapped
// (inverse FFT == swap o forward FFT o swap); `IDX` maps a spectrum index to its (possibly skewed) buffer slot.
template <class IDX>
__device__ __forceinline__ void fill_spectrum(const float* __restrict__ Zs, float2* b0, int bufstride, int pairs, int nthreads, const FftParams& prm,
                                              int k0, int r, IDX idx) {
  const int N = prm.nlon, mmax = prm.mmax;
  {
    const int lo = mmax, span = N - 2 * mmax + 1;  // untouched middle of the spectrum: indices [mmax, N - mmax]
    if (span > 0)
      for (int e = threadIdx.x; e < pairs * span; e += nthreads) {
        const int q = e / span, i = e - q * span;
        b0[q * bufstride + idx(lo + i)] = make_float2(0.f, 0.f);
      }
  }
  for (int e = threadIdx.x; e < mmax * pairs; e += nthreads) {
    const int q = e % pairs, m = e / pairs;
    const int ka = k0 + 2 * q;
    float2 re2 = make_float2(0.f, 0.f), im2 = re2;
    if (ka < prm.kp) {  // kp is a multiple of 8 and ka is even: ka + 1 < kp as well
      re2 = *reinterpret_cast<const float2*>(Zs + (((size_t)m * 2 + 0) * prm.R + r) * prm.kp + ka);
      im2 = *reinterpret_cast<const float2*>(Zs + (((size_t)m * 2 + 1) * prm.R + r) * prm.kp + ka);
    }
    float ar = re2.x, br = re2.y, ai = im2.x, bi = im2.y;
    if (ka >= prm.nlat) { ar = 0.f; ai = 0.f; }
    if (ka + 1 >= prm.nlat) { br = 0.f; bi = 0.f; }
    const bool self_conj = (m == 0) || (2 * m == N);
    if (self_conj) { ai = 0.f; bi = 0.f; }
    if (prm.scale_mode == 1 && !self_conj) { ar *= 0.5f; ai *= 0.5f; br *= 0.5f; bi *= 0.5f; }
    // V[m] = (ar - bi) + i (ai + br);  V[N-m] = (ar + bi) + i (br - ai)   -- stored swapped (y, x)
    b0[q * bufstride + idx(m)] = make_float2(ai + br, ar - bi);
    if (!self_conj) b0[q * bufstride + idx(N - m)] = make_float2(br - ai, ar + bi);
  }
}

// ================================================================================================ runtime plans
constexpr int kFftThreads = 256;

template <int R>
__device__ __forceinline__ void run_stage(const float2* in, float2* out, const float2* tw, int N, int Ns, int pairs, int bufstride) {
  const int nb = N / R;
  for (int w = threadIdx.x; w < pairs * nb; w += kFftThreads) {
    const int q = w / nb, j = w - q * nb;
    stage_butterfly<R>(in + q * bufstride, out + q * bufstride, tw, N, Ns, j);
  }
}

// runs all stages; returns pointer to the buffer holding the result
__device__ __forceinline__ float2* run_fft(float2* b0, float2* b1, const float2* tw, const FftPlan& fp, int pairs, int bufstride) {
  float2* in = b0;
  float2* out = b1;
  int Ns = 1;
  const int N = fp.N;
  for (int s = 0; s < fp.nstages; ++s) {
    const int R = fp.radix[s];
#define CALL(RR) run_stage<RR>(in, out, tw, N, Ns, pairs, bufstride)
    B200_RADIX_SWITCH(R, CALL)
#undef CALL
    Ns *= R;
    __syncthreads();
    float2* t = in; in = out; out = t;
  }
  return in;
}

template <typename T, int PAIRS>
__global__ void __launch_bounds__(kFftThreads) fft_analysis_rt_kernel(const T* __restrict__ x, float* __restrict__ X, const FftParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = prm.nlon;
  const int NS = N + 1;  // padded per-pair stride (float2 units)
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* b0 = tw + N;
  float2* b1 = b0 + PAIRS * NS;
  constexpr int KC = 2 * PAIRS;
  const int k0 = blockIdx.x * KC;
  const int r = blockIdx.y;
  for (int t = threadIdx.x; t < N; t += kFftThreads) tw[t] = prm.twiddle[t];
  {
    float* bf = reinterpret_cast<float*>(b0);
    const T* base = x + ((size_t)r * prm.nlat + k0) * N;
    const int rows_valid = min(KC, prm.nlat - k0);
    for (int e = threadIdx.x; e < KC * N; e += kFftThreads) {
      const int kk = e / N, j = e - kk * N;
      const float v = (kk < rows_valid) ? ld_as_float(base + (size_t)kk * N + j) : 0.f;
      bf[((size_t)(kk >> 1) * NS + j) * 2 + (kk & 1)] = v;
    }
  }
  __syncthreads();
  float2* res = run_fft(b0, b1, tw, prm.fp, PAIRS, NS);
  const int total = prm.mmax * 2 * KC;
  for (int e = threadIdx.x; e < total; e += kFftThreads) {
    const int kk = e % KC;
    const int mp = e / KC;
    const int p = mp & 1, m = mp >> 1;
    const int q = kk >> 1;
    float2 A, Bq;
    split_pair(res[q * NS + m], res[q * NS + (m == 0 ? 0 : N - m)], A, Bq);
    const float2 val = (kk & 1) ? Bq : A;
    const int k = k0 + kk;
    X[(((size_t)m * 2 + p) * prm.R + r) * prm.kp + k] = finish_analysis(prm, (p ? val.y : val.x) * mode_scale_analysis(prm, m, k));
  }
}

template <typename T, int PAIRS>
__global__ void __launch_bounds__(kFftThreads) fft_synthesis_rt_kernel(const float* __restrict__ Zs, T* __restrict__ y, const FftParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = prm.nlon;
  const int NS = N + 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* b0 = tw + N;
  float2* b1 = b0 + PAIRS * NS;
  constexpr int KC = 2 * PAIRS;
  const int k0 = blockIdx.x * KC;
  const int r = blockIdx.y;
  for (int t = threadIdx.x; t < N; t += kFftThreads) tw[t] = prm.twiddle[t];
  fill_spectrum(Zs, b0, NS, PAIRS, kFftThreads, prm, k0, r, [](int i) { return i; });
  __syncthreads();
  float2* res = run_fft(b0, b1, tw, prm.fp, PAIRS, NS);
  {
    const float* rf = reinterpret_cast<const float*>(res);
    T* base = y + ((size_t)r * prm.nlat + k0) * N;
    const int rows_valid = min(KC, prm.nlat - k0);
    const float bias = prm.bias ? prm.bias[r % prm.C] : 0.f;
    for (int e = threadIdx.x; e < rows_valid * N; e += kFftThreads) {
      const int kk = e / N, j = e - kk * N;
      const float sc = (prm.scale_mode == 1) ? prm.rowscale[k0 + kk] : 1.f;
      st_from_float(base + (size_t)kk * N + j, rf[((size_t)(kk >> 1) * NS + j) * 2 + (1 - (kk & 1))] * sc + bias);
    }
  }
}

// ===================================================================================================== dispatch
static size_t rt_smem_bytes(int N, int pairs) { return sizeof(float2) * ((size_t)N + 2 * (size_t)pairs * (N + 1)); }

static int rt_pick_pairs(int N) {
  if (rt_smem_bytes(N, 4) <= 110 * 1024) return 4;
  if (rt_smem_bytes(N, 2) <= 220 * 1024) return 2;
  if (rt_smem_bytes(N, 1) <= 220 * 1024) return 1;
  return 0;
}

static FftParams make_params(const Plan* pl, int B, int C, int scale_mode, const float* bias) {
  FftParams prm;
  prm.fp = pl->fft;
  prm.nlat = pl->nlat; prm.nlon = pl->nlon; prm.mmax = pl->mmax; prm.kp = pl->kp;
  prm.R = B * C; prm.C = C; prm.scale_mode = scale_mode & 1; prm.round_tf32 = (scale_mode >> 1) & 1;
  prm.twiddle = pl->d_twiddle; prm.rowscale = pl->d_rowscale; prm.bias = bias;
  return prm;
}

template <typename T, int ROWS, int GROUPS, int TPG, int R0, int R1, int R2, int MINB>
static int launch_ct(const Plan* pl, int dir, const void* in, void* out, const FftParams& prm, cudaStream_t st) {
  constexpr int H = R0 * R1 * R2;
  constexpr size_t smem = sizeof(float2) * ((size_t)ct_tw_elems<R0, R1, R2>() + 2 * ROWS * ct_bufstride<H, R0>());
  static_assert(smem <= 227 * 1024, "plan does not fit in shared memory");
  // persistent CTAs: as many as fit concurrently (by shared memory), each walks tiles blockIdx.x, + gridDim.x, ...
  const int ntiles = ceil_div(pl->kp, ROWS) * prm.R;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const int sms = pl->sm_count > 0 ? pl->sm_count : 148;
  dim3 grid(ntiles < per_sm * sms ? ntiles : per_sm * sms);
  if (dir == 0) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(fft_analysis_ct_kernel<T, ROWS, GROUPS, TPG, R0, R1, R2, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fft_analysis_ct_kernel<T, ROWS, GROUPS, TPG, R0, R1, R2, MINB><<<grid, GROUPS * TPG, smem, st>>>(static_cast<const T*>(in), static_cast<float*>(out), prm);
  } else {
    B200_CHECK_CUDA(cudaFuncSetAttribute(fft_synthesis_ct_kernel<T, ROWS, GROUPS, TPG, R0, R1, R2, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fft_synthesis_ct_kernel<T, ROWS, GROUPS, TPG, R0, R1, R2, MINB><<<grid, GROUPS * TPG, smem, st>>>(static_cast<const float*>(in), static_cast<T*>(out), prm);
  }
  B200_CHECK_LAUNCH();
  return 0;
}

// lengths with a compile-time plan: (ROWS, GROUPS, TPG, R0, R1, R2) for H = nlon / 2 = R0*R1*R2.  R0 is a power of two (the skew
// i + i/R0 is a shift); TPG ~ max_s H/R_s.  Other lengths (odd, or not listed) run the runtime-plan kernels.
#define CT_PLANS(X)             \
  X(8, 2, 96, 8, 10, 9, 2)      /* nlon 1440: 2 groups x 4 rows per thread (no spills at 2 CTAs/SM; measured faster than 4 x 2) */ \
  X(8, 4, 96, 8, 9, 5, 2)       /* nlon  720 */ \
  X(8, 4, 64, 8, 6, 5, 2)       /* nlon  480 */ \
  X(8, 4, 64, 4, 9, 5, 2)       /* nlon  360 */ \
  X(8, 4, 64, 8, 5, 3, 2)       /* nlon  240 */ \
  X(8, 4, 64, 2, 9, 5, 2)       /* nlon  180 */ \
  X(8, 4, 32, 8, 3, 3, 2)       /* nlon  144 */ \
  X(8, 8, 32, 4, 4, 4, 2)       /* nlon  128 */ \
  X(8, 8, 32, 4, 4, 3, 2)       /* nlon   96 */ \
  X(8, 8, 32, 4, 3, 3, 2)       /* nlon   72 */ \
  X(8, 8, 32, 4, 8, 1, 2)       /* nlon   64 */ \
  X(8, 4, 32, 8, 4, 4, 2)       /* nlon  256 */ \
  X(8, 4, 64, 8, 8, 4, 2)       /* nlon  512 */ \
  X(8, 4, 64, 8, 8, 8, 2)       /* nlon 1024 */ \
  X(8, 2, 160, 16, 10, 9, 1)    /* nlon 2880 */

template <typename T>
static int dispatch_ct(const Plan* pl, int dir, const void* in, void* out, const FftParams& prm, cudaStream_t st, bool* handled) {
  *handled = false;
  // the compile-time plans move element pairs / quads with vector loads: both tensors must be 16-byte aligned
  if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return 0;
  *handled = true;
  // experiment switch (B200SHT_FFT_VARIANT=1): 1440-point rows with 4 thread groups x 2 rows per thread
  static const int variant = [] { const char* e = getenv("B200SHT_FFT_VARIANT"); return e ? atoi(e) : 0; }();
  if (variant == 1 && pl->nlon == 1440) return launch_ct<T, 8, 4, 96, 8, 10, 9, 2>(pl, dir, in, out, prm, st);
#define X(RW, G, TP, A, B_, C_, MB) \
  if (pl->nlon == 2 * (A) * (B_) * (C_)) return launch_ct<T, RW, G, TP, A, B_, C_, MB>(pl, dir, in, out, prm, st);
  CT_PLANS(X)
#undef X
  *handled = false;
  return 0;
}

// pairs per CTA for plan lookup by the grid computation (kp / (2 * pairs) must be integral: kp is a multiple of 8)
template <typename T>
static int launch_rt(const Plan* pl, int dir, const void* in, void* out, const FftParams& prm, cudaStream_t st) {
  const int pairs = rt_pick_pairs(pl->nlon);
  if (pairs == 0) { set_error("fft: nlon=%d too large for shared memory", pl->nlon); return B200SHT_ERR_UNSUPPORTED; }
  const size_t smem = rt_smem_bytes(pl->nlon, pairs);
  dim3 grid(pl->kp / (2 * pairs), prm.R);
#define LAUNCH(P)                                                                                                                          \
  if (dir == 0) {                                                                                                                          \
    B200_CHECK_CUDA(cudaFuncSetAttribute(fft_analysis_rt_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    fft_analysis_rt_kernel<T, P><<<grid, kFftThreads, smem, st>>>(static_cast<const T*>(in), static_cast<float*>(out), prm);                \
  } else {                                                                                                                                 \
    B200_CHECK_CUDA(cudaFuncSetAttribute(fft_synthesis_rt_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    fft_synthesis_rt_kernel<T, P><<<grid, kFftThreads, smem, st>>>(static_cast<const float*>(in), static_cast<T*>(out), prm);               \
  }
  if (pairs == 4) { LAUNCH(4) } else if (pairs == 2) { LAUNCH(2) } else { LAUNCH(1) }
#undef LAUNCH
  B200_CHECK_LAUNCH();
  return 0;
}

template <typename T>
static int run_fft_dir(const Plan* pl, int dir, const void* in, void* out, const FftParams& prm, cudaStream_t st) {
  bool handled = false;
  int rc = dispatch_ct<T>(pl, dir, in, out, prm, st, &handled);
  if (handled) return rc;
  return launch_rt<T>(pl, dir, in, out, prm, st);
}

int fft_analysis(const Plan* pl, const void* x, int dtype, int B, int C, float* X, int scale_mode, cudaStream_t st) {
  B200_REQUIRE(B > 0 && C > 0 && (long long)B * C <= 65535, "fft_analysis: B*C=%lld out of range", (long long)B * C);
  FftParams prm = make_params(pl, B, C, scale_mode, nullptr);
  if (dtype == B200SHT_F32) return run_fft_dir<float>(pl, 0, x, X, prm, st);
  if (dtype == B200SHT_BF16) return run_fft_dir<__nv_bfloat16>(pl, 0, x, X, prm, st);
  set_error("fft_analysis: unknown dtype %d", dtype);
  return B200SHT_ERR_INVALID;
}

int fft_synthesis(const Plan* pl, const float* Z, void* y, int dtype, int B, int C, const float* bias, int scale_mode, cudaStream_t st) {
  B200_REQUIRE(B > 0 && C > 0 && (long long)B * C <= 65535, "fft_synthesis: B*C=%lld out of range", (long long)B * C);
  FftParams prm = make_params(pl, B, C, scale_mode, bias);
  if (dtype == B200SHT_F32) return run_fft_dir<float>(pl, 1, Z, y, prm, st);
  if (dtype == B200SHT_BF16) return run_fft_dir<__nv_bfloat16>(pl, 1, Z, y, prm, st);
  set_error("fft_synthesis: unknown dtype %d", dtype);
  return B200SHT_ERR_INVALID;
}

// ---------------------------------------------------------------------------- host emulation (CPU tests)
static void host_fft(std::vector<float2>& a, const std::vector<float2>& tw, const FftPlan& fp) {
  const int N = fp.N;
  std::vector<float2> b(N);
  float2* in = a.data();
  float2* out = b.data();
  int Ns = 1;
  for (int s = 0; s < fp.nstages; ++s) {
    const int R = fp.radix[s];
    for (int j = 0; j < N / R; ++j) {
#define CALL(RR) stage_butterfly<RR>(in, out, tw.data(), N, Ns, j)
      B200_RADIX_SWITCH(R, CALL)
#undef CALL
    }
    Ns *= R;
    float2* t = in; in = out; out = t;
  }
  if (in != a.data()) for (int i = 0; i < N; ++i) a[i] = in[i];
}

void make_twiddles_host(int N, std::vector<float2>& tw) {
  tw.resize(N);
  for (int t = 0; t < N; ++t) {
    const double ang = -2.0 * M_PI * (double)t / (double)N;
    tw[t] = make_float2((float)cos(ang), (float)sin(ang));
  }
}

}  // namespace b200sht

using namespace b200sht;

// Debug entry points: run the *same* stage/butterfly/split code on the host (no GPU needed).
//   analysis : rows a, b (float[N]) -> Xa, Xb (float[2*mmax] interleaved), unscaled rfft
//   synthesis: Za, Zb (float[2*mmax]) -> rows a, b (float[N]) with irfft(norm="forward") semantics
extern "C" int b200sht_debug_fft_host(int N, int mmax, int direction, const float* in_a, const float* in_b, float* out_a, float* out_b) {
  FftPlan fp;
  if (!make_fft_plan(N, &fp)) { set_error("debug_fft_host: unsupported length %d", N); return B200SHT_ERR_UNSUPPORTED; }
  std::vector<float2> tw;
  make_twiddles_host(N, tw);
  std::vector<float2> buf(N);
  if (direction == 0) {
    for (int j = 0; j < N; ++j) buf[j] = make_float2(in_a[j], in_b[j]);
    host_fft(buf, tw, fp);
    for (int m = 0; m < mmax; ++m) {
      float2 A, Bq;
      split_pair(buf[m], buf[m == 0 ? 0 : N - m], A, Bq);
      out_a[2 * m] = A.x; out_a[2 * m + 1] = A.y;
      out_b[2 * m] = Bq.x; out_b[2 * m + 1] = Bq.y;
    }
  } else {
    for (int j = 0; j < N; ++j) buf[j] = make_float2(0.f, 0.f);
    for (int m = 0; m < mmax; ++m) {
      float ar = in_a[2 * m], ai = in_a[2 * m + 1], br = in_b[2 * m], bi = in_b[2 * m + 1];
      const bool self_conj = (m == 0) || (2 * m == N);
      if (self_conj) { ai = 0.f; bi = 0.f; }
      buf[m] = make_float2(ai + br, ar - bi);
      if (!self_conj) buf[N - m] = make_float2(br - ai, ar + bi);
    }
    host_fft(buf, tw, fp);
    for (int j = 0; j < N; ++j) { out_a[j] = buf[j].y; out_b[j] = buf[j].x; }
  }
  return 0;
}

extern "C" int b200sht_debug_fft_plan(int N, int* radices, int max_radices) {
  FftPlan fp;
  if (!make_fft_plan(N, &fp)) return B200SHT_ERR_UNSUPPORTED;
  for (int i = 0; i < fp.nstages && i < max_radices; ++i) radices[i] = fp.radix[i];
  return fp.nstages;
}
