// Longitude FFT stage of the SHT (replaces torch.fft.rfft / irfft inside torch_harmonics.RealSHT /
// InverseRealSHT; reference call sites /root/reference/makani/models/common/spectral_convolution.py:239,253 and
// the FFT twin /root/reference/makani/mpu/fft.py:62,109).
//
// One CTA transforms KC = 2*PAIRS consecutive latitude rows of one (batch, channel) image: the rows are contiguous
// in memory (coalesced, vectorised loads), two real rows are packed into one complex sequence, transformed with a
// mixed-radix Stockham FFT in shared memory, split back into the two half spectra, truncated to mmax, scaled and
// written in the "latspec" layout [m][re/im][row r][k] so that the KC results of one (m, re/im) form one contiguous
// 32-byte sector and the Legendre GEMM can read K-major operands straight from it.
//
// The stage / butterfly code is __host__ __device__ so that the same arithmetic is unit-tested on the CPU
// (b200sht_debug_fft_host) without a GPU.
#include "common.cuh"
#include <cmath>
#include <vector>

namespace b200sht {

// ------------------------------------------------------------------------------------------------ plan
bool make_fft_plan(int N, FftPlan* p) {
  p->N = N;
  p->nstages = 0;
  if (N < 2) return false;
  int n = N;
  int twos = 0;
  while (n % 2 == 0) { n /= 2; ++twos; }
  const int odd[] = {3, 5, 7, 11, 13};
  int tmp[20];
  int cnt = 0;
  for (int r : odd)
    while (n % r == 0) { n /= r; if (cnt >= 16) return false; tmp[cnt++] = r; }
  if (n != 1) return false;
  // powers of two: as many radix-8 as possible, remainder as 4 or 2 (4*4 preferred over 8*2)
  int e8 = twos / 3, rem = twos % 3;
  int e4 = 0, e2 = 0;
  if (rem == 1) { if (e8 >= 1) { e8 -= 1; e4 = 2; } else e2 = 1; }
  if (rem == 2) e4 += 1;
  // small radices first (keeps Ns small while strides are large), odd ones last
  for (int i = 0; i < e2; ++i) p->radix[p->nstages++] = 2;
  for (int i = 0; i < e4; ++i) p->radix[p->nstages++] = 4;
  for (int i = 0; i < e8; ++i) p->radix[p->nstages++] = 8;
  for (int i = 0; i < cnt; ++i) p->radix[p->nstages++] = tmp[i];
  return p->nstages <= 20;
}

// ------------------------------------------------------------------------------------------ butterflies
HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
HD float2 cmul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)

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

// One Stockham butterfly (index j of N/R) of a stage with sub-transform length Ns:  in -> out.
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

// split the FFT of z = a + i b (a, b real rows) into the half spectra of a and b at mode m
HD void split_pair(float2 Z, float2 Zm /* = FFT(z)[(N-m)%N] */, float2& A, float2& Bq) {
  A = make_float2(0.5f * (Z.x + Zm.x), 0.5f * (Z.y - Zm.y));
  Bq = make_float2(0.5f * (Z.y + Zm.y), -0.5f * (Z.x - Zm.x));
}

// ------------------------------------------------------------------------------------------- kernels
constexpr int kFftThreads = 256;

template <int R>
__device__ __forceinline__ void run_stage(const float2* in, float2* out, const float2* tw, int N, int Ns, int pairs,
                                          int bufstride) {
  const int nb = N / R;
  for (int w = threadIdx.x; w < pairs * nb; w += kFftThreads) {
    const int q = w / nb, j = w - q * nb;
    stage_butterfly<R>(in + q * bufstride, out + q * bufstride, tw, N, Ns, j);
  }
}

// runs all stages; returns pointer to the buffer holding the result
__device__ __forceinline__ float2* run_fft(float2* b0, float2* b1, const float2* tw, const FftPlan& fp, int pairs,
                                           int bufstride) {
  float2* in = b0;
  float2* out = b1;
  int Ns = 1;
  const int N = fp.N;
  for (int s = 0; s < fp.nstages; ++s) {
    const int R = fp.radix[s];
    switch (R) {
      case 2: run_stage<2>(in, out, tw, N, Ns, pairs, bufstride); break;
      case 3: run_stage<3>(in, out, tw, N, Ns, pairs, bufstride); break;
      case 4: run_stage<4>(in, out, tw, N, Ns, pairs, bufstride); break;
      case 5: run_stage<5>(in, out, tw, N, Ns, pairs, bufstride); break;
      case 7: run_stage<7>(in, out, tw, N, Ns, pairs, bufstride); break;
      case 8: run_stage<8>(in, out, tw, N, Ns, pairs, bufstride); break;
      case 11: run_stage<11>(in, out, tw, N, Ns, pairs, bufstride); break;
      default: run_stage<13>(in, out, tw, N, Ns, pairs, bufstride); break;
    }
    Ns *= R;
    __syncthreads();
    float2* t = in; in = out; out = t;
  }
  return in;
}

struct FftParams {
  FftPlan fp;
  int nlat, nlon, mmax, kp;
  int R;            // B*C image rows
  int C;            // channels (bias index = r % C)
  int scale_mode;
  const float2* twiddle;
  const float* rowscale;
  const float* bias;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// x [R][nlat][nlon] -> latspec [mmax][2][R][kp]
template <typename T, int PAIRS>
__global__ void __launch_bounds__(kFftThreads) fft_analysis_kernel(const T* __restrict__ x, float* __restrict__ X,
                                                                   const FftParams prm) {
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

  // ---- load KC rows (contiguous in memory), packing row pairs into complex sequences
  {
    float* bf = reinterpret_cast<float*>(b0);
    const T* base = x + ((size_t)r * prm.nlat + k0) * N;
    const int rows_valid = min(KC, prm.nlat - k0);
    constexpr int V = 16 / sizeof(T);
    const bool vec_ok = (N % V == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    if (vec_ok) {
      const int nv = N / V;
      for (int e = threadIdx.x; e < KC * nv; e += kFftThreads) {
        const int kk = e / nv, jv = e - kk * nv;
        float vals[V];
        if (kk < rows_valid) {
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + (size_t)kk * N) + jv);
          const T* pv = reinterpret_cast<const T*>(&raw);
#pragma unroll
          for (int i = 0; i < V; ++i) vals[i] = ld_as_float(pv + i);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) vals[i] = 0.f;
        }
        float* dst = bf + ((size_t)(kk >> 1) * NS + jv * V) * 2 + (kk & 1);
#pragma unroll
        for (int i = 0; i < V; ++i) dst[2 * i] = vals[i];
      }
    } else {
      for (int e = threadIdx.x; e < KC * N; e += kFftThreads) {
        const int kk = e / N, j = e - kk * N;
        const float v = (kk < rows_valid) ? ld_as_float(base + (size_t)kk * N + j) : 0.f;
        bf[((size_t)(kk >> 1) * NS + j) * 2 + (kk & 1)] = v;
      }
    }
  }
  __syncthreads();

  float2* res = run_fft(b0, b1, tw, prm.fp, PAIRS, NS);

  // ---- split, truncate, scale, store: item = (m, p, kk), kk fastest -> 32-byte sectors
  const int total = prm.mmax * 2 * KC;
  for (int e = threadIdx.x; e < total; e += kFftThreads) {
    const int kk = e % KC;
    const int mp = e / KC;
    const int p = mp & 1, m = mp >> 1;
    const int q = kk >> 1;
    const float2 Z = res[q * NS + m];
    const float2 Zm = res[q * NS + (m == 0 ? 0 : N - m)];
    float2 A, Bq;
    split_pair(Z, Zm, A, Bq);
    const float2 val = (kk & 1) ? Bq : A;
    float v = p ? val.y : val.x;
    const int k = k0 + kk;
    float sc;
    if (prm.scale_mode == 0) sc = prm.rowscale[k];
    else sc = (m == 0 || 2 * m == N) ? 1.f : 2.f;
    // rows in the k padding (k >= nlat) are written as exact zeros (pair splitting leaves rounding noise there)
    X[(((size_t)m * 2 + p) * prm.R + r) * prm.kp + k] = (k < prm.nlat) ? v * sc : 0.f;
  }
}

// latspec [mmax][2][R][kp] -> y [R][nlat][nlon]
template <typename T, int PAIRS>
__global__ void __launch_bounds__(kFftThreads) fft_synthesis_kernel(const float* __restrict__ Zs, T* __restrict__ y,
                                                                    const FftParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = prm.nlon;
  const int NS = N + 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* b0 = tw + N;
  float2* b1 = b0 + PAIRS * NS;
  constexpr int KC = 2 * PAIRS;
  const int k0 = blockIdx.x * KC;
  const int r = blockIdx.y;
  const int mmax = prm.mmax;

  for (int t = threadIdx.x; t < N; t += kFftThreads) tw[t] = prm.twiddle[t];
  // zero the untouched middle of the spectrum: indices [mmax, N - mmax]
  {
    const int lo = mmax, hi = N - mmax;  // inclusive
    const int span = hi - lo + 1;
    if (span > 0)
      for (int e = threadIdx.x; e < PAIRS * span; e += kFftThreads) {
        const int q = e / span, i = e - q * span;
        b0[q * NS + lo + i] = make_float2(0.f, 0.f);
      }
  }
  // Hermitian-extend the truncated half spectra of rows a = k0+2q, b = a+1 into V = Za + i Zb and store it with
  // real/imag swapped (inverse FFT == swap o forward FFT o swap).
  for (int e = threadIdx.x; e < mmax * PAIRS; e += kFftThreads) {
    const int q = e % PAIRS, m = e / PAIRS;
    const int ka = k0 + 2 * q;
    const float2 re2 = *reinterpret_cast<const float2*>(Zs + (((size_t)m * 2 + 0) * prm.R + r) * prm.kp + ka);
    const float2 im2 = *reinterpret_cast<const float2*>(Zs + (((size_t)m * 2 + 1) * prm.R + r) * prm.kp + ka);
    float ar = re2.x, br = re2.y, ai = im2.x, bi = im2.y;
    if (ka >= prm.nlat) { ar = 0.f; ai = 0.f; }
    if (ka + 1 >= prm.nlat) { br = 0.f; bi = 0.f; }
    const bool self_conj = (m == 0) || (2 * m == N);
    if (self_conj) { ai = 0.f; bi = 0.f; }
    if (prm.scale_mode == 1 && !self_conj) { ar *= 0.5f; ai *= 0.5f; br *= 0.5f; bi *= 0.5f; }
    // V[m] = (ar - bi) + i (ai + br);  V[N-m] = (ar + bi) + i (br - ai)   -- stored swapped (y, x)
    b0[q * NS + m] = make_float2(ai + br, ar - bi);
    if (!self_conj) b0[q * NS + N - m] = make_float2(br - ai, ar + bi);
  }
  __syncthreads();

  float2* res = run_fft(b0, b1, tw, prm.fp, PAIRS, NS);

  // ---- store rows: a[j] = res.y, b[j] = res.x (swapped back)
  {
    const float* rf = reinterpret_cast<const float*>(res);
    T* base = y + ((size_t)r * prm.nlat + k0) * N;
    const int rows_valid = min(KC, prm.nlat - k0);
    const float bias = prm.bias ? prm.bias[r % prm.C] : 0.f;
    constexpr int V = 16 / sizeof(T);
    const bool vec_ok = (N % V == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    if (vec_ok) {
      const int nv = N / V;
      for (int e = threadIdx.x; e < rows_valid * nv; e += kFftThreads) {
        const int kk = e / nv, jv = e - kk * nv;
        const float sc = (prm.scale_mode == 1) ? prm.rowscale[k0 + kk] : 1.f;
        const float* src = rf + ((size_t)(kk >> 1) * NS + jv * V) * 2 + (1 - (kk & 1));
        uint4 raw;
        T* pv = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int i = 0; i < V; ++i) st_from_float(pv + i, src[2 * i] * sc + bias);
        *(reinterpret_cast<uint4*>(base + (size_t)kk * N) + jv) = raw;
      }
    } else {
      for (int e = threadIdx.x; e < rows_valid * N; e += kFftThreads) {
        const int kk = e / N, j = e - kk * N;
        const float sc = (prm.scale_mode == 1) ? prm.rowscale[k0 + kk] : 1.f;
        st_from_float(base + (size_t)kk * N + j, rf[((size_t)(kk >> 1) * NS + j) * 2 + (1 - (kk & 1))] * sc + bias);
      }
    }
  }
}

static size_t fft_smem_bytes(int N, int pairs) { return sizeof(float2) * ((size_t)N + 2 * (size_t)pairs * (N + 1)); }

template <typename T, int PAIRS>
static int launch_analysis(const Plan* pl, const T* x, float* X, FftParams prm, cudaStream_t st) {
  const size_t smem = fft_smem_bytes(pl->nlon, PAIRS);
  B200_CHECK_CUDA(cudaFuncSetAttribute(fft_analysis_kernel<T, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(pl->kp / (2 * PAIRS), prm.R);
  fft_analysis_kernel<T, PAIRS><<<grid, kFftThreads, smem, st>>>(x, X, prm);
  B200_CHECK_LAUNCH();
  return 0;
}

template <typename T, int PAIRS>
static int launch_synthesis(const Plan* pl, const float* Z, T* y, FftParams prm, cudaStream_t st) {
  const size_t smem = fft_smem_bytes(pl->nlon, PAIRS);
  B200_CHECK_CUDA(cudaFuncSetAttribute(fft_synthesis_kernel<T, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(pl->kp / (2 * PAIRS), prm.R);
  fft_synthesis_kernel<T, PAIRS><<<grid, kFftThreads, smem, st>>>(Z, y, prm);
  B200_CHECK_LAUNCH();
  return 0;
}

static int pick_pairs(int N) {
  // shared memory budget: aim for >= 2 CTAs per SM when possible (227 KB per SM)
  if (fft_smem_bytes(N, 4) <= 110 * 1024) return 4;
  if (fft_smem_bytes(N, 2) <= 220 * 1024) return 2;
  if (fft_smem_bytes(N, 1) <= 220 * 1024) return 1;
  return 0;
}

static FftParams make_params(const Plan* pl, int B, int C, int scale_mode, const float* bias) {
  FftParams prm;
  prm.fp = pl->fft;
  prm.nlat = pl->nlat; prm.nlon = pl->nlon; prm.mmax = pl->mmax; prm.kp = pl->kp;
  prm.R = B * C; prm.C = C; prm.scale_mode = scale_mode;
  prm.twiddle = pl->d_twiddle; prm.rowscale = pl->d_rowscale; prm.bias = bias;
  return prm;
}

int fft_analysis(const Plan* pl, const void* x, int dtype, int B, int C, float* X, int scale_mode, cudaStream_t st) {
  B200_REQUIRE(B > 0 && C > 0 && (long long)B * C <= 65535, "fft_analysis: B*C=%lld out of range", (long long)B * C);
  const int pairs = pick_pairs(pl->nlon);
  if (pairs == 0) { set_error("fft_analysis: nlon=%d too large for shared memory", pl->nlon); return B200SHT_ERR_UNSUPPORTED; }
  FftParams prm = make_params(pl, B, C, scale_mode, nullptr);
#define DISPATCH(T, P) return launch_analysis<T, P>(pl, static_cast<const T*>(x), X, prm, st)
  if (dtype == B200SHT_F32) {
    if (pairs == 4) DISPATCH(float, 4);
    if (pairs == 2) DISPATCH(float, 2);
    DISPATCH(float, 1);
  } else if (dtype == B200SHT_BF16) {
    if (pairs == 4) DISPATCH(__nv_bfloat16, 4);
    if (pairs == 2) DISPATCH(__nv_bfloat16, 2);
    DISPATCH(__nv_bfloat16, 1);
  }
#undef DISPATCH
  set_error("fft_analysis: unknown dtype %d", dtype);
  return B200SHT_ERR_INVALID;
}

int fft_synthesis(const Plan* pl, const float* Z, void* y, int dtype, int B, int C, const float* bias, int scale_mode,
                  cudaStream_t st) {
  B200_REQUIRE(B > 0 && C > 0 && (long long)B * C <= 65535, "fft_synthesis: B*C=%lld out of range", (long long)B * C);
  const int pairs = pick_pairs(pl->nlon);
  if (pairs == 0) { set_error("fft_synthesis: nlon=%d too large for shared memory", pl->nlon); return B200SHT_ERR_UNSUPPORTED; }
  FftParams prm = make_params(pl, B, C, scale_mode, bias);
#define DISPATCH(T, P) return launch_synthesis<T, P>(pl, Z, static_cast<T*>(y), prm, st)
  if (dtype == B200SHT_F32) {
    if (pairs == 4) DISPATCH(float, 4);
    if (pairs == 2) DISPATCH(float, 2);
    DISPATCH(float, 1);
  } else if (dtype == B200SHT_BF16) {
    if (pairs == 4) DISPATCH(__nv_bfloat16, 4);
    if (pairs == 2) DISPATCH(__nv_bfloat16, 2);
    DISPATCH(__nv_bfloat16, 1);
  }
#undef DISPATCH
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
      switch (R) {
        case 2: stage_butterfly<2>(in, out, tw.data(), N, Ns, j); break;
        case 3: stage_butterfly<3>(in, out, tw.data(), N, Ns, j); break;
        case 4: stage_butterfly<4>(in, out, tw.data(), N, Ns, j); break;
        case 5: stage_butterfly<5>(in, out, tw.data(), N, Ns, j); break;
        case 7: stage_butterfly<7>(in, out, tw.data(), N, Ns, j); break;
        case 8: stage_butterfly<8>(in, out, tw.data(), N, Ns, j); break;
        case 11: stage_butterfly<11>(in, out, tw.data(), N, Ns, j); break;
        default: stage_butterfly<13>(in, out, tw.data(), N, Ns, j); break;
      }
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
extern "C" int b200sht_debug_fft_host(int N, int mmax, int direction, const float* in_a, const float* in_b, float* out_a,
                                      float* out_b) {
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
