// Longitude transform on the tensor cores (B200SHT_PREC_TF32): the truncated real DFT of every latitude row as
//   radix-8 butterflies + twiddles on the CUDA cores  x  one class-independent [M2 x N2/2] DFT matrix on tcgen05 (kind::tf32).
// Replaces the CUDA-core Stockham kernels of fft.cu for nlon = 8 * N2, N2 <= 190, mmax <= 256 (every shipped grid); see dft_math.cuh
// for the factorisation.  Reference semantics: 2 pi * torch.fft.rfft(x, norm="forward")[..., :mmax] and torch.fft.irfft(Z, n=nlon,
// norm="forward") inside torch_harmonics.RealSHT / InverseRealSHT (call sites makani/models/common/spectral_convolution.py:239,253).
//
// Why: the Stockham kernels were issue-bound on the CUDA cores (64.7 M warp instructions, 0.26-0.29 of HBM bandwidth):
// a B200 has ~37 TFLOP/s of fp32 add/mul against 6.5 TB/s, and a 1440-point row is 34 kflop for 4.8 KB.  Here two of the three
// radix stages (the 31 x 180 sub-transform, 86 % of the flops) run on the tensor pipe at TF32 and only one radix-8 stage stays on the
// CUDA cores; no shared-memory exchange between stages is left.
//
// synthesis kernel (latspec -> rows):   TMEM lane = column j2 (<= N2/2, replicated when N2/2 < 64 so that all four SM sub-partitions
//   work), accumulator columns = (class c, latitude k) of an 8-row tile.  A = E^T resident in shared memory (32 KB), B = the raw
//   latspec tile [m2][(c, k)] streamed by TMA (16 KB per 8 rows), four accumulators S1..S4 (cos/sin x re/im), double buffered.
//   Epilogue warps: tcgen05.ld -> V(j2), V(N2-j2) -> twiddle -> radix-8 -> scale/bias -> bf16; a warp stores 32 consecutive
//   longitudes of one row per instruction.
// analysis kernel (rows -> latspec):    producer warps load the eight samples x[N2 j1 + j2] of a column (lanes = consecutive j2),
//   butterfly + twiddle them and write the even/odd combinations (Ye, Yo) as K-major TF32 operand tiles [(c, k)][j2] (128-byte
//   swizzle, conflict-free row stores); B = E resident (<= 24 KB); D[(c,k)][m2] in TMEM; epilogue scales and writes latspec
//   (64-byte runs along k).
#include "umma_common.cuh"
#include "dft_math.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace b200sht {

int umma_available();   // umma.cu

constexpr int kDftMaxHalf = 95;    // N2 / 2 <= 95: three 32-lane quadrants (synthesis) / three K-blocks (analysis)
constexpr int kDftSynThreads = 512;
constexpr int kDftSynStages = 8;   // 16 KB each

struct DftTables {
  float* et;      // synthesis A: [2][128 rows = lane -> j2][32 m2]  (cos, sin), TF32-rounded
  float* eb;      // analysis  B: [nkb][2][32 rows m2][32 j2 local]  (cos, sin), TF32-rounded
  float2* tw;     // [8][N2]  exp(+2 pi i c j2 / nlon)
  float* trash;   // 64 x (4 * nlon + 256) floats: store target of the synthesis rows / lanes without output (keeps the stores unconditional); one slab per CTA % 64
  float* zeros;   // 8 * N2 floats of zeros: load target of the analysis lanes / rows that carry no sample (keeps the loads unconditional)
  int N2, half, M2, qpr, nrep, nkb;
};

bool dft_shape_ok(int nlon, int mmax) {
  if (nlon % 8 != 0) return false;
  const int N2 = nlon / 8;
  return N2 >= 2 && N2 / 2 <= kDftMaxHalf && (mmax + 7) / 8 <= 32 && mmax <= nlon / 2 + 1;
}

__global__ void dft_tables_kernel(float* et, float* eb, float2* tw, int N2, int half, int M2, int qpr, int nrep, int nkb, int nlon) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // E^T tiles: [2][128][32]
  if (i < 2 * 128 * 32) {
    const int m2 = i % 32, lane = (i / 32) % 128, p = i / (32 * 128);
    const int rep = lane / (32 * qpr), j2 = lane - rep * 32 * qpr;
    float v = 0.f;
    if (rep < nrep && j2 <= half && m2 < M2) {
      const long long t = ((long long)m2 * j2) % N2;
      const double ang = 2.0 * M_PI * (double)t / (double)N2;
      v = tf32_rn((float)(p == 0 ? cos(ang) : sin(ang)));
    }
    et[i] = v;
  }
  if (i < nkb * 2 * 32 * 32) {
    const int jl = i % 32, m2 = (i / 32) % 32, p = (i / 1024) % 2, kb = i / 2048;
    const int j2 = kb * 32 + jl;
    float v = 0.f;
    if (j2 <= half && m2 < M2) {
      const long long t = ((long long)m2 * j2) % N2;
      const double ang = 2.0 * M_PI * (double)t / (double)N2;
      v = tf32_rn((float)(p == 0 ? cos(ang) : sin(ang)));
      if (2 * j2 == N2) v *= 0.5f;   // the column N2 / 2 is its own partner: the producers count it twice (no mask)
    }
    eb[i] = v;
  }
  if (i < 8 * N2) {
    const int j2 = i % N2, c = i / N2;
    const double ang = 2.0 * M_PI * (double)(c * j2) / (double)nlon;
    tw[i] = make_float2((float)cos(ang), (float)sin(ang));
  }
}

int dft_plan_init(Plan* pl) {
  pl->dft_state = nullptr;
  if (!umma_available()) return -1;
  if (!dft_shape_ok(pl->nlon, pl->mmax)) return -1;
  DftTables* t = new DftTables();
  t->N2 = pl->nlon / 8; t->half = t->N2 / 2; t->M2 = (pl->mmax + 7) / 8;
  t->qpr = (t->half + 1 + 31) / 32;
  t->nrep = t->qpr == 1 ? 4 : (t->qpr == 2 ? 2 : 1);
  t->nkb = t->qpr;
  t->et = nullptr; t->eb = nullptr; t->tw = nullptr; t->zeros = nullptr; t->trash = nullptr;
  const size_t neb = (size_t)t->nkb * 2 * 32 * 32;
  cudaError_t e = cudaMalloc(&t->et, sizeof(float) * 2 * 128 * 32);
  if (e == cudaSuccess) e = cudaMalloc(&t->eb, sizeof(float) * neb);
  if (e == cudaSuccess) e = cudaMalloc(&t->tw, sizeof(float2) * 8 * t->N2);
  if (e == cudaSuccess) e = cudaMalloc(&t->zeros, sizeof(float) * 8 * t->N2);
  if (e == cudaSuccess) e = cudaMemset(t->zeros, 0, sizeof(float) * 8 * t->N2);
  if (e == cudaSuccess) e = cudaMalloc(&t->trash, sizeof(float) * 64 * (4 * (size_t)pl->nlon + 256));   // + 256: idle lanes index up to j2 = 127 + 7 N2 past a row
  if (e == cudaSuccess) {
    const int n = 8192 > 8 * t->N2 ? 8192 : 8 * t->N2;
    dft_tables_kernel<<<(n + 255) / 256, 256>>>(t->et, t->eb, t->tw, t->N2, t->half, t->M2, t->qpr, t->nrep, t->nkb, pl->nlon);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  }
  if (e != cudaSuccess) {
    cudaFree(t->et); cudaFree(t->eb); cudaFree(t->tw); cudaFree(t->zeros); cudaFree(t->trash);
    delete t;
    return -1;
  }
  pl->dft_state = t;
  return 0;
}

void dft_plan_destroy(Plan* pl) {
  DftTables* t = static_cast<DftTables*>(pl->dft_state);
  if (!t) return;
  cudaFree(t->et); cudaFree(t->eb); cudaFree(t->tw); cudaFree(t->zeros); cudaFree(t->trash);
  delete t;
  pl->dft_state = nullptr;
}

static bool dft_enabled() {
  static const int on = [] { const char* e = getenv("B200SHT_DFT"); return e ? atoi(e) : 1; }();
  return on != 0;
}
bool dft_usable(const Plan* pl) { return pl->dft_state != nullptr && dft_enabled(); }

// ----------------------------------------------------------------------------------------- wait-time profile
// B200SHT_DFT_PROF=1: every role accumulates the SM clocks it spends in its mbarrier waits (one atomic per wait, lane 0 of the warp) into 16
// counters, read back and cleared by b200sht_debug_dft_profile().  Slots -- analysis: 0 producers / raw samples, 1 producers / operand stage free,
// 2 loader / raw stage free, 3 MMA / operand stage full, 4 MMA / accumulator free, 5 epilogue / accumulator full, 6 CTA lifetime, 7 producer items;
// synthesis: 8 TMA / stage free, 9 MMA / stage full, 10 MMA / accumulator free, 11 epilogue / accumulator full, 12 CTA lifetime, 13 epilogue tiles.
static unsigned long long* g_dft_prof = nullptr;
static unsigned long long* dft_prof_buffer() {
  static const int on = [] { const char* e = getenv("B200SHT_DFT_PROF"); return e ? atoi(e) : 0; }();
  if (!on) return nullptr;
  if (!g_dft_prof) {
    if (cudaMalloc(&g_dft_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) { g_dft_prof = nullptr; return nullptr; }
    cudaMemset(g_dft_prof, 0, 16 * sizeof(unsigned long long));
  }
  return g_dft_prof;
}
int dft_profile_read(unsigned long long* out16) {
  if (!g_dft_prof) { for (int i = 0; i < 16; ++i) out16[i] = 0; return 0; }
  B200_CHECK_CUDA(cudaDeviceSynchronize());
  B200_CHECK_CUDA(cudaMemcpy(out16, g_dft_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  B200_CHECK_CUDA(cudaMemset(g_dft_prof, 0, 16 * sizeof(unsigned long long)));
  return 0;
}
#ifdef B200SHT_DFT_PROFILE
constexpr bool kDftProfile = true;
#else
constexpr bool kDftProfile = false;   // the counters cost a few instructions per wait: compiled in only by `python -m makani_b200.build --profile`
#endif
__device__ __forceinline__ void prof_wait(unsigned long long* prof, int slot, uint64_t* bar, uint32_t parity, bool lead) {
  if (!kDftProfile || prof == nullptr) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  if (lead) atomicAdd(prof + slot, (unsigned long long)(clock64() - t0));
}

// ------------------------------------------------------------------------------------------------ small PTX
__device__ __forceinline__ pr tmem_ld2(uint32_t taddr) {
  uint32_t a, b;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
  return make_pr(__uint_as_float(a), __uint_as_float(b));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <typename T> __device__ __forceinline__ void st_out(T* p, float v);
template <> __device__ __forceinline__ void st_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
// raw sample bits of the next item (converted when consumed): float bits, or the bf16 pattern in the low half
template <typename T> __device__ __forceinline__ uint32_t ld_raw(const T* p);
template <> __device__ __forceinline__ uint32_t ld_raw<float>(const float* p) { return __float_as_uint(__ldg(p)); }
template <> __device__ __forceinline__ uint32_t ld_raw<__nv_bfloat16>(const __nv_bfloat16* p) { return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)); }

// How the CUDA-core producers turn an fp32 value into a kind::tf32 operand (the MMA truncates the 13 low mantissa bits):
//   0  cvt.rna.tf32.f32                 3 instructions on sm_100a (FSETP + IADD + LOP3)
//   1  (bits + 0x1000) & ~0x1fff        2 instructions, same result for finite values
//   2  bias-compensated truncation      0 instructions: the value is pre-scaled by (1 + 2^-10 / 3) -- folded into the twiddle
//      factors -- so that the hardware truncation error x f - delta, delta ~ U[0, ulp), has zero mean over a binade; its rms is
//      0.304 ulp against 0.289 ulp for round-to-nearest, and TF32-exact inputs stay exact (x f < ulp / 3).
#ifndef B200_DFT_TF32_MODE
#define B200_DFT_TF32_MODE 2
#endif
constexpr float kTruncComp = 1.0f + 0.0009765625f / 3.0f;
__device__ __forceinline__ float tf32_operand(float v) {
#if B200_DFT_TF32_MODE == 0
  return tf32_rn(v);
#elif B200_DFT_TF32_MODE == 1
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
#else
  return v;   // already scaled through the twiddles
#endif
}
__device__ __forceinline__ float tf32_operand_unscaled(float v) {   // class 0 carries no twiddle
#if B200_DFT_TF32_MODE == 2
  return v * kTruncComp;
#else
  return tf32_operand(v);
#endif
}

__device__ __forceinline__ pr pr_operand(pr v) { return make_pr(tf32_operand(v.v.x), tf32_operand(v.v.y)); }
__device__ __forceinline__ pr pr_operand_unscaled(pr v) {
#if B200_DFT_TF32_MODE == 2
  return rmul(v, kTruncComp);
#else
  return pr_operand(v);
#endif
}

template <typename T> __device__ __forceinline__ float ld_in(const T* p);
template <> __device__ __forceinline__ float ld_in<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_in<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

// ================================================================================================ synthesis
struct DftSynParams {
  alignas(64) CUtensorMap tmZ;   // tiled latspec as ((c % 4, k % 8), c / 4, m2, p, tile), box (32, 1, 32, 1, 1): MN-major B operand, N = (c, k)
  alignas(64) CUtensorMap tmE;   // E^T tiles (m2, 256 rows), box (32, 128): K-major A operand
  const float* Z;
  void* y;
  const float2* tw;
  const float* rowscale;
  const float* bias;
  void* trash;
  unsigned long long* prof;
  int R, C, nlat, nlon, kp, mmax, N2, half, M2, qpr, nrep, mode, ntiles, ktiles, has_nyq;
  int kt0, kt_all;   // latitude range of this launch: first 8-row tile, tiles per image in the whole tensor (ktiles = tiles per image in the range)
  uint32_t idesc;
};

// shared memory: [A: cos 16 KB | sin 16 KB][B ring: kDftSynStages x 16 KB][tw table 8 x N2 float2][barriers]
template <typename T, int N2T>
__global__ void __launch_bounds__(kDftSynThreads, 1) dft_synthesis_kernel(const __grid_constant__ DftSynParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ unsigned long long prof_s[16];   // wait-time profile (B200SHT_DFT_PROF): accumulated per CTA, flushed once at the end
  unsigned long long* const prof = (kDftProfile && p.prof) ? prof_s : nullptr;
  if (kDftProfile && threadIdx.x < 16) prof_s[threadIdx.x] = 0;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + 32768;
  float2* tws = reinterpret_cast<float2*>(gbase + 32768 + kDftSynStages * 16384);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + 32768 + kDftSynStages * 16384 + ((8 * p.N2 * 8 + 15) & ~15));
  uint64_t* full = bars;
  uint64_t* empty = full + kDftSynStages;
  uint64_t* acc_full = empty + kDftSynStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* e_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_full + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  pdl_trigger();
  const int quad = warp & 3, sub = warp >> 2;
  const int subs = 4 / p.nrep;                                  // k pairs per replica
  const bool is_tma = (warp == 15), is_mma = (warp == 11);
  const bool is_epi = !is_tma && !is_mma && quad < p.qpr * p.nrep && sub < subs;
  const int n_epi = p.qpr * p.nrep * subs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDftSynStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], n_epi); }
    mbar_init(e_full, 1);
    fence_barrier_init();
    prefetch_tmap(&p.tmZ);
    prefetch_tmap(&p.tmE);
  }
  for (int i = threadIdx.x; i < 8 * p.N2; i += blockDim.x) tws[i] = p.tw[i];
  if (is_mma) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long t_cta0 = (kDftProfile && p.prof && threadIdx.x == 0) ? clock64() : 0;
  pdl_wait();   // the prologue read plan constants only (twiddles); the latspec tiles, bias and y belong to other kernels until here

  if (is_tma) {
    if (lane == 0) {
      mbar_expect_tx(e_full, 32768);
      tma_load_2d(sA, &p.tmE, e_full, 0, 0);
      tma_load_2d(sA + 16384, &p.tmE, e_full, 0, 128);
      int n = 0;
      for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
        const int s = n % kDftSynStages, it = n / kDftSynStages;
        if (it > 0) prof_wait(prof, 8, &empty[s], (it - 1) & 1, true);
        mbar_expect_tx(&full[s], 16384);
        const uint32_t st = sB + s * 16384;
        const int ta = (ti / p.ktiles) * p.kt_all + p.kt0 + ti % p.ktiles;   // tile index in the whole tensor
        tma_load_5d(st, &p.tmZ, &full[s], 0, 0, 0, 0, ta);           // re, classes 0..3
        tma_load_5d(st + 4096, &p.tmZ, &full[s], 0, 1, 0, 0, ta);    // re, classes 4..7
        tma_load_5d(st + 8192, &p.tmZ, &full[s], 0, 0, 0, 1, ta);    // im
        tma_load_5d(st + 12288, &p.tmZ, &full[s], 0, 1, 0, 1, ta);
      }
    }
    __syncwarp();
  } else if (is_mma) {
    {   // all 32 lanes run the loop (converged); the MMAs / commits are issued by an elected lane (umma_*_ws)
      mbar_wait(e_full, 0);
      const uint64_t dE0 = desc_kmajor(sA, 0), dZ0 = desc_mnmajor(sB, 0, 4096);
      int n = 0;
      for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
        const int s = n % kDftSynStages, it = n / kDftSynStages;
        const int buf = n & 1, use = n >> 1;
        if (use > 0) { prof_wait(prof, 10, &acc_empty[buf], (use - 1) & 1, lane == 0); }
        prof_wait(prof, 9, &full[s], it & 1, lane == 0);
        tc_fence_after();
        const uint64_t z0 = desc_advance(dZ0, (uint32_t)s * 16384u);
        const uint32_t d = tmem + buf * 256;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t ac = desc_advance(dE0, 32 * j), as = desc_advance(dE0, 16384 + 32 * j);
          const uint64_t zr = desc_advance(z0, 1024 * j), zi = desc_advance(z0, 8192 + 1024 * j);
          const uint32_t acc = j > 0 ? 1u : 0u;
          umma_tf32_ws(d, ac, zr, p.idesc, acc);          // S1 = cos . Zr
          umma_tf32_ws(d + 64, as, zi, p.idesc, acc);     // S2 = sin . Zi
          umma_tf32_ws(d + 128, as, zr, p.idesc, acc);    // S3 = sin . Zr
          umma_tf32_ws(d + 192, ac, zi, p.idesc, acc);    // S4 = cos . Zi
        }
        umma_commit_ws(&empty[s]);
        umma_commit_ws(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (is_epi) {
    const int N2 = N2T > 0 ? N2T : p.N2;
    const int nlon = 8 * N2;
    const int rep = quad / p.qpr;
    const int j2 = 32 * (quad - rep * p.qpr) + lane;
    const bool valid = j2 <= p.half;
    const bool paired = valid && j2 != 0 && 2 * j2 != N2;
    const int jp = N2 - j2;
    const int kpi = rep * subs + sub;
    const bool n2odd = (N2 & 1) != 0;
    T* const y = static_cast<T*>(p.y);
    const float smul = p.mode == 0 ? 2.f : 1.f;
    const int nyq_m = nlon / 2;
    T* const trash = static_cast<T*>(p.trash) + (size_t)(blockIdx.x & 63) * (4 * nlon + 256);
    float2 tw[8], tp[8];
    tw[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int c = 1; c < 8; ++c) tw[c] = valid ? tws[c * N2 + j2] : make_float2(1.f, 0.f);
    dft_partner_twiddles(tw, tp);
    int n = 0;
    for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
      const int r = ti / p.ktiles, k0 = (p.kt0 + ti - r * p.ktiles) * 8;
      const int ta = r * p.kt_all + p.kt0 + (ti - r * p.ktiles);   // tile index in the whole tensor
      const int ka = k0 + 2 * kpi;
      const int buf = n & 1, use = n >> 1;
      // per-row output factors:  out = x * sc + off(parity of the longitude)
      float rsa = 1.f, rsb = 1.f, z0a = 0.f, z0b = 0.f, zna = 0.f, znb = 0.f;
      if (p.mode == 1) {
        const float2 rs = *reinterpret_cast<const float2*>(p.rowscale + ka);
        rsa = rs.x; rsb = rs.y;
      } else {
        // tiled latspec: element (m, plane, r, k) at ((tile * 2 + plane) * M2 + m / 8) * 64 + (m % 8) * 8 + k % 8, tile = r * ktiles + k / 8
        const float* zt = p.Z + (size_t)ta * 2 * p.M2 * 64 + 2 * kpi;
        const float2 z0 = *reinterpret_cast<const float2*>(zt);
        z0a = z0.x; z0b = z0.y;
        if (p.has_nyq) {
          const float2 zn = *reinterpret_cast<const float2*>(zt + (nyq_m >> 3) * 64 + (nyq_m & 7) * 8);
          zna = zn.x; znb = zn.y;
        }
      }
      const float bias = p.bias ? __ldg(p.bias + r % p.C) : 0.f;
      const pr sc = make_pr(smul * rsa, smul * rsb);
      const pr off_e = make_pr(bias - rsa * (z0a + zna), bias - rsb * (z0b + znb));   // even longitude j
      const pr off_o = make_pr(bias - rsa * (z0a - zna), bias - rsb * (z0b - znb));   // odd longitude j
      prof_wait(prof, 11, &acc_full[buf], use & 1, lane == 0);
      if (prof && lane == 0) atomicAdd(prof + 13, 1ull);
      tc_fence_after();
      const uint32_t t0 = tmem + ((uint32_t)(quad * 32) << 16) + buf * 256 + 2 * kpi;
      // rows beyond nlat (last tile of an image) are stored into a scratch row: no predicates / branches around the 32 stores
      T* const pa = (valid && ka < p.nlat) ? y + ((size_t)r * p.nlat + ka) * nlon + j2 : trash + j2;
      T* const pb = (valid && ka + 1 < p.nlat) ? y + ((size_t)r * p.nlat + ka + 1) * nlon + j2 : trash + nlon + j2;
      const int dq = paired ? jp - j2 : 0;
      // The four accumulators are read twice (TMEM reads are cheap) and reduced to V of one column right away, so that only 16 register
      // pairs are live through each radix-8 pass: holding S1..S4 (64 registers) next to the twiddles made ptxas rematerialise the
      // 64-bit store address for every one of the 32 stores (10 instructions each).
      {
        pr vr[8], vi[8], x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const pr a1 = tmem_ld2(t0 + c * 8), a2 = tmem_ld2(t0 + 64 + c * 8), a3 = tmem_ld2(t0 + 128 + c * 8), a4 = tmem_ld2(t0 + 192 + c * 8);
          tmem_ld_wait();
          vr[c] = a1 - a2;
          vi[c] = a3 + a4;
        }
        dft_syn_radix8<pr>(vr, vi, tw, x);
        const pr o0 = (j2 & 1) ? off_o : off_e, o1 = (j2 & 1) ? off_e : off_o;
#pragma unroll
        for (int j1 = 0; j1 < 8; ++j1) {
          const pr o = rfma(x[j1], sc, (n2odd && (j1 & 1)) ? o1 : o0);
          st_out<T>(pa + N2 * j1, o.v.x);
          st_out<T>(pb + N2 * j1, o.v.y);
        }
      }
      {
        pr vr[8], vi[8], x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const pr a1 = tmem_ld2(t0 + c * 8), a2 = tmem_ld2(t0 + 64 + c * 8), a3 = tmem_ld2(t0 + 128 + c * 8), a4 = tmem_ld2(t0 + 192 + c * 8);
          tmem_ld_wait();
          vr[c] = a1 + a2;
          vi[c] = a4 - a3;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);   // both reads done: release the accumulator set
        dft_syn_radix8<pr>(vr, vi, tp, x);
        T* const qa = paired ? pa + dq : trash + 2 * nlon + j2;     // unpaired columns (j2 = 0, N2 / 2) and idle lanes: scratch
        T* const qb = paired ? pb + dq : trash + 3 * nlon + j2;
        const pr o0 = (jp & 1) ? off_o : off_e, o1 = (jp & 1) ? off_e : off_o;
#pragma unroll
        for (int j1 = 0; j1 < 8; ++j1) {
          const pr o = rfma(x[j1], sc, (n2odd && (j1 & 1)) ? o1 : o0);
          st_out<T>(qa + N2 * j1, o.v.x);
          st_out<T>(qb + N2 * j1, o.v.y);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kDftProfile && p.prof && threadIdx.x == 0) prof_s[12] = (unsigned long long)(clock64() - t_cta0);
  if (kDftProfile) __syncthreads();
  if (kDftProfile && p.prof && threadIdx.x < 16) atomicAdd(p.prof + threadIdx.x, prof_s[threadIdx.x]);
  if (is_mma) tmem_dealloc(tmem, 512);
}

// k_begin / k_end: latitude range [k_begin, k_end) to produce (k_begin a multiple of 8; k_end < 0: all rows) -- the other rows of y are not touched
int dft_synthesis(const Plan* pl, const float* Z, void* y, int dtype, int B, int C, const float* bias, int mode, cudaStream_t st, int k_begin, int k_end) {
  const DftTables* t = static_cast<const DftTables*>(pl->dft_state);
  B200_REQUIRE(t != nullptr, "dft_synthesis: plan has no DFT tables");
  const int R = B * C;
  DftSynParams p;
  memset(&p, 0, sizeof(p));
  p.Z = Z; p.y = y; p.tw = t->tw; p.rowscale = pl->d_rowscale; p.bias = bias; p.trash = t->trash; p.prof = dft_prof_buffer();
  p.R = R; p.C = C; p.nlat = pl->nlat; p.nlon = pl->nlon; p.kp = pl->kp; p.mmax = pl->mmax;
  p.N2 = t->N2; p.half = t->half; p.qpr = t->qpr; p.nrep = t->nrep; p.mode = mode;
  if (k_end < 0 || k_end > pl->kp) k_end = pl->kp;
  B200_REQUIRE(k_begin >= 0 && k_begin % 8 == 0 && k_begin < k_end, "dft_synthesis: bad latitude range [%d, %d)", k_begin, k_end);
  p.kt_all = pl->kp / 8; p.kt0 = k_begin / 8;
  p.ktiles = (k_end - k_begin + 7) / 8; p.ntiles = R * p.ktiles;
  p.has_nyq = (pl->mmax == pl->nlon / 2 + 1) ? 1 : 0;
  p.idesc = make_idesc(64, 0, 1, 0);
  p.M2 = t->M2;
  {
    // tiled latspec (written by legendre_synthesis_umma(tiled = 1)): [tile = r * ktiles + k / 8][plane][m2][c = m % 8][k % 8]; a tile is 16 KB
    // contiguous, the 128-byte rows of the TMA box are (4 classes x 8 latitudes) of one m2: the MN-major B operand, N = (c, k)
    long long d[5] = {32, 2, t->M2, 2, (long long)R * (pl->kp / 8)}, s[5] = {1, 32, 64, (long long)t->M2 * 64, 2ll * t->M2 * 64};
    int bx[5] = {32, 1, 32, 1, 1};
    int rc = make_tmap(&p.tmZ, Z, 5, d, s, bx, true);
    if (rc) return rc;
  }
  {
    long long d[2] = {32, 256}, s[2] = {1, 32};
    int bx[2] = {32, 128};
    int rc = make_tmap(&p.tmE, t->et, 2, d, s, bx);
    if (rc) return rc;
  }
  const size_t smem = 1024 + 32768 + (size_t)kDftSynStages * 16384 + ((8 * (size_t)t->N2 * 8 + 15) & ~(size_t)15) + (2 * kDftSynStages + 5) * 8 + 16;
  const int sms = usable_sms(pl->sm_count > 0 ? pl->sm_count : 148);
  const int ctas = p.ntiles < sms ? p.ntiles : sms;
#define B200_LAUNCH_SYN(TT, NN)                                                                                                          \
  do {                                                                                                                                  \
    struct Tag {};                                                                                                                        \
    B200_CHECK_CUDA((ensure_dynamic_smem<Tag>(dft_synthesis_kernel<TT, NN>, smem)));                                                     \
    B200_CHECK_CUDA(launch_pdl(dft_synthesis_kernel<TT, NN>, dim3(ctas), dim3(kDftSynThreads), smem, st, p));                                                                 \
  } while (0)
#define B200_DISPATCH_SYN(TT)                                            \
  switch (t->N2) {                                                       \
    case 180: B200_LAUNCH_SYN(TT, 180); break; /* nlon 1440 */           \
    case 90: B200_LAUNCH_SYN(TT, 90); break;   /* nlon  720 */           \
    case 60: B200_LAUNCH_SYN(TT, 60); break;   /* nlon  480 */           \
    default: B200_LAUNCH_SYN(TT, 0); break;                              \
  }
  if (dtype == B200SHT_BF16) { B200_DISPATCH_SYN(__nv_bfloat16) } else { B200_DISPATCH_SYN(float) }
#undef B200_DISPATCH_SYN
#undef B200_LAUNCH_SYN
  B200_CHECK_LAUNCH();
  return 0;
}

// ================================================================================================= analysis
__host__ __device__ constexpr int dft_box_group(int N2, int es) { return (N2 * es) % 16 == 0 ? 1 : (2 * N2 * es) % 16 == 0 ? 2 : (4 * N2 * es) % 16 == 0 ? 4 : 8; }

// 3-D view of the samples for the analysis loader: (column inside a group of gs row segments, group, row), box (box_cols, 8 / gs, 16), no swizzle
static int make_tmap_segments(CUtensorMap* tm, const void* base, bool bf16, int nlon, int gs, long long rows, int box_cols) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = 3; key.kind = bf16 ? 5 : 4;
  key.dims[0] = nlon; key.dims[1] = gs; key.dims[2] = rows; key.box[0] = box_cols;
  int slot = 0;
  if (tmap_lookup(key, tm, &slot)) return 0;
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return B200SHT_ERR_UNSUPPORTED; }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  const int es = bf16 ? 2 : 4, N2 = nlon / 8;
  cuuint64_t gd[3] = {(cuuint64_t)gs * N2, (cuuint64_t)(8 / gs), (cuuint64_t)rows};
  cuuint64_t gst[2] = {(cuuint64_t)gs * N2 * es, (cuuint64_t)nlon * es};
  cuuint32_t bx[3] = {(cuuint32_t)box_cols, (cuuint32_t)(8 / gs), 16}, el[3] = {1, 1, 1};
  if (gst[0] % 16 != 0 || gst[1] % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (box_cols * es) % 16 != 0) {
    set_error("tensor map (segments): base / pitch / box row not 16-byte aligned");
    return B200SHT_ERR_INVALID;
  }
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gd, gst, bx, el,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (segments) failed (%d)", (int)r); return B200SHT_ERR_CUDA; }
  tmap_store(key, tm, slot);
  return 0;
}

constexpr int kDftAnaStages = 2;   // operand ring: one stage = one K-block (32 columns) of a 16-row tile = 4 planes x 16 KB

struct DftAnaParams {
  alignas(64) CUtensorMap tmB;   // E tiles (32 j2 local, nkb * 64 rows), box (32, 32): K-major B operand
  alignas(64) CUtensorMap tmXc;  // the input as [R * nlat rows][8 / gs groups][gs * N2], box (Wc columns, 8 / gs, 16 rows), no swizzle: columns j2
  alignas(64) CUtensorMap tmXp;  // same, box (Wp columns, 8 / gs, 16 rows): partner columns N2 - j2
  float* X;
  const float2* tw;
  const float* rowscale;
  unsigned long long* prof;
  int R, nlat, nlon, kp, mmax, N2, half, M2, nkb, mode, round_tf32, ntiles, ktiles, nraw, gs;
  int kt0;   // first 16-row tile of the latitude range this launch transforms (ktiles = tiles in the range; latitude-chunked analysis, capi.cu)
  uint32_t idesc, idesc_neg;
};

// warps: 0..3 epilogue (TMEM quadrant = warp), 4 MMA issuer (+ TMEM owner, loads the resident B), 5 sample loader (TMA), 6.. producers
// shared memory: [B resident: nkb x (cos 4 KB | sin 4 KB)][A ring: 2 x 4 planes x 16 KB][raw ring: nraw x 16 boxes][twiddles][barriers]
//
// Data flow of one (tile, K-block): the loader thread brings the 8 + 8 sample boxes the K-block needs -- for each j1 the 32 columns
// j2 = 32 kb .. + 31 and their 32 partner columns N2 - j2, 16 rows each -- into a raw stage with 16 TMA boxes (deep asynchronous prefetch,
// no registers: the first version's LDG -> register path stalled 2.1 cycles per issued instruction on the loads, with the 96-register cap
// allowing only half an item of prefetch).  Eight producer warps (one row pair each) read their samples with LDS, run the two radix-8
// butterflies + twiddles with the two rows packed in f32x2, and write Ye / Yo of the 8 classes into the operand stage; then the MMA thread
// contracts the stage with E.  N2T > 0: nlon / 8 as a compile-time constant.
template <typename T, int N2T>
__global__ void __launch_bounds__(576, 1) dft_analysis_kernel(const __grid_constant__ DftAnaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ unsigned long long prof_s[16];   // wait-time profile (B200SHT_DFT_PROF): accumulated per CTA, flushed once at the end
  unsigned long long* const prof = (kDftProfile && p.prof) ? prof_s : nullptr;
  if (kDftProfile && threadIdx.x < 16) prof_s[threadIdx.x] = 0;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  // TMA needs 16-byte aligned box starts (found on the GPU: an unaligned inner coordinate is an illegal instruction): a box starts at the
  // column rounded down to kAl elements and is wide enough to still contain the 32 wanted ones.  fp32 column boxes are exact (N2 % 4 == 0 is
  // required by the host for fp32 input).
  constexpr int kAl = 16 / (int)sizeof(T);                       // 8 (bf16) / 4 (fp32)
  constexpr int kWc = (sizeof(T) == 2) ? 40 : 32, kWp = (sizeof(T) == 2) ? 40 : 36;
  constexpr uint32_t kColBytes = 16u * kWc * sizeof(T), kParBytes = 16u * kWp * sizeof(T);   // one box: 16 rows
  constexpr uint32_t kRawBytes = 8u * (kColBytes + kParBytes);
  const uint32_t oB = 0, oA = 3 * 8192, oR = oA + kDftAnaStages * 65536, oT = oR + (uint32_t)p.nraw * kRawBytes, oBar = oT + 3 * 7 * 32 * 8;
  const uint32_t sBm = base + oB, sAr = base + oA, sRaw = base + oR;
  uint8_t* gA = gbase + oA;
  const uint8_t* gR = gbase + oR;
  float2* twS = reinterpret_cast<float2*>(gbase + oT);   // [nkb][7][32] twiddles of the producer lanes
  uint64_t* full = reinterpret_cast<uint64_t*>(gbase + oBar);
  uint64_t* empty = full + kDftAnaStages;
  uint64_t* acc_full = empty + kDftAnaStages;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* raw_full = acc_empty + 4;
  uint64_t* raw_empty = raw_full + 4;
  uint64_t* b_full = raw_empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  pdl_trigger();
  const int nkb = p.nkb;
  const int N2 = N2T > 0 ? N2T : p.N2;
  // row segments per sample box group: the smallest gs with (gs * N2 elements) a multiple of 16 bytes, so that the segments gm, gm + gs, ...
  // of all rows form one 3-D TMA box (2 gs boxes per K-block instead of 16)
  const int gs = N2T > 0 ? dft_box_group(N2T, (int)sizeof(T)) : p.gs;
  for (int i = threadIdx.x; i < nkb * 7 * 32; i += blockDim.x) {
    const int ln = i & 31, c = (i >> 5) % 7 + 1, kb = i / 224;
    const int j2 = 32 * kb + ln;
    float2 w = (j2 <= p.half) ? p.tw[c * N2 + j2] : make_float2(1.f, 0.f);
#if B200_DFT_TF32_MODE == 2
    w.x *= kTruncComp; w.y *= kTruncComp;
#endif
    twS[i] = w;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDftAnaStages; ++s) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); }   // a K-block = 8 row pairs, one arrival each
    for (int b = 0; b < 4; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    for (int s = 0; s < p.nraw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 8); }
    mbar_init(b_full, 1);
    fence_barrier_init();
    prefetch_tmap(&p.tmB);
    prefetch_tmap(&p.tmXc);
    prefetch_tmap(&p.tmXp);
  }
  // imaginary parts of class 0 (rows 0..15 of the planes Ye_i, Yo_i of every operand stage): zero, never written again
  for (int i = threadIdx.x; i < kDftAnaStages * 2 * 512; i += blockDim.x) {
    const int st = i >> 10, pl = (i >> 9) & 1, off = i & 511;
    reinterpret_cast<float*>(gA + (size_t)st * 65536)[(pl ? 12288 : 4096) + off] = 0.f;
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long t_cta0 = (kDftProfile && p.prof && threadIdx.x == 0) ? clock64() : 0;
  pdl_wait();   // the prologue read plan constants only (twiddles); samples and latspec belong to other kernels until here

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(b_full, (uint32_t)nkb * 8192);
      for (int kb = 0; kb < nkb; ++kb) {
        tma_load_2d(sBm + kb * 8192, &p.tmB, b_full, 0, kb * 64);
        tma_load_2d(sBm + kb * 8192 + 4096, &p.tmB, b_full, 0, kb * 64 + 32);
      }
    }
    __syncwarp();
    {   // all 32 lanes run the loop (converged); the MMAs / commits are issued by an elected lane (umma_*_ws)
      mbar_wait(b_full, 0);
      const uint64_t dA0 = desc_kmajor(sAr, 0), dB0 = desc_kmajor(sBm, 0);   // every other descriptor = one of these + a byte offset
      int n = 0;
      for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
        const int buf = n & 3, use = n >> 2;
        if (use > 0) prof_wait(prof, 4, &acc_empty[buf], (use - 1) & 1, lane == 0);
        tc_fence_after();
        const uint32_t d = tmem + buf * 64;
        for (int kb = 0; kb < nkb; ++kb) {
          const int g = n * nkb + kb, s = g % kDftAnaStages, it = g / kDftAnaStages;
          prof_wait(prof, 3, &full[s], it & 1, lane == 0);   // precise wake-up: the stage is released (empty) only after these MMAs
          tc_fence_after();
          const uint64_t a0 = desc_advance(dA0, (uint32_t)s * 65536u);
          const uint64_t bc = desc_advance(dB0, (uint32_t)kb * 8192u), bs = desc_advance(bc, 4096);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
            const uint64_t bcj = desc_advance(bc, 32 * j), bsj = desc_advance(bs, 32 * j);
            umma_tf32_ws(d, desc_advance(a0, 32 * j), bcj, p.idesc, acc);                       // Xre  = Ye_r cos
            umma_tf32_ws(d, desc_advance(a0, 49152 + 32 * j), bsj, p.idesc, 1u);               // Xre += Yo_i sin
            umma_tf32_ws(d + 32, desc_advance(a0, 16384 + 32 * j), bcj, p.idesc, acc);         // Xim  = Ye_i cos
            umma_tf32_ws(d + 32, desc_advance(a0, 32768 + 32 * j), bsj, p.idesc_neg, 1u);      // Xim -= Yo_r sin
          }
          umma_commit_ws(&empty[s]);
        }
        umma_commit_ws(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------------------------------- sample loader
    if (lane == 0) {
      int n = 0;
      for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
        const int r = ti / p.ktiles, row0 = r * p.nlat + (p.kt0 + ti - r * p.ktiles) * 16;
        for (int kb = 0; kb < nkb; ++kb) {
          const int g = n * nkb + kb, rs = g % p.nraw, it = g / p.nraw;
          if (it > 0) prof_wait(prof, 2, &raw_empty[rs], (it - 1) & 1, true);
          mbar_expect_tx(&raw_full[rs], kRawBytes);
          const uint32_t dst = sRaw + rs * kRawBytes;
#pragma unroll
          for (int gm = 0; gm < 8; ++gm) {
            if (gm >= gs) break;
            // one box = the columns of the row segments j1 = gm, gm + gs, ...: (kW columns) x (8 / gs segments) x (16 rows)
            const int sc = N2 * gm + 32 * kb, sp = N2 * gm + N2 - 32 * kb - 31;   // first wanted column / partner column (sp < 0 only for N2 < 31)
            tma_load_3d(dst + gm * (8 / gs) * kColBytes, &p.tmXc, &raw_full[rs], (sc / kAl) * kAl, 0, row0);
            tma_load_3d(dst + 8 * kColBytes + gm * (8 / gs) * kParBytes, &p.tmXp, &raw_full[rs], sp < 0 ? 0 : (sp / kAl) * kAl, 0, row0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ------------------------------------------------------------------------------------------- epilogue
    const int c = 2 * warp + (lane >> 4), kr = lane & 15;
    const size_t plane = (size_t)p.R * p.kp;
    int n = 0;
    for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
      const int r = ti / p.ktiles, k0 = (p.kt0 + ti - r * p.ktiles) * 16;
      const int k = k0 + kr;
      const int buf = n & 3, use = n >> 2;
      const bool kok = k < p.kp;
      const float rs = (p.mode == 0) ? ((k < p.nlat) ? __ldg(p.rowscale + k) : 0.f) : 1.f;
      const float tcomp = p.round_tf32 ? kTruncComp : 1.f;
      {
        const long long tw0 = prof ? clock64() : 0;
        mbar_wait_relaxed(&acc_full[buf], use & 1, 1000);
        if (prof && lane == 0) atomicAdd(prof + 5, (unsigned long long)(clock64() - tw0));
      }
      tc_fence_after();
      float vr[32], vi[32];
      const uint32_t t0 = tmem + ((uint32_t)(warp * 32) << 16) + buf * 64;
      tmem_ld32_nowait(t0, vr);
      tmem_ld32_nowait(t0 + 32, vi);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (!kok) continue;
      float* xb = p.X + (size_t)r * p.kp + k;
#pragma unroll
      for (int m2 = 0; m2 < 32; ++m2) {
        const int m = c + 8 * m2;
        if (m >= p.mmax) break;
        // round_tf32: the consumer is the kind::tf32 Legendre GEMM, which truncates its operands -> bias-compensated truncation folded
        // into the scale factor (see B200_DFT_TF32_MODE above) instead of 3 instructions of cvt.rna per value
        const float sc = ((p.mode == 0) ? rs : ((m == 0 || 2 * m == p.nlon) ? 1.f : 2.f)) * tcomp;
        const float a = vr[m2] * sc, b = vi[m2] * sc;
        float* dst = xb + (size_t)m * 2 * plane;
        dst[0] = a;
        dst[plane] = b;
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------- producers
    // Work item = (K-block kb, row pair q): rows 2q, 2q + 1 of the tile in the halves of packed f32x2 registers, lanes = the 32 columns of the
    // K-block.  Items are taken in K-block-major order (item = kb * 8 + q; warp w does w, w + nprod, ...): the MMAs of K-block kb run while
    // the warps work on kb + 1.  12 producer warps for three K-blocks (2 items per warp and tile), 8 otherwise: the kernel is bound by the
    // latency of the dependent butterfly chains, so the tile time is (items per warp) x (item latency).
    const int pw = warp - 6;
    const int nprod = (int)(blockDim.x >> 5) - 6;
    const int ipw = (8 * nkb) / nprod;
    constexpr bool kBf16 = (sizeof(T) == 2);
    const T* const rawS = reinterpret_cast<const T*>(gR);
    int n = 0;
    for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, ++n) {
      const int r = ti / p.ktiles, kt16 = (p.kt0 + ti - r * p.ktiles) * 16;
      for (int ii = 0; ii < ipw; ++ii) {
        const int item = pw + ii * nprod;
        const int kb = item >> 3, q = item & 7;
        const int k0 = kt16 + 2 * q;
        const int g = n * nkb + kb;
        const int j2 = 32 * kb + lane;
        const int rs = g % p.nraw;
        prof_wait(prof, 0, &raw_full[rs], (g / p.nraw) & 1, lane == 0);
        if (prof && lane == 0) atomicAdd(prof + 7, 1ull);
        const T* const rb = rawS + (size_t)rs * (kRawBytes / sizeof(T));
        pr xa[8], xb[8];
        // No per-lane masks on the samples: lanes beyond N2 / 2 feed rows of E that are zero, and the column N2 / 2 (its own partner) is
        // simply counted twice against a halved row of E; only column 0 (no partner) and rows beyond nlat are patched below, in branches
        // that are uniform (and rarely taken).
#pragma unroll
        for (int j1 = 0; j1 < 8; ++j1) {
          // column box gm: [16 rows][npb segments][kWc]; partner box: [16 rows][npb][kWp]; both start at the wanted column rounded down to kAl
          const int gm = j1 % gs, ga = j1 / gs, npb = 8 / gs;                  // box gm, segment ga of its npb segments
          const int sc = N2 * gm + 32 * kb, sp = N2 * gm + N2 - 32 * kb - 31;
          const int ic = sc - (sc / kAl) * kAl + lane;                          // column j2 = 32 kb + lane
          int ip = N2 * gm + N2 - j2 - (sp < 0 ? 0 : (sp / kAl) * kAl);          // column N2 - j2
          ip = ip < 0 ? 0 : (ip > kWp - 1 ? kWp - 1 : ip);                      // lanes without a partner read anything inside the box
          const T* b0 = rb + gm * (16 * npb * kWc) + (2 * q) * (npb * kWc) + ga * kWc;
          const T* b1 = rb + 8 * (16 * kWc) + gm * (16 * npb * kWp) + (2 * q) * (npb * kWp) + ga * kWp;
          const int rowc = npb * kWc, rowp = npb * kWp;                         // row pitch inside a box
          if constexpr (kBf16) {
            xa[j1] = make_pr(__uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(b0)[ic] << 16),
                             __uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(b0)[rowc + ic] << 16));
            xb[j1] = make_pr(__uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(b1)[ip] << 16),
                             __uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(b1)[rowp + ip] << 16));
          } else {
            xa[j1] = make_pr(reinterpret_cast<const float*>(b0)[ic], reinterpret_cast<const float*>(b0)[rowc + ic]);
            xb[j1] = make_pr(reinterpret_cast<const float*>(b1)[ip], reinterpret_cast<const float*>(b1)[rowp + ip]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);   // samples are in registers: the raw stage may be refilled
        if (kb == 0) {                                 // column 0 has no partner
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) xb[j1] = (lane == 0) ? make_pr(0.f, 0.f) : xb[j1];
        }
        if (k0 + 1 >= p.nlat) {                        // rows beyond nlat belong to the next image (or are out of bounds): zeros
          const bool row0ok = k0 < p.nlat;
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) {
            xa[j1] = make_pr(row0ok ? xa[j1].v.x : 0.f, 0.f);
            xb[j1] = make_pr(row0ok ? xb[j1].v.x : 0.f, 0.f);
          }
        }
        pr er[8], ei[8], br[8], bi[8];
        float2 tw[8], tp[8];
        tw[0] = make_float2(1.f, 0.f);
        const float2* const twl = twS + kb * 224 + lane;   // tw[c] = twl[(c - 1) * 32], already scaled for the truncation compensation
#pragma unroll
        for (int c = 1; c < 8; ++c) tw[c] = twl[(c - 1) * 32];
        dft_partner_twiddles(tw, tp);   // products of the tw components with constants of modulus 1: they carry the (1 + f) factor too
        dft_ana_radix8<pr>(xa, tw, er, ei);
        dft_ana_radix8<pr>(xb, tp, br, bi);
        const int s = g % kDftAnaStages, it = g / kDftAnaStages;
        if (it > 0) prof_wait(prof, 1, &empty[s], (it - 1) & 1, lane == 0);
        float* const stg = reinterpret_cast<float*>(gA + (size_t)s * 65536);
        const int kr0 = 2 * q;
        // swizzled K-major position of (row c * 16 + kr, column lane): the XOR term depends on kr only (16 c is a multiple of 8)
        float* const d0 = stg + kr0 * 32 + ((((lane >> 2) ^ (kr0 & 7)) << 2) | (lane & 3));
        float* const d1 = stg + (kr0 + 1) * 32 + ((((lane >> 2) ^ ((kr0 + 1) & 7)) << 2) | (lane & 3));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          pr ye_r = er[c] + br[c], yo_r = er[c] - br[c];
          if (c == 0) { ye_r = pr_operand_unscaled(ye_r); yo_r = pr_operand_unscaled(yo_r); }
          else { ye_r = pr_operand(ye_r); yo_r = pr_operand(yo_r); }
          d0[c * 512] = ye_r.v.x; d1[c * 512] = ye_r.v.y;
          d0[8192 + c * 512] = yo_r.v.x; d1[8192 + c * 512] = yo_r.v.y;
          if (c != 0) {   // the imaginary parts of class 0 are zero: those 16 rows of the two planes are cleared once at kernel start
            const pr ye_i = pr_operand(ei[c] + bi[c]), yo_i = pr_operand(ei[c] - bi[c]);
            d0[4096 + c * 512] = ye_i.v.x; d1[4096 + c * 512] = ye_i.v.y;
            d0[12288 + c * 512] = yo_i.v.x; d1[12288 + c * 512] = yo_i.v.y;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kDftProfile && p.prof && threadIdx.x == 0) prof_s[6] = (unsigned long long)(clock64() - t_cta0);
  if (kDftProfile) __syncthreads();
  if (kDftProfile && p.prof && threadIdx.x < 16) atomicAdd(p.prof + threadIdx.x, prof_s[threadIdx.x]);
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// k_begin / k_end: latitude range [k_begin, k_end) to transform (k_begin a multiple of 16; k_end < 0: up to kp) -- the other rows of X are not touched
int dft_analysis(const Plan* pl, const void* x, int dtype, int B, int C, float* X, int mode, int round_tf32, cudaStream_t st, int k_begin, int k_end) {
  const DftTables* t = static_cast<const DftTables*>(pl->dft_state);
  B200_REQUIRE(t != nullptr, "dft_analysis: plan has no DFT tables");
  const int R = B * C;
  DftAnaParams p;
  memset(&p, 0, sizeof(p));
  p.X = X; p.tw = t->tw; p.rowscale = pl->d_rowscale; p.prof = dft_prof_buffer();
  p.R = R; p.nlat = pl->nlat; p.nlon = pl->nlon; p.kp = pl->kp; p.mmax = pl->mmax;
  p.N2 = t->N2; p.half = t->half; p.M2 = t->M2; p.nkb = t->nkb; p.mode = mode; p.round_tf32 = round_tf32;
  if (k_end < 0 || k_end > pl->kp) k_end = pl->kp;
  B200_REQUIRE(k_begin >= 0 && k_begin % 16 == 0 && k_begin < k_end, "dft_analysis: bad latitude range [%d, %d)", k_begin, k_end);
  p.kt0 = k_begin / 16;
  p.ktiles = (k_end - k_begin + 15) / 16; p.ntiles = R * p.ktiles;
  const bool bf16 = (dtype == B200SHT_BF16);
  p.nraw = bf16 ? 3 : 2;   // raw stages of 20 / 34 KB
  p.idesc = make_idesc(32, 0, 0, 0);
  p.idesc_neg = make_idesc(32, 0, 0, 1);
  {
    long long d[2] = {32, (long long)t->nkb * 64}, s[2] = {1, 32};
    int bx[2] = {32, 32};
    int rc = make_tmap(&p.tmB, t->eb, 2, d, s, bx);
    if (rc) return rc;
  }
  B200_REQUIRE(bf16 || t->N2 % 4 == 0, "dft_analysis: fp32 input needs nlon %% 32 == 0 (16-byte aligned TMA boxes)");
  {
    p.gs = dft_box_group(t->N2, bf16 ? 2 : 4);
    int rc = make_tmap_segments(&p.tmXc, x, bf16, pl->nlon, p.gs, (long long)R * pl->nlat, bf16 ? 40 : 32);
    if (!rc) rc = make_tmap_segments(&p.tmXp, x, bf16, pl->nlon, p.gs, (long long)R * pl->nlat, bf16 ? 40 : 36);
    if (rc) return rc;
  }
  const size_t raw_bytes = bf16 ? (size_t)8 * 16 * (40 + 40) * 2 : (size_t)8 * 16 * (32 + 36) * 4;
  const size_t smem = 1024 + 3 * 8192 + (size_t)kDftAnaStages * 65536 + p.nraw * raw_bytes + 3 * 7 * 32 * 8 + 256;
  const int sms = usable_sms(pl->sm_count > 0 ? pl->sm_count : 148);
  const int ctas = p.ntiles < sms ? p.ntiles : sms;
  const int threads = 32 * (6 + (t->nkb == 3 ? 12 : 8));
#define B200_LAUNCH_ANA(TT, NN)                                                                                                          \
  do {                                                                                                                                  \
    struct Tag {};                                                                                                                        \
    B200_CHECK_CUDA((ensure_dynamic_smem<Tag>(dft_analysis_kernel<TT, NN>, smem)));                                                      \
    B200_CHECK_CUDA(launch_pdl(dft_analysis_kernel<TT, NN>, dim3(ctas), dim3(threads), smem, st, p));                                                                         \
  } while (0)
#define B200_DISPATCH_ANA(TT)                                            \
  switch (t->N2) {                                                       \
    case 180: B200_LAUNCH_ANA(TT, 180); break; /* nlon 1440 */           \
    case 90: B200_LAUNCH_ANA(TT, 90); break;   /* nlon  720 */           \
    case 60: B200_LAUNCH_ANA(TT, 60); break;   /* nlon  480 */           \
    default: B200_LAUNCH_ANA(TT, 0); break;                              \
  }
  if (bf16) { B200_DISPATCH_ANA(__nv_bfloat16) } else { B200_DISPATCH_ANA(float) }
#undef B200_DISPATCH_ANA
#undef B200_LAUNCH_ANA
  B200_CHECK_LAUNCH();
  return 0;
}

// ============================================================================================ host emulation
// The same factorisation and the same __host__ __device__ radix-8 code as the kernels, with the tensor-core sums done in double
// precision on the host: unit-tests the index maps, twiddles and butterflies without a GPU (b200sht_debug_dft_host).
int dft_host(int N, int mmax, int direction, int mode, const float* rowscale, const float* in, float* out) {
  if (!dft_shape_ok(N, mmax)) { set_error("debug_dft_host: unsupported (nlon=%d, mmax=%d)", N, mmax); return B200SHT_ERR_UNSUPPORTED; }
  const int N2 = N / 8, half = N2 / 2, M2 = (mmax + 7) / 8;
  const float rs = rowscale ? rowscale[0] : 1.f;
  auto twid = [&](int j, float2* tw) {
    for (int c = 0; c < 8; ++c) {
      const double a = 2.0 * M_PI * (double)(c * j) / (double)N;
      tw[c] = make_float2((float)cos(a), (float)sin(a));
    }
  };
  if (direction == 1) {
    // in: float[2 * mmax] interleaved (re, im) -> out: float[N]
    std::vector<double> zr(8 * M2, 0.0), zi(8 * M2, 0.0);
    for (int m = 0; m < mmax; ++m) { zr[m] = in[2 * m]; zi[m] = in[2 * m + 1]; }
    const bool has_nyq = (mmax == N / 2 + 1);
    for (int j2 = 0; j2 <= half; ++j2) {
      float s1[8], s2[8], s3[8], s4[8];
      for (int c = 0; c < 8; ++c) {
        double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        for (int m2 = 0; m2 < M2; ++m2) {
          const double b = 2.0 * M_PI * (double)(((long long)m2 * j2) % N2) / (double)N2;
          a1 += cos(b) * zr[c + 8 * m2]; a2 += sin(b) * zi[c + 8 * m2]; a3 += sin(b) * zr[c + 8 * m2]; a4 += cos(b) * zi[c + 8 * m2];
        }
        s1[c] = (float)a1; s2[c] = (float)a2; s3[c] = (float)a3; s4[c] = (float)a4;
      }
      float2 tw[8], tp[8];
      twid(j2, tw);
      dft_partner_twiddles(tw, tp);
      for (int side = 0; side < 2; ++side) {
        if (side == 1 && (j2 == 0 || 2 * j2 == N2)) continue;
        const int jj = side ? N2 - j2 : j2;
        float vr[8], vi[8], x[8];
        for (int c = 0; c < 8; ++c) {
          vr[c] = side ? s1[c] + s2[c] : s1[c] - s2[c];
          vi[c] = side ? s4[c] - s3[c] : s3[c] + s4[c];
        }
        dft_syn_radix8<float>(vr, vi, side ? tp : tw, x);
        for (int j1 = 0; j1 < 8; ++j1) {
          const int j = N2 * j1 + jj;
          float v = x[j1];
          if (mode == 0) {
            v = 2.f * v - (float)zr[0];
            if (has_nyq) v -= (float)zr[N / 2] * ((j & 1) ? -1.f : 1.f);
          } else {
            v *= rs;
          }
          out[j] = v;
        }
      }
    }
    return 0;
  }
  // analysis: in float[N] -> out float[2 * mmax]
  std::vector<double> dre(8 * M2, 0.0), dim(8 * M2, 0.0);
  for (int j2 = 0; j2 <= half; ++j2) {
    float2 tw[8], tp[8];
    twid(j2, tw);
    dft_partner_twiddles(tw, tp);
    float xa[8], er[8], ei[8], orr[8], oi[8];
    for (int j1 = 0; j1 < 8; ++j1) xa[j1] = in[N2 * j1 + j2];
    dft_ana_radix8<float>(xa, tw, er, ei);
    const bool paired = (j2 != 0 && 2 * j2 != N2);
    if (paired) {
      float xb[8], br[8], bi[8];
      for (int j1 = 0; j1 < 8; ++j1) xb[j1] = in[N2 * j1 + N2 - j2];
      dft_ana_radix8<float>(xb, tp, br, bi);
      for (int c = 0; c < 8; ++c) { orr[c] = er[c] - br[c]; oi[c] = ei[c] - bi[c]; er[c] += br[c]; ei[c] += bi[c]; }
    } else {
      for (int c = 0; c < 8; ++c) { orr[c] = 0.f; oi[c] = 0.f; }
    }
    for (int c = 0; c < 8; ++c)
      for (int m2 = 0; m2 < M2; ++m2) {
        const double b = 2.0 * M_PI * (double)(((long long)m2 * j2) % N2) / (double)N2;
        dre[c + 8 * m2] += cos(b) * er[c] + sin(b) * oi[c];
        dim[c + 8 * m2] += cos(b) * ei[c] - sin(b) * orr[c];
      }
  }
  for (int m = 0; m < mmax; ++m) {
    const double sc = mode == 0 ? (double)rs : ((m == 0 || 2 * m == N) ? 1.0 : 2.0);
    out[2 * m] = (float)(dre[m] * sc);
    out[2 * m + 1] = (float)(dim[m] * sc);
  }
  return 0;
}

}  // namespace b200sht
