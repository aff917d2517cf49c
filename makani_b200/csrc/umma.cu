// tcgen05 / TMA path (B200SHT_PREC_TF32): the Legendre contractions and the dense channel mix as TMA-fed tensor-core
// GEMMs with fp32 accumulators in TMEM.
//
//   one persistent engine (umma_kernel<Traits>, one CTA per SM walking the tile list): warp 0 lane 0 = TMA producer, warp 1
//   lane 0 = tcgen05.mma issuer, warps 2..5 = epilogue (warp w owns TMEM lanes 32(w%4)..+31).  A ring of `stages` operand
//   stages guarded by full/empty mbarriers runs continuously across tiles; two accumulator sets in TMEM (acc_full/acc_empty)
//   let the epilogue of tile i overlap the main loop of tile i+1.
//
//   five Traits supply the per-operation pieces (tile coordinates, TMA boxes, MMA issue list, epilogue):
//     AnaTraits   spec[l][m][n]  = sum_k P[m][l][k] X[m][n][k]          A K-major,  B K-major     (RealSHT einsum "...km,mlk->...lm")
//     SynTraits   Z[m][n][k]     = sum_l P[m][l][k] spec[l][m][n]       A MN-major, B MN-major    (InverseRealSHT "...lm,mlk->...km")
//     MixFwd      y[row][o]      = sum_i x[row][i] w[i][o]   (complex)  A K-major,  B MN-major    (contractions.py:23 "bgixy,giox->bgoxy")
//     MixDgrad    gx[row][i]     = sum_o gy[row][o] conj(w[i][o])       A K-major,  B K-major
//     MixWgrad    gw[i][o]       = sum_row conj(x[row][i]) gy[row][o]   A MN-major, B MN-major
//   complex products use planar operands: 4 real MMAs into two accumulators (real, imaginary), one of them with the
//   instruction descriptor's negate-A bit.
//
// All shared-memory operand tiles use the 128-byte swizzle; every TMA box is [rows][32 floats] so it lands as rows of 128 B.
#include "umma_common.cuh"
#include <mutex>

namespace b200sht {

// ======================================================================================================= engine
constexpr int kUmmaThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int kMaxStages = 8;
constexpr int kEpiScratch = 256;     // ints of per-tile epilogue scratch (one per accumulator column)

struct EngineParams {
  int stages;
  uint32_t stage_bytes, tx_bytes, tmem_cols;
  int gx, gy, gz;          // logical tile grid (x fastest); CTAs walk it round-robin
  int acc_cols, nbuf;      // TMEM columns of one accumulator set, number of sets (2: epilogue of tile i overlaps main loop of i+1)
  int split;               // 3 x TF32 (strict fp32 on the tensor cores): every stage also holds the residual tiles of both operands, `lo_off`
  uint32_t lo_off;         // bytes after the main tiles, and each MMA becomes hi.hi + hi.lo + lo.hi into the same accumulator
};


// Persistent engine: one CTA per SM loops over tiles.  The operand ring (full/empty) runs continuously across tiles, so the
// TMA producer prefetches the next tile while the tensor core finishes the current one and the four epilogue warps drain the
// previous accumulator set (acc_full / acc_empty).
template <class T>
__global__ void __launch_bounds__(kUmmaThreads, 1) umma_kernel(const __grid_constant__ typename T::Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int stages = p.stages;
  const uint32_t stage_bytes = p.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(gbase + (size_t)stages * stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* epi_scratch = reinterpret_cast<int*>(tmem_slot + 4);   // 2 x kEpiScratch ints, double-buffered by tile parity (epilogue warps only)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  pdl_trigger();   // the next kernel of the stream may be scheduled while this one runs (it waits for our completion before touching data)
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
    T::prefetch(p);
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles = p.gx * p.gy * p.gz;
  const int nbuf = p.nbuf;
  pdl_wait();      // prologue done (barriers, TMEM, tensor-map prefetch): from here on this kernel reads what its predecessors wrote

  if (warp == 0) {
    if (lane == 0) {
      int kbg = 0;   // k-block counter across tiles: position in the operand ring
      for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        typename T::Tile tile;
        if (!T::make_tile(p, tile, ti % p.gx, (ti / p.gx) % p.gy, ti / (p.gx * p.gy))) continue;
        const int nk = T::num_kblocks(p, tile);
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const int s = kbg % stages, it = kbg / stages;
          if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
          mbar_expect_tx(&full[s], p.tx_bytes);
          T::load(p, tile, kb, base + s * stage_bytes, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // all 32 lanes run the loop (converged); MMAs and commits are issued by an elected lane (umma_*_ws)
      int kbg = 0, i = 0;
      for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        typename T::Tile tile;
        if (!T::make_tile(p, tile, ti % p.gx, (ti / p.gx) % p.gy, ti / (p.gx * p.gy))) continue;
        const int nk = T::num_kblocks(p, tile);
        const int buf = i % nbuf, use = i / nbuf;
        if (use > 0) {  // the epilogue must have drained this accumulator set
          mbar_wait(&acc_empty[buf], (use - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const int s = kbg % stages, it = kbg / stages;
          mbar_wait(&full[s], it & 1);
          tc_fence_after();
          T::mma(p, tile, base + s * stage_bytes, tmem + buf * p.acc_cols, kb > 0);
          umma_commit_ws(&empty[s]);
        }
        umma_commit_ws(&acc_full[buf]);
        ++i;
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;   // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    int i = 0;
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
      typename T::Tile tile;
      if (!T::make_tile(p, tile, ti % p.gx, (ti / p.gx) % p.gy, ti / (p.gx * p.gy))) continue;
      const int nk = T::num_kblocks(p, tile);
      const int buf = i % nbuf, use = i / nbuf;
      mbar_wait(&acc_full[buf], use & 1);
      tc_fence_after();
      T::epilogue(p, tile, tmem + buf * p.acc_cols, quad, lane, nk, epi_scratch + (i & 1) * kEpiScratch);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      ++i;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

// ================================================================================================ AnaTraits
struct AnaTraits {
  struct Params : EngineParams {
    alignas(64) CUtensorMap tmA;  // table  (k, l, m)       box (32, 128, 1)
    alignas(64) CUtensorMap tmB;  // X      (k, c, pb, m)   box (32, Cc, PBc, 1)
    alignas(64) CUtensorMap tmA_lo, tmB_lo;   // residuals of the table and of X (split mode)
    float* spec;
    int L, M, nlat, C, cp, PB, Cc, PBc, n_ct, N, m0;
    int kb0, nkb;        // latitude range of this launch in 32-row K-blocks (latitude-chunked analysis: partial sums over a chunk of rows)
    int acc_in, round_out;   // add to the spec values already stored (chunks after the first) / round the result to TF32 (last chunk)
    uint32_t idesc;
  };
  struct Tile { int m, l0, c0, pb0; };
  __device__ static bool make_tile(const Params& p, Tile& t, int bx, int by, int bz) {
    t.m = bz;
    t.l0 = lstart(p.m0 + t.m) + 128 * bx;
    t.c0 = (by % p.n_ct) * p.Cc;
    t.pb0 = (by / p.n_ct) * p.PBc;
    return t.l0 < p.L;
  }
  __device__ static void prefetch(const Params& p) { prefetch_tmap(&p.tmA); prefetch_tmap(&p.tmB); }
  __device__ static int num_kblocks(const Params& p, const Tile&) { return p.nkb; }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint32_t st, uint64_t* bar) {
    kb += p.kb0;
    tma_load_3d(st, &p.tmA, bar, kb * 32, t.l0, t.m);
    tma_load_4d(st + 16384, &p.tmB, bar, kb * 32, t.c0, t.pb0, t.m);
    if (p.split) {
      tma_load_3d(st + p.lo_off, &p.tmA_lo, bar, kb * 32, t.l0, t.m);
      tma_load_4d(st + p.lo_off + 16384, &p.tmB_lo, bar, kb * 32, t.c0, t.pb0, t.m);
    }
  }
  __device__ static void mma(const Params& p, const Tile&, uint32_t st, uint32_t tmem, bool acc) {
    const uint64_t a = desc_kmajor(st, 0), b = desc_kmajor(st + 16384, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) umma_tf32_ws(tmem, desc_advance(a, 32 * j), desc_advance(b, 32 * j), p.idesc, (acc || j > 0) ? 1u : 0u);
    if (p.split) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma_tf32_ws(tmem, desc_advance(a, 32 * j), desc_advance(b, p.lo_off + 32 * j), p.idesc, 1u);   // hi . lo
        umma_tf32_ws(tmem, desc_advance(a, p.lo_off + 32 * j), desc_advance(b, 32 * j), p.idesc, 1u);   // lo . hi
      }
    }
  }
  __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int warp, int lane, int nk, int* scratch) {
    const int l = t.l0 + warp * 32 + lane;
    const int ncols = p.Cc * p.PBc;
    const size_t JP = (size_t)p.PB * p.cp;
    float* orow = p.spec + ((size_t)(l < p.L ? l : 0) * p.M + t.m) * JP;
    float v[32];
    for (int n0 = 0; n0 < ncols; n0 += 32) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
      if (l >= p.L) continue;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int n = n0 + q * 4;
        if (n >= ncols) break;
        const int pbi = n / p.Cc, ci = n - pbi * p.Cc;
        const int pb = t.pb0 + pbi, c = t.c0 + ci;
        if (pb < p.PB && c < p.cp) {
          float4* dst = reinterpret_cast<float4*>(orow + (size_t)pb * p.cp + c);
          float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          if (p.acc_in) {   // partial sums of the earlier latitude chunks (unrounded fp32)
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          // strict fp32 (split) and the partial sums of a chunked analysis stay as accumulated; otherwise the consumers are kind::tf32
          // MMAs: round to nearest here
          if (p.round_out) o = make_float4(tf32_rn(o.x), tf32_rn(o.y), tf32_rn(o.z), tf32_rn(o.w));
          *dst = o;
        }
      }
    }
  }
};

// base[o] = v when o >= 0: compare + one wide multiply-add + predicated store, no branch
__device__ __forceinline__ void st_if_nonneg(float* base, int o, float v) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .s64 a;\n"
      "setp.ge.s32 p, %1, 0;\n"
      "mad.wide.s32 a, %1, 4, %0;\n"
      "@p st.global.f32 [a], %2;\n"
      "}\n" ::"l"(base), "r"(o), "f"(v) : "memory");
}

// ================================================================================================ SynTraits
struct SynTraits {
  struct Params : EngineParams {
    alignas(64) CUtensorMap tmA;  // table (k, l, m)   box (32, 32, 1)   MN-major A (M = k)
    alignas(64) CUtensorMap tmB;  // spec  (n, m, l)   box (32, 1, 32)   MN-major B (N = n)
    alignas(64) CUtensorMap tmA_lo, tmB_lo;   // residuals of the table and of spec (split mode)
    float* Z;
    int L, M, nlat, kp, C, cp, PB, nblk, N, m0;
    int kc0, kc1;           // latitude range [kc0, kc1) of this launch (kc0 a multiple of 128): the tiles cover these rows only
    int tiled, M2, KT, B;   // tiled output for the tensor-core DFT (dft.cu): Z[r][k / 8][p][m / 8][m % 8][k % 8], orders padded to 8 * M2
    uint32_t idesc;
  };
  struct Tile { int m, k0, n0, lbeg; };
  __device__ static bool make_tile(const Params& p, Tile& t, int bx, int by, int bz) {
    t.m = bz;
    t.k0 = p.kc0 + 128 * bx;
    t.n0 = p.N * by;
    t.lbeg = lstart(p.m0 + t.m);
    return true;
  }
  __device__ static void prefetch(const Params& p) { prefetch_tmap(&p.tmA); prefetch_tmap(&p.tmB); }
  // orders m >= M exist only in the tiled layout (padding up to a multiple of 8): no degree contributes, the tile is written as zeros
  __device__ static int num_kblocks(const Params& p, const Tile& t) { return (t.m < p.M && t.lbeg < p.L) ? (p.L - t.lbeg + 31) / 32 : 0; }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint32_t st, uint64_t* bar) {
    const int l = t.lbeg + kb * 32;
#pragma unroll
    for (int b = 0; b < 4; ++b) tma_load_3d(st + b * 4096, &p.tmA, bar, t.k0 + 32 * b, l, t.m);
    for (int b = 0; b < p.nblk; ++b) tma_load_3d(st + 16384 + b * 4096, &p.tmB, bar, t.n0 + 32 * b, t.m, l);
    if (p.split) {
#pragma unroll
      for (int b = 0; b < 4; ++b) tma_load_3d(st + p.lo_off + b * 4096, &p.tmA_lo, bar, t.k0 + 32 * b, l, t.m);
      for (int b = 0; b < p.nblk; ++b) tma_load_3d(st + p.lo_off + 16384 + b * 4096, &p.tmB_lo, bar, t.n0 + 32 * b, t.m, l);
    }
  }
  __device__ static void mma(const Params& p, const Tile&, uint32_t st, uint32_t tmem, bool acc) {
    const uint64_t a = desc_mnmajor(st, 0, 4096), b = desc_mnmajor(st + 16384, 0, 4096);
#pragma unroll
    for (int j = 0; j < 4; ++j) umma_tf32_ws(tmem, desc_advance(a, 1024 * j), desc_advance(b, 1024 * j), p.idesc, (acc || j > 0) ? 1u : 0u);
    if (p.split) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma_tf32_ws(tmem, desc_advance(a, 1024 * j), desc_advance(b, p.lo_off + 1024 * j), p.idesc, 1u);   // hi . lo
        umma_tf32_ws(tmem, desc_advance(a, p.lo_off + 1024 * j), desc_advance(b, 1024 * j), p.idesc, 1u);   // lo . hi
      }
    }
  }
  __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int warp, int lane, int nk, int* scratch) {
    const int k = t.k0 + warp * 32 + lane;
    const int JP = p.PB * p.cp;
    // column jp = pb * cp + c -> element offset of row (pb, c) of this order's slab of Z, -1 for the channel padding: one division per
    // column per tile, shared by the four epilogue warps through `scratch` (the per-element bookkeeping of an incremental row
    // pointer made this kernel epilogue-bound: 60 % of its stall samples)
    for (int n = warp * 32 + lane; n < p.N; n += 128) {
      const int jp = t.n0 + n;
      int o = -1;
      if (jp < JP) {
        const int pb = jp / p.cp, c = jp - pb * p.cp;
        if (c < p.C) {
          if (!p.tiled) o = (pb * p.C + c) * p.kp;
          else {   // pb = plane * B + b, image r = b * C + c:  Z[r][kt][plane][m2][c8][k8]
            const int pl = pb / p.B, b = pb - pl * p.B;
            o = (((b * p.C + c) * p.KT) * 2 + pl) * p.M2 * 64;
          }
        }
      }
      scratch[n] = o;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
    const bool kok = k < p.kc1;
    const int kk = kok ? k : 0;
    float* zb = p.tiled ? p.Z + (size_t)(kk >> 3) * 2 * p.M2 * 64 + (t.m >> 3) * 64 + (t.m & 7) * 8 + (kk & 7)
                        : p.Z + (size_t)t.m * p.PB * p.C * p.kp + kk;
    float v[32];
    for (int n0 = 0; n0 < p.N; n0 += 32) {
      if (t.n0 + n0 >= JP) break;
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
      if (nk == 0) {   // no degree contributes to this order (cannot happen for m < lmax; kept for safety)
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
      }
      if (kok) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const int4 o = *reinterpret_cast<const int4*>(scratch + n0 + 4 * q4);
          st_if_nonneg(zb, o.x, v[4 * q4 + 0]);
          st_if_nonneg(zb, o.y, v[4 * q4 + 1]);
          st_if_nonneg(zb, o.z, v[4 * q4 + 2]);
          st_if_nonneg(zb, o.w, v[4 * q4 + 3]);
        }
      }
    }
  }
};

// ====================================================================================================== mix
// spec tensor map: dims (c, b, p, m, l) with strides (1, cp, B*cp, 2*B*cp, M*2*B*cp) floats.
// weight tensor map (planar packed weight [Lw][G][Cig][2][cop]): dims (o, p, i, lg) strides (1, cop, 2*cop, Cig*2*cop).
struct MixParams : EngineParams {
  alignas(64) CUtensorMap tmX;   // operand read as rows (m, b)
  alignas(64) CUtensorMap tmX2;  // second spec operand (wgrad: gy)
  alignas(64) CUtensorMap tmW;
  float* out;              // spec (fwd / dgrad) or packed weight gradient (wgrad)
  const float2* cbias;
  int L, M, B, G, Cig, Cog, cpi, cpo, cop;
  int Mt;                  // m values per row tile (Mt * B <= 128 rows)
  int nblk, N;             // N = output columns per tile; nblk = N / 32 (MN-major B operands)
  int n_nt;                // output tiles per group
  int shared_w;            // weight has no l dimension
  int dense;               // spec tensors store every (l, m) entry
  long long wl_stride;
  uint32_t idesc, idesc_neg;
  uint32_t offA_i, offB_r, offB_i;  // stage offsets of the imaginary A tile and the two B tiles (A_r at 0)
};

struct MixFwdTraits {
  using Params = MixParams;
  struct Tile { int l, m0, g, o0, lg; };
  __device__ static bool make_tile(const Params& p, Tile& t, int bx, int by, int bz) {
    t.l = bz;
    t.m0 = bx * p.Mt;
    t.g = by / p.n_nt;
    t.o0 = (by % p.n_nt) * p.N;
    t.lg = (p.shared_w ? 0 : t.l * p.G) + t.g;
    return t.m0 < mend_d(t.l, p.M, p.dense);
  }
  __device__ static void prefetch(const Params& p) { prefetch_tmap(&p.tmX); prefetch_tmap(&p.tmW); }
  __device__ static int num_kblocks(const Params& p, const Tile&) { return (p.Cig + 31) / 32; }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint32_t st, uint64_t* bar) {
    const int c = t.g * p.Cig + kb * 32;
    tma_load_5d(st, &p.tmX, bar, c, 0, 0, t.m0, t.l);
    tma_load_5d(st + p.offA_i, &p.tmX, bar, c, 0, 1, t.m0, t.l);
    for (int b = 0; b < p.nblk; ++b) {
      tma_load_4d(st + p.offB_r + b * 4096, &p.tmW, bar, t.o0 + 32 * b, 0, kb * 32, t.lg);
      tma_load_4d(st + p.offB_i + b * 4096, &p.tmW, bar, t.o0 + 32 * b, 1, kb * 32, t.lg);
    }
  }
  __device__ static void mma(const Params& p, const Tile&, uint32_t st, uint32_t tmem, bool acc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t ar = desc_advance(desc_kmajor(st, 0), 32 * j), ai = desc_advance(desc_kmajor(st + p.offA_i, 0), 32 * j);
      const uint64_t br = desc_advance(desc_mnmajor(st + p.offB_r, 0, 4096), 1024 * j), bi = desc_advance(desc_mnmajor(st + p.offB_i, 0, 4096), 1024 * j);
      const uint32_t a0 = (acc || j > 0) ? 1u : 0u;
      umma_tf32_ws(tmem, ar, br, p.idesc, a0);            // yr  = xr wr
      umma_tf32_ws(tmem, ai, bi, p.idesc_neg, 1u);        // yr -= xi wi
      umma_tf32_ws(tmem + p.N, ar, bi, p.idesc, a0);      // yi  = xr wi
      umma_tf32_ws(tmem + p.N, ai, br, p.idesc, 1u);      // yi += xi wr
    }
  }
  // rows (mi, b) -> spec rows; columns -> output channels of group g; handles cbias and the zero channel padding
  __device__ static void store_rows(const Params& p, int l, int m0, int g, int o0, int NOg, int cp_out, uint32_t tmem, int warp, int lane,
                                    bool with_bias) {
    const int r = warp * 32 + lane;
    const int m = m0 + r / p.B, b = r % p.B;
    const bool row_ok = (r < p.Mt * p.B) && (m < mend_d(l, p.M, p.dense));
    const int pad = cp_out - NOg * p.G;
    const int limit = NOg + ((g == p.G - 1) ? pad : 0);  // columns of this group incl. trailing zero padding
    float* yr = p.out + ((size_t)(row_ok ? l : 0) * p.M + (row_ok ? m : 0)) * 2 * p.B * cp_out + (size_t)b * cp_out + g * NOg;
    float* yi = yr + (size_t)p.B * cp_out;
    float vr[32], vi[32];
    for (int n0 = 0; n0 < p.N; n0 += 32) {
      if (o0 + n0 >= limit) break;
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, vr);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + p.N + n0, vi);
      if (!row_ok) continue;
      // 16-byte stores (one row per thread: a warp store touches 32 rows, so wide stores cut the L2 write transactions 4x)
      const bool vec_ok = (((g * NOg) & 3) == 0);   // cp_out and o0 + n0 are multiples of 4
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const int ob = o0 + n0 + q4 * 4;
        if (ob >= limit) break;
        float a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int o = ob + u;
          a[u] = vr[q4 * 4 + u];
          c[u] = vi[q4 * 4 + u];
          if (o >= NOg) { a[u] = 0.f; c[u] = 0.f; }
          else if (with_bias) { const float2 cb = p.cbias[g * NOg + o]; a[u] += cb.x; c[u] += cb.y; }
          a[u] = tf32_rn(a[u]);
          c[u] = tf32_rn(c[u]);
        }
        if (vec_ok && ob + 3 < limit) {
          *reinterpret_cast<float4*>(yr + ob) = make_float4(a[0], a[1], a[2], a[3]);
          *reinterpret_cast<float4*>(yi + ob) = make_float4(c[0], c[1], c[2], c[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (ob + u < limit) { yr[ob + u] = a[u]; yi[ob + u] = c[u]; }
        }
      }
    }
  }
  __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int warp, int lane, int nk, int* scratch) {
    store_rows(p, t.l, t.m0, t.g, t.o0, p.Cog, p.cpo, tmem, warp, lane, p.cbias != nullptr);
  }
};

struct MixDgradTraits {
  using Params = MixParams;   // tmX = gy (channels = Cout), out = gx; N tiles over i
  using Tile = MixFwdTraits::Tile;  // o0 is the first input channel i0 of the tile
  __device__ static bool make_tile(const Params& p, Tile& t, int bx, int by, int bz) { return MixFwdTraits::make_tile(p, t, bx, by, bz); }
  __device__ static void prefetch(const Params& p) { prefetch_tmap(&p.tmX); prefetch_tmap(&p.tmW); }
  __device__ static int num_kblocks(const Params& p, const Tile&) { return (p.Cog + 31) / 32; }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint32_t st, uint64_t* bar) {
    const int c = t.g * p.Cog + kb * 32;
    tma_load_5d(st, &p.tmX, bar, c, 0, 0, t.m0, t.l);
    tma_load_5d(st + p.offA_i, &p.tmX, bar, c, 0, 1, t.m0, t.l);
    tma_load_4d(st + p.offB_r, &p.tmW, bar, kb * 32, 0, t.o0, t.lg);   // box (32 o, 1, N i, 1): K-major rows i
    tma_load_4d(st + p.offB_i, &p.tmW, bar, kb * 32, 1, t.o0, t.lg);
  }
  __device__ static void mma(const Params& p, const Tile&, uint32_t st, uint32_t tmem, bool acc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t ar = desc_advance(desc_kmajor(st, 0), 32 * j), ai = desc_advance(desc_kmajor(st + p.offA_i, 0), 32 * j);
      const uint64_t br = desc_advance(desc_kmajor(st + p.offB_r, 0), 32 * j), bi = desc_advance(desc_kmajor(st + p.offB_i, 0), 32 * j);
      const uint32_t a0 = (acc || j > 0) ? 1u : 0u;
      umma_tf32_ws(tmem, ar, br, p.idesc, a0);            // gxr  = gr wr
      umma_tf32_ws(tmem, ai, bi, p.idesc, 1u);            // gxr += gi wi
      umma_tf32_ws(tmem + p.N, ai, br, p.idesc, a0);      // gxi  = gi wr
      umma_tf32_ws(tmem + p.N, ar, bi, p.idesc_neg, 1u);  // gxi -= gr wi
    }
  }
  __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int warp, int lane, int nk, int* scratch) {
    MixFwdTraits::store_rows(p, t.l, t.m0, t.g, t.o0, p.Cig, p.cpi, tmem, warp, lane, false);
  }
};

struct MixWgradTraits {
  using Params = MixParams;   // tmX = x (A, rows i), tmX2 = gy (B, cols o); K = spectral rows (m, b)
  struct Tile { int lz, i0, g, o0; };
  __device__ static bool make_tile(const Params& p, Tile& t, int bx, int by, int bz) {
    t.lz = bz;
    t.i0 = bx * 128;
    t.g = by / p.n_nt;
    t.o0 = (by % p.n_nt) * p.N;
    return true;
  }
  __device__ static void prefetch(const Params& p) { prefetch_tmap(&p.tmX); prefetch_tmap(&p.tmX2); }
  __device__ static int kb_of_l(const Params& p, int l) { return (mend_d(l, p.M, p.dense) * p.B + 31) / 32; }
  __device__ static int num_kblocks(const Params& p, const Tile& t) {
    if (!p.shared_w) return kb_of_l(p, t.lz);
    int n = 0;
    for (int l = 0; l < p.L; ++l) n += kb_of_l(p, l);
    return n;
  }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint32_t st, uint64_t* bar) {
    int l = t.lz;
    if (p.shared_w) {
      l = 0;
      int n = kb_of_l(p, 0);
      while (kb >= n) { kb -= n; ++l; n = kb_of_l(p, l); }
    }
    const int m = kb * (32 / p.B);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      tma_load_5d(st + a * 4096, &p.tmX, bar, t.g * p.Cig + t.i0 + 32 * a, 0, 0, m, l);
      tma_load_5d(st + p.offA_i + a * 4096, &p.tmX, bar, t.g * p.Cig + t.i0 + 32 * a, 0, 1, m, l);
    }
    for (int b = 0; b < p.nblk; ++b) {
      tma_load_5d(st + p.offB_r + b * 4096, &p.tmX2, bar, t.g * p.Cog + t.o0 + 32 * b, 0, 0, m, l);
      tma_load_5d(st + p.offB_i + b * 4096, &p.tmX2, bar, t.g * p.Cog + t.o0 + 32 * b, 0, 1, m, l);
    }
  }
  __device__ static void mma(const Params& p, const Tile&, uint32_t st, uint32_t tmem, bool acc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t ar = desc_advance(desc_mnmajor(st, 0, 4096), 1024 * j), ai = desc_advance(desc_mnmajor(st + p.offA_i, 0, 4096), 1024 * j);
      const uint64_t br = desc_advance(desc_mnmajor(st + p.offB_r, 0, 4096), 1024 * j), bi = desc_advance(desc_mnmajor(st + p.offB_i, 0, 4096), 1024 * j);
      const uint32_t a0 = (acc || j > 0) ? 1u : 0u;
      umma_tf32_ws(tmem, ar, br, p.idesc, a0);            // gwr  = xr gr
      umma_tf32_ws(tmem, ai, bi, p.idesc, 1u);            // gwr += xi gi
      umma_tf32_ws(tmem + p.N, ar, bi, p.idesc, a0);      // gwi  = xr gi
      umma_tf32_ws(tmem + p.N, ai, br, p.idesc_neg, 1u);  // gwi -= xi gr
    }
  }
  __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int warp, int lane, int nk, int* scratch) {
    const int i = t.i0 + warp * 32 + lane;
    const bool ok = i < p.Cig;
    float* row = p.out + (size_t)(p.shared_w ? 0 : t.lz) * p.wl_stride + (size_t)(t.g * p.Cig + (ok ? i : 0)) * 2 * p.cop;
    float vr[32], vi[32];
    for (int n0 = 0; n0 < p.N; n0 += 32) {
      if (t.o0 + n0 >= p.cop) break;
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, vr);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + p.N + n0, vi);
      if (!ok) continue;
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {   // cop is a multiple of 4: whole float4 groups, 16-byte aligned
        const int ob = t.o0 + n0 + q4 * 4;
        if (ob >= p.cop) break;
        float a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = (ob + u < p.Cog) ? vr[q4 * 4 + u] : 0.f;
          c[u] = (ob + u < p.Cog) ? vi[q4 * 4 + u] : 0.f;
        }
        *reinterpret_cast<float4*>(row + ob) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(row + p.cop + ob) = make_float4(c[0], c[1], c[2], c[3]);
      }
    }
  }
};

// ================================================================================================== host side
// per device (a process may hold tensors on several GPUs): -1 unknown, 0 / 1
constexpr int kMaxDevices = 64;
static int g_umma_ok[kMaxDevices];
static int g_sm_count[kMaxDevices];
static std::once_flag g_dev_once;
static void init_dev_caches() { for (int i = 0; i < kMaxDevices; ++i) { g_umma_ok[i] = -1; g_sm_count[i] = 0; } }
int umma_available() {
  std::call_once(g_dev_once, init_dev_caches);
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < kMaxDevices && g_umma_ok[dev] >= 0) return g_umma_ok[dev];
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  const int ok = (major == 10 && get_encode() != nullptr) ? 1 : 0;
  if (dev >= 0 && dev < kMaxDevices) g_umma_ok[dev] = ok;
  return ok;
}

int round_table_tf32(const float* src, float* dst, size_t n, cudaStream_t st);  // legendre.cu

int umma_plan_init(Plan* pl) {
  pl->umma_state = nullptr;
  pl->d_table_tf32 = nullptr;
  pl->d_table_lo = nullptr;
  if (!umma_available()) return -1;
  const size_t n = (size_t)pl->mmax * pl->lmax * pl->kp;
  if (cudaMalloc(&pl->d_table_tf32, n * sizeof(float)) != cudaSuccess) { pl->d_table_tf32 = nullptr; return -1; }
  if (round_table_tf32(pl->d_table, pl->d_table_tf32, n, 0) != 0 || cudaStreamSynchronize(0) != cudaSuccess) {
    cudaFree(pl->d_table_tf32);
    pl->d_table_tf32 = nullptr;
    return -1;
  }
  return 0;
}
void umma_plan_destroy(Plan* pl) {
  if (pl->d_table_tf32) cudaFree(pl->d_table_tf32);
  pl->d_table_tf32 = nullptr;
  if (pl->d_table_lo) cudaFree(pl->d_table_lo);
  pl->d_table_lo = nullptr;
}

int table_residual(const float* full, const float* hi, float* lo, size_t n, cudaStream_t st);  // legendre.cu
// residual table of the 3 x TF32 mode: built the first time a strict-fp32 Legendre stage runs on the tensor cores (most plans never need it)
int umma_plan_table_lo(const Plan* cpl) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Plan* pl = const_cast<Plan*>(cpl);
  if (pl->d_table_lo) return 0;
  B200_REQUIRE(pl->d_table && pl->d_table_tf32, "3 x TF32: the plan has no Legendre table");
  const size_t n = (size_t)pl->mmax * pl->lmax * pl->kp;
  float* lo = nullptr;
  B200_CHECK_CUDA(cudaMalloc(&lo, n * sizeof(float)));
  int rc = table_residual(pl->d_table, pl->d_table_tf32, lo, n, 0);
  if (!rc && cudaStreamSynchronize(0) != cudaSuccess) rc = B200SHT_ERR_CUDA;
  if (rc) { cudaFree(lo); return rc; }
  pl->d_table_lo = lo;
  return 0;
}

constexpr size_t kSmemMax = 232448 - 4096;  // 227 KB minus barriers / epilogue scratch / alignment slack

static void pick_stages(EngineParams* e, uint32_t stage_bytes, int /*k-blocks per tile: the ring runs across tiles*/) {
  e->stage_bytes = stage_bytes;
  int s = (int)(kSmemMax / stage_bytes);   // persistent: one CTA per SM owns the whole shared memory
  if (s > kMaxStages) s = kMaxStages;
  if (s < 2) s = 2;
  e->stages = s;
}
static size_t smem_bytes(const EngineParams& e) { return (size_t)e.stages * e.stage_bytes + 1024 /*align*/ + (2 * kMaxStages + 4) * 8 + 16 + 2 * kEpiScratch * 4; }
static uint32_t tmem_cols_pow2(int cols) { uint32_t c = 32; while ((int)c < cols) c <<= 1; return c; }
// accumulator sets: `cols` TMEM columns per tile; two sets (double buffering) when they fit in the 512 columns
static void set_accumulators(EngineParams* e, int cols) {
  e->acc_cols = round_up(cols, 32);
  e->nbuf = (2 * e->acc_cols <= 512) ? 2 : 1;
  e->tmem_cols = tmem_cols_pow2(e->acc_cols * e->nbuf);
}

static int sm_count() {   // of the current device (the launch device: _lib.call makes the tensor's device current)
  std::call_once(g_dev_once, init_dev_caches);
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < kMaxDevices && g_sm_count[dev] > 0) return g_sm_count[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < kMaxDevices) g_sm_count[dev] = n;
  return n;
}

template <class T>
static int launch(typename T::Params& p, dim3 grid, cudaStream_t st) {
  p.gx = (int)grid.x; p.gy = (int)grid.y; p.gz = (int)grid.z;
  const long long ntiles = (long long)grid.x * grid.y * grid.z;
  if (ntiles <= 0) return 0;
  const size_t smem = smem_bytes(p);
  if (smem > 232448) { set_error("umma: %zu bytes of shared memory needed", smem); return B200SHT_ERR_UNSUPPORTED; }
  B200_CHECK_CUDA((ensure_dynamic_smem<T>(umma_kernel<T>, smem)));
  // static round-robin over tiles: an odd CTA count not divisible by 3 keeps tile-grid periods (2 l- or m-tiles, 3 or 6 n/k-tiles)
  // from locking heavy tiles onto the same CTAs
  const int sms = usable_sms(sm_count());
  int ctas = (int)(ntiles < sms ? ntiles : sms);
  while (ctas > 1 && (ctas % 2 == 0 || ctas % 3 == 0)) --ctas;
  B200_CHECK_CUDA(launch_pdl(umma_kernel<T>, dim3(ctas), dim3(kUmmaThreads), smem, st, p));
  B200_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------- Legendre
// k_begin / k_end: latitude range [k_begin, k_end) to reduce over (k_begin a multiple of 32; k_end < 0: all rows); accumulate: add to the spec
// values stored by the launches of the earlier ranges; last: this is the final range (TF32 rounding of the result happens here)
int legendre_analysis_umma(const Plan* pl, const float* X, float* spec, int B, int C, cudaStream_t st, const float* X_lo, int k_begin, int k_end,
                           int accumulate, int last) {
  AnaTraits::Params p;
  memset(&p, 0, sizeof(p));
  if (k_end < 0 || k_end > pl->nlat) k_end = pl->nlat;
  B200_REQUIRE(k_begin >= 0 && k_begin % 32 == 0 && k_begin < k_end, "legendre_analysis: bad latitude range [%d, %d)", k_begin, k_end);
  p.kb0 = k_begin / 32;
  p.nkb = ceil_div(k_end - k_begin, 32);
  p.acc_in = accumulate;
  p.round_out = (last && X_lo == nullptr) ? 1 : 0;
  const int cp = round_up(C, 4), PB = 2 * B;
  p.spec = spec; p.L = pl->lmax; p.M = pl->mmax; p.nlat = pl->nlat; p.C = C; p.cp = cp; p.PB = PB; p.m0 = pl->m0;
  if (cp <= 128) { p.Cc = cp; p.n_ct = 1; p.PBc = 256 / cp < PB ? 256 / cp : PB; }
  else { p.n_ct = ceil_div(cp, 128); p.Cc = round_up(ceil_div(cp, p.n_ct), 4); p.PBc = (2 * p.Cc <= 256 && PB >= 2) ? 2 : 1; }
  const int rows = p.Cc * p.PBc;
  p.N = round_up(rows, 16);
  p.idesc = make_idesc(p.N, 0, 0, 0);
  {
    long long d[3] = {pl->nlat, pl->lmax, pl->mmax}, s[3] = {1, pl->kp, (long long)pl->lmax * pl->kp};
    int bx[3] = {32, 128, 1};
    int rc = make_tmap(&p.tmA, pl->d_table_tf32, 3, d, s, bx);
    if (rc) return rc;
  }
  {
    long long d[4] = {pl->nlat, C, PB, pl->mmax}, s[4] = {1, pl->kp, (long long)C * pl->kp, (long long)PB * C * pl->kp};
    int bx[4] = {32, p.Cc, p.PBc, 1};
    int rc = make_tmap(&p.tmB, X, 4, d, s, bx);
    if (rc) return rc;
  }
  const uint32_t bbytes = (uint32_t)round_up(p.N * 128, 1024);
  p.split = X_lo != nullptr;
  if (p.split) {
    B200_REQUIRE(pl->d_table_lo != nullptr, "legendre_analysis (3 x TF32): the residual table is missing");
    long long d[3] = {pl->nlat, pl->lmax, pl->mmax}, s[3] = {1, pl->kp, (long long)pl->lmax * pl->kp};
    int bx[3] = {32, 128, 1};
    int rc = make_tmap(&p.tmA_lo, pl->d_table_lo, 3, d, s, bx);
    long long d4[4] = {pl->nlat, C, PB, pl->mmax}, s4[4] = {1, pl->kp, (long long)C * pl->kp, (long long)PB * C * pl->kp};
    int bx4[4] = {32, p.Cc, p.PBc, 1};
    if (!rc) rc = make_tmap(&p.tmB_lo, X_lo, 4, d4, s4, bx4);
    if (rc) return rc;
    p.lo_off = 16384 + bbytes;
  }
  pick_stages(&p, (16384 + bbytes) * (p.split ? 2 : 1), ceil_div(pl->nlat, 32));
  p.tx_bytes = (16384 + (uint32_t)rows * 128) * (p.split ? 2 : 1);
  set_accumulators(&p, p.N);
  dim3 grid(ceil_div(pl->lmax, 128), p.n_ct * ceil_div(PB, p.PBc), pl->mmax);
  return launch<AnaTraits>(p, grid, st);
}

// k_begin / k_end: latitude range [k_begin, k_end) to produce (k_begin a multiple of 128; k_end < 0: up to kp)
int legendre_synthesis_umma(const Plan* pl, const float* spec, float* Z, int B, int C, int tiled, cudaStream_t st, const float* spec_lo, int k_begin,
                            int k_end) {
  SynTraits::Params p;
  memset(&p, 0, sizeof(p));
  if (k_end < 0 || k_end > pl->kp) k_end = pl->kp;
  B200_REQUIRE(k_begin >= 0 && k_begin % 128 == 0 && k_begin < k_end, "legendre_synthesis: bad latitude range [%d, %d)", k_begin, k_end);
  p.kc0 = k_begin; p.kc1 = k_end;
  const int cp = round_up(C, 4), PB = 2 * B, JP = PB * cp;
  p.Z = Z; p.L = pl->lmax; p.M = pl->mmax; p.nlat = pl->nlat; p.kp = pl->kp; p.C = C; p.cp = cp; p.PB = PB; p.m0 = pl->m0;
  p.tiled = tiled; p.M2 = (pl->mmax + 7) / 8; p.KT = pl->kp / 8; p.B = B;
  B200_REQUIRE(!tiled || (long long)B * C * p.KT * 2 * p.M2 * 64 < (1ll << 31), "legendre_synthesis: tiled latspec of %d images exceeds 2^31 floats", B * C);
  p.nblk = ceil_div(JP, 32) < 8 ? ceil_div(JP, 32) : 8;
  p.N = 32 * p.nblk;
  p.idesc = make_idesc(p.N, 1, 1, 0);
  {
    long long d[3] = {pl->nlat, pl->lmax, pl->mmax}, s[3] = {1, pl->kp, (long long)pl->lmax * pl->kp};
    int bx[3] = {32, 32, 1};
    int rc = make_tmap(&p.tmA, pl->d_table_tf32, 3, d, s, bx, true);
    if (rc) return rc;
  }
  {
    long long d[3] = {JP, pl->mmax, pl->lmax}, s[3] = {1, JP, (long long)pl->mmax * JP};
    int bx[3] = {32, 1, 32};
    int rc = make_tmap(&p.tmB, spec, 3, d, s, bx, true);
    if (rc) return rc;
  }
  p.split = spec_lo != nullptr;
  if (p.split) {
    B200_REQUIRE(pl->d_table_lo != nullptr, "legendre_synthesis (3 x TF32): the residual table is missing");
    long long d[3] = {pl->nlat, pl->lmax, pl->mmax}, s[3] = {1, pl->kp, (long long)pl->lmax * pl->kp};
    int bx[3] = {32, 32, 1};
    int rc = make_tmap(&p.tmA_lo, pl->d_table_lo, 3, d, s, bx, true);
    long long d2[3] = {JP, pl->mmax, pl->lmax}, s2[3] = {1, JP, (long long)pl->mmax * JP};
    int bx2[3] = {32, 1, 32};
    if (!rc) rc = make_tmap(&p.tmB_lo, spec_lo, 3, d2, s2, bx2, true);
    if (rc) return rc;
    p.lo_off = 16384 + 4096 * p.nblk;
  }
  pick_stages(&p, (16384 + 4096 * p.nblk) * (p.split ? 2 : 1), ceil_div(pl->lmax, 32));
  p.tx_bytes = (16384 + 4096 * p.nblk) * (p.split ? 2 : 1);
  set_accumulators(&p, p.N);
  dim3 grid(ceil_div(k_end - k_begin, 128), ceil_div(JP, p.N), tiled ? 8 * p.M2 : pl->mmax);
  return launch<SynTraits>(p, grid, st);
}

// --------------------------------------------------------------------------------------------------- mix
static int spec_tmap(CUtensorMap* tm, const float* base, int L, int M, int B, int Ctot, int cp, int box_c, int box_b, int box_m, bool mn_major = false) {
  long long d[5] = {Ctot, B, 2, M, L};
  long long s[5] = {1, cp, (long long)B * cp, 2ll * B * cp, (long long)M * 2 * B * cp};
  int bx[5] = {box_c, box_b, 1, box_m, 1};
  return make_tmap(tm, base, 5, d, s, bx, mn_major);
}
static int weight_tmap(CUtensorMap* tm, const float* base, int Lw, int G, int Cig, int Cog, int cop, int box_o, int box_i, bool mn_major = false) {
  long long d[4] = {Cog, 2, Cig, (long long)Lw * G};
  long long s[4] = {1, cop, 2ll * cop, (long long)Cig * 2 * cop};
  int bx[4] = {box_o, 1, box_i, 1};
  return make_tmap(tm, base, 4, d, s, bx, mn_major);
}

static int fill_mix(const Plan* pl, int op, int B, int G, int Ci, int Co, MixParams* p) {
  B200_REQUIRE(B >= 1 && 32 % B == 0, "tcgen05 mix: batch %d must divide 32 (use precision fp32 otherwise)", B);
  B200_REQUIRE(G == 1 || ((Ci / G) % 4 == 0 && (Co / G) % 4 == 0), "tcgen05 mix: group slices (%d, %d channels) must be 16-byte aligned", Ci / G, Co / G);
  memset(p, 0, sizeof(*p));
  p->dense = pl->dense;
  p->L = pl->lmax; p->M = pl->mmax; p->B = B; p->G = G; p->Cig = Ci / G; p->Cog = Co / G;
  p->cpi = round_up(Ci, 4); p->cpo = round_up(Co, 4); p->cop = round_up(Co / G, 4);
  p->shared_w = (op == B200SHT_OP_SHARED);
  p->wl_stride = p->shared_w ? 0 : (long long)G * p->Cig * 2 * p->cop;
  p->Mt = 128 / B;
  return 0;
}

// Column tiling of a mix GEMM: one tile of up to 256 columns when that covers everything; otherwise equal tiles of at most 128
// columns (no half-empty last tile -- at C = 384 a 256 + 128 split wasted a quarter of the MMAs -- and 2 N <= 256 TMEM columns per
// complex accumulator leaves room for two accumulator sets, so the epilogue overlaps the next tile).
static void split_cols(int cols, int gran, int* N, int* n_nt) {
  if (cols <= 256) { *n_nt = 1; *N = round_up(cols, gran); return; }
  *n_nt = ceil_div(cols, 128);
  *N = round_up(ceil_div(cols, *n_nt), 32);   // the epilogues drain 32 columns at a time: a tile must not end inside a chunk
}

int mix_forward_umma(const Plan* pl, int op, const float* x, const void* w, const void* cbias, float* y, int B, int G, int Ci, int Co, cudaStream_t st) {
  MixParams p;
  int rc = fill_mix(pl, op, B, G, Ci, Co, &p);
  if (rc) return rc;
  p.out = y; p.cbias = static_cast<const float2*>(cbias);
  const int cols = p.Cog + ((p.cpo - Co) > 0 ? (p.cpo - Co) : 0);   // last group's tile also writes the zero padding
  split_cols(cols, 32, &p.N, &p.n_nt);
  p.nblk = p.N / 32;
  p.idesc = make_idesc(p.N, 0, 1, 0);
  p.idesc_neg = make_idesc(p.N, 0, 1, 1);
  p.offA_i = 16384; p.offB_r = 32768; p.offB_i = 32768 + 4096 * p.nblk;
  rc = spec_tmap(&p.tmX, x, p.L, p.M, B, Ci, p.cpi, 32, B, p.Mt);
  if (!rc) rc = weight_tmap(&p.tmW, static_cast<const float*>(w), p.shared_w ? 1 : p.L, G, p.Cig, p.Cog, p.cop, 32, 32, true);
  if (rc) return rc;
  pick_stages(&p, 32768 + 8192 * p.nblk, ceil_div(p.Cig, 32));
  p.tx_bytes = 2u * (uint32_t)(p.Mt * B) * 128 + 8192u * p.nblk;
  set_accumulators(&p, 2 * p.N);
  dim3 grid(ceil_div(p.M, p.Mt), p.n_nt * G, p.L);
  return launch<MixFwdTraits>(p, grid, st);
}

int mix_backward_umma(const Plan* pl, int op, const float* x, const void* w, const float* gy, float* gx, void* gw, void* gcbias, int B, int G,
                      int Ci, int Co, cudaStream_t st);

int mix_dgrad_umma(const Plan* pl, int op, const void* w, const float* gy, float* gx, int B, int G, int Ci, int Co, cudaStream_t st) {
  MixParams p;
  int rc = fill_mix(pl, op, B, G, Ci, Co, &p);
  if (rc) return rc;
  p.out = gx;
  const int cols = p.Cig + ((p.cpi - Ci) > 0 ? (p.cpi - Ci) : 0);
  split_cols(cols, 16, &p.N, &p.n_nt);
  p.idesc = make_idesc(p.N, 0, 0, 0);
  p.idesc_neg = make_idesc(p.N, 0, 0, 1);
  const uint32_t bb = (uint32_t)round_up(p.N * 128, 1024);
  p.offA_i = 16384; p.offB_r = 32768; p.offB_i = 32768 + bb;
  rc = spec_tmap(&p.tmX, gy, p.L, p.M, B, Co, p.cpo, 32, B, p.Mt);
  if (!rc) rc = weight_tmap(&p.tmW, static_cast<const float*>(w), p.shared_w ? 1 : p.L, G, p.Cig, p.Cog, p.cop, 32, p.N);
  if (rc) return rc;
  pick_stages(&p, 32768 + 2 * bb, ceil_div(p.Cog, 32));
  p.tx_bytes = 2u * (uint32_t)(p.Mt * B) * 128 + 2u * (uint32_t)p.N * 128;
  set_accumulators(&p, 2 * p.N);
  dim3 grid(ceil_div(p.M, p.Mt), p.n_nt * G, p.L);
  return launch<MixDgradTraits>(p, grid, st);
}

int mix_wgrad_umma(const Plan* pl, int op, const float* x, const float* gy, float* gw, int B, int G, int Ci, int Co, cudaStream_t st) {
  MixParams p;
  int rc = fill_mix(pl, op, B, G, Ci, Co, &p);
  if (rc) return rc;
  p.out = gw;
  split_cols(p.cop, 32, &p.N, &p.n_nt);
  p.nblk = p.N / 32;
  p.idesc = make_idesc(p.N, 1, 1, 0);
  p.idesc_neg = make_idesc(p.N, 1, 1, 1);
  p.offA_i = 16384; p.offB_r = 32768; p.offB_i = 32768 + 4096 * p.nblk;
  rc = spec_tmap(&p.tmX, x, p.L, p.M, B, Ci, p.cpi, 32, B, 32 / B, true);
  if (!rc) rc = spec_tmap(&p.tmX2, gy, p.L, p.M, B, Co, p.cpo, 32, B, 32 / B, true);
  if (rc) return rc;
  pick_stages(&p, 32768 + 8192 * p.nblk, 8);
  p.tx_bytes = 32768u + 8192u * p.nblk;
  set_accumulators(&p, 2 * p.N);
  dim3 grid(ceil_div(p.Cig, 128), p.n_nt * G, p.shared_w ? 1 : p.L);
  return launch<MixWgradTraits>(p, grid, st);
}

int mix_cbias_grad(const float* gy, void* gcb, int L, int M, int B, int Co, int dense, cudaStream_t st);  // mix.cu

int mix_backward_umma(const Plan* pl, int op, const float* x, const void* w, const float* gy, float* gx, void* gw, void* gcbias, int B, int G,
                      int Ci, int Co, cudaStream_t st) {
  int rc = 0;
  if (gx) rc = mix_dgrad_umma(pl, op, w, gy, gx, B, G, Ci, Co, st);
  if (!rc && gw) rc = mix_wgrad_umma(pl, op, x, gy, static_cast<float*>(gw), B, G, Ci, Co, st);
  if (!rc && gcbias) rc = mix_cbias_grad(gy, gcbias, pl->lmax, pl->mmax, B, Co, pl->dense, st);
  return rc;
}

}  // namespace b200sht
