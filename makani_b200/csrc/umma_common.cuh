// tcgen05 / TMA / mbarrier PTX wrappers, UMMA descriptors and the tensor-map helper shared by umma.cu (Legendre + mix engine)
// and dft.cu (tensor-core longitude DFT).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <cstring>

namespace b200sht {

// =============================================================================================== PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done;
}
// try_wait with a suspend-time hint: the warp sleeps in hardware (no issue slots) until the phase completes or ~`ns` elapse
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return done;
}
// bounded wait: a lost arrival traps (kernel error) instead of hanging the GPU.  The wall-clock check runs once per 256 wake-ups: the
// polling loop of the first version (clock64 + compare every iteration) cost the CUDA-core warps of dft.cu a third of their issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if ((++spins & 255u) == 0) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) {
        printf("b200sht: mbarrier timeout block (%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
        __trap();
      }
    }
  }
}

// wait of a warp that is not on the critical path (epilogue / MMA issuer of a CUDA-core-bound kernel): poll every `ns` nanoseconds
// instead of waking on every arrival, so that the waiting warps leave the issue slots to the working ones
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    __nanosleep(ns);
    if (mbar_try_wait(bar, parity)) return;
    if ((++spins & 255u) == 0) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) {
        printf("b200sht: mbarrier timeout block (%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Warp-synchronous forms: called by all 32 lanes of a converged warp, one elected lane issues.  Under `if (lane == 0)` the compiler wraps
// every UTCHMMA / UTCBAR in a five-instruction lane-serialising loop (the instruction is uniform, the predicate is not); with elect.sync it
// emits the instruction alone.
__device__ __forceinline__ void umma_tf32_ws(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_ws(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO (all >> 4), version = 1 (bit 46),
// layout type SWIZZLE_128B = 2 (bits 61..63).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= layout_type << 61;
  return d;
}
// Descriptor of the same operand `bytes` further on in shared memory: one 32-bit add on the address field.  (Rebuilding a descriptor
// from an address costs ~5 uniform-datapath instructions; a single thread issuing 16 MMAs per stage was spending most of its time there.)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) {
  return (d & 0xffffffff00000000ull) | (uint64_t)((uint32_t)d + (bytes >> 4));
}
// K-major operand tile: rows of 128 B (32 floats of K), 8-row groups 1024 B apart.  kstep selects the K = 8 slice (32 B).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kstep) { return make_smem_desc(tile + kstep * 32, 16, 1024, 2 /*SWIZZLE_128B*/); }
// MN-major operand tile: blocks of [32 K-rows][32 floats of M/N]; blocks `blk_bytes` apart; kstep selects 8 K-rows (1024 B).
// 32-bit MN-major operands use SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>): atoms of 4 K-rows x 128 B,
// so one K = 8 MMA spans two atoms SBO = 512 B apart; LBO = distance between 32-float M/N blocks.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kstep, uint32_t blk_bytes) {
  return make_smem_desc(tile + kstep * 1024, blk_bytes, 512, 1 /*SWIZZLE_128B_BASE32B*/);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32, fp32 accumulate, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int N, int a_mn_major, int b_mn_major, int negate_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)negate_a << 13) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ================================================================================================== host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// Encoded tensor maps are a pure function of (base, shape, strides, box, swizzle): in steady state the caching allocator hands the same
// buffers to every step, so the 2-3 driver encodes per launch are replaced by a lookup in a small per-thread direct-mapped cache.
struct TmapKey {
  const void* base;
  long long dims[5], strides[5];
  int box[5];
  int rank, kind;   // kind: 0 fp32 / 128-byte swizzle, 1 same with 32-byte atoms (MN-major), 2 fp32 rows, 3 bf16 rows, 4 fp32 segments, 5 bf16 segments
};
struct TmapCache {
  static constexpr int kSlots = 128;
  CUtensorMap maps[kSlots];
  TmapKey keys[kSlots];
  bool used[kSlots];
};
inline TmapCache& tmap_cache() {
  static thread_local TmapCache* c = [] { TmapCache* p = new TmapCache(); memset(p->used, 0, sizeof(p->used)); return p; }();
  return *c;
}
inline int tmap_slot(const TmapKey& k) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(TmapKey); ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return (int)(h % TmapCache::kSlots);
}
inline bool tmap_lookup(const TmapKey& k, CUtensorMap* tm, int* slot) {
  TmapCache& c = tmap_cache();
  *slot = tmap_slot(k);
  if (c.used[*slot] && memcmp(&c.keys[*slot], &k, sizeof(TmapKey)) == 0) { memcpy(tm, &c.maps[*slot], sizeof(CUtensorMap)); return true; }
  return false;
}
inline void tmap_store(const TmapKey& k, const CUtensorMap* tm, int slot) {
  TmapCache& c = tmap_cache();
  memcpy(&c.maps[slot], tm, sizeof(CUtensorMap));
  c.keys[slot] = k;
  c.used[slot] = true;
}

// fp32 tensor map, 128-byte swizzle.  dims[0] is the contiguous dimension; strides (in floats) for dims 1..rank-1.
// mn_major: the operand is read M/N-major by kind::tf32, which needs the 128-byte swizzle with 32-byte atoms
inline int make_tmap(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides, const int* box, bool mn_major = false) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.kind = mn_major ? 1 : 0;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.strides[i] = i ? strides[i] : 1; key.box[i] = box[i]; }
  int slot = 0;
  if (tmap_lookup(key, tm, &slot)) return 0;
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return B200SHT_ERR_UNSUPPORTED; }
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a current context.  A thread that has made no runtime call yet (PyTorch's autograd
  // thread entering a backward whose first action is this encode) has none -> CUDA_ERROR_INVALID_CONTEXT (201).  Bind the primary context once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = 1; }
  for (int i = 1; i < rank; ++i) {
    gs[i - 1] = (cuuint64_t)strides[i] * 4;
    if (gs[i - 1] % 16 != 0) { set_error("tensor map stride %lld floats is not 16-byte aligned", strides[i]); return B200SHT_ERR_INVALID; }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) { set_error("tensor map base is not 16-byte aligned"); return B200SHT_ERR_INVALID; }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank); return B200SHT_ERR_CUDA; }
  tmap_store(key, tm, slot);
  return 0;
}


// 2-D tensor map over a row-major matrix of fp32 or bf16 elements without swizzle: box rows land densely (box_cols * esize bytes apart)
inline int make_tmap_rows(CUtensorMap* tm, const void* base, bool bf16, long long cols, long long rows, int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return B200SHT_ERR_UNSUPPORTED; }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  const int es = bf16 ? 2 : 4;
  cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gs[1] = {(cuuint64_t)cols * es};
  cuuint32_t bx[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, el[2] = {1, 1};
  if (gs[0] % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (box_cols * es) % 16 != 0) {
    set_error("tensor map (rows): base / row pitch / box row not 16-byte aligned");
    return B200SHT_ERR_INVALID;
  }
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gd, gs, bx, el,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (rows) failed (%d)", (int)r); return B200SHT_ERR_CUDA; }
  return 0;
}

}  // namespace b200sht
