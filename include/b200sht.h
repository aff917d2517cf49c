/*
 * b200sht -- C ABI of the B200-native spherical-harmonic hot path (RealSHT / InverseRealSHT / SpectralConv).
 *
 * Nothing equivalent exists in the reference: NVIDIA/makani has no native code and reaches this
 * arithmetic through the Python package torch-harmonics.  Each entry point below names the reference interface
 * it replaces (paths relative to /root/reference):
 *
 *   b200sht_plan_create          <- torch_harmonics.RealSHT.__init__ / InverseRealSHT.__init__ as constructed at
 *                                   makani/models/networks/sfnonet.py:792-805 (Legendre table + quadrature precompute)
 *   b200sht_sht_forward          <- RealSHT.forward            (call site makani/models/common/spectral_convolution.py:239)
 *   b200sht_sht_inverse          <- InverseRealSHT.forward     (call sites spectral_convolution.py:241,253)
 *   b200sht_sht_forward_adjoint  <- autograd backward of RealSHT.forward (rfft + einsum adjoints)
 *   b200sht_sht_inverse_adjoint  <- autograd backward of InverseRealSHT.forward
 *   b200sht_mix_forward/backward <- makani/models/common/contractions.py:19-54 (_contract_* einsums) and :62-151
 *   b200sht_spectral_conv_forward<- SpectralConv.forward       (spectral_convolution.py:213-264), one call
 *   b200sht_complex_relu_*       <- makani/models/common/activations.py:88-127 (ComplexReLU)
 *
 * Conventions
 *   - every function returns 0 on success, a negative b200sht_status otherwise; b200sht_last_error() gives text.
 *   - all data pointers are DEVICE pointers owned by the caller (e.g. the PyTorch caching allocator) unless a
 *     parameter is documented as host memory.  The library allocates only inside plans (Legendre table, FFT
 *     twiddles, TMA descriptors).
 *   - every launch takes the cudaStream_t to enqueue on (as void*), is asynchronous and never synchronises.
 *   - plans are immutable after creation and may be shared by concurrent calls on different streams.
 *   - no thread-local state except the last-error string.
 *
 * Internal ("packed") tensor formats -- opaque to callers that only use the *_sht_* / spectral_conv entry points,
 * documented in DESIGN.md section 3:
 *   latspec  float [mmax8][2][B*C][kp]          (after the longitude FFT;  kp = nlat rounded up to 8, mmax8 = mmax rounded up to 8)
 *   spec     float [lmax][mmax][2][B][cp]       (spectral coefficients;    cp = C rounded up to 4)
 */
#ifndef B200SHT_H
#define B200SHT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200SHT_OK = 0,
  B200SHT_ERR_INVALID = -1,      /* bad argument (shape, dtype, null pointer)            */
  B200SHT_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed                    */
  B200SHT_ERR_UNSUPPORTED = -3,  /* valid request this build cannot serve (e.g. FFT len) */
  B200SHT_ERR_NOMEM = -4
} b200sht_status;

typedef enum { B200SHT_F32 = 0, B200SHT_BF16 = 1 } b200sht_dtype;

/* arithmetic of the Legendre / channel-mix contractions */
typedef enum {
  B200SHT_PREC_FP32 = 0, /* fp32 FMA on CUDA cores (reference tests run with TF32 disabled)          */
  B200SHT_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM (reference training: allow_tf32) */
  B200SHT_PREC_FP32X3 = 2 /* fp32 operands on the tensor cores: Legendre stages as 3 x TF32 (hi.hi + hi.lo + lo.hi into one TMEM accumulator),
                             longitude transform and channel mix as in FP32.  Element errors stay inside rtol 1e-5 (atol = rtol max|ref|), relative
                             L2 ~ 1e-6 .. 8e-6 growing with nlat (the tensor core truncates its fp32 accumulator on every add; the CUDA-core FP32 mode
                             rounds to nearest and stays at ~ 3e-7): ~1.7 x the speed of FP32 at 721 x 1440.  Uses a per-device scratch buffer for the
                             operand residuals: issue calls of this mode from one stream per device.                                          */
} b200sht_precision;

typedef enum {
  B200SHT_OP_DHCONV = 0,      /* weight [G][Ci][Co][L]        contractions.py:23  */
  B200SHT_OP_DIAGONAL = 1,    /* weight [G][Ci][Co][L][M]     contractions.py:19  */
  B200SHT_OP_SEP_DHCONV = 2,  /* weight [G][Ci][L]            contractions.py:31  */
  B200SHT_OP_SEP_DIAGONAL = 3,/* weight [G][Ci][L][M]         contractions.py:27  */
  B200SHT_OP_SHARED = 4,      /* weight [Ci][Co]              contractions.py:62  (compl_mul2d_fwd)     */
  B200SHT_OP_LDEP = 5         /* weight [L][Ci][Co]           contractions.py:106 (compl_exp_mul2d_fwd) */
} b200sht_mix_op;

/* or-ed into `op` (mix) / `mode` (ComplexReLU): the packed spec operands store every (l, m) entry (no block triangle) */
#define B200SHT_DENSE_FLAG 0x100

typedef struct b200sht_plan b200sht_plan;

const char* b200sht_last_error(void);
int b200sht_version(void);

/* ---------------------------------------------------------------------------------------------- plan */
/* cost/quad_w: HOST arrays [nlat]: cos(colatitude) in row order (row 0 = north) and quadrature weights on [-1,1].
 * The Legendre table P[m][l][k] (orthonormal, optional Condon-Shortley phase) is built on the device in fp64
 * and stored as fp32 [mmax][lmax][kp]. */
int b200sht_plan_create(b200sht_plan** plan, int nlat, int nlon, int lmax, int mmax,
                        const double* cost, const double* quad_w, int csphase, void* stream);
/* Extended creation for the h x w model-parallel (distributed) SHT:
 *   m_offset : the plan's orders are m_offset .. m_offset + mmax - 1 (this rank's shard of the orders)
 *   flags & 1: FFT-only plan (no Legendre table): nlat is this rank's latitude count, quad_w its slice of the weights */
int b200sht_plan_create_ex(b200sht_plan** plan, int nlat, int nlon, int lmax, int mmax, int m_offset, int flags,
                           const double* cost, const double* quad_w, int csphase, void* stream);
int b200sht_plan_destroy(b200sht_plan* plan);
/* what: 0 nlat, 1 nlon, 2 lmax, 3 mmax, 4 kp, 5 table bytes, 6 tcgen05 path available (0/1), 7 m_offset,
 *       8 tensor-core longitude DFT available for this grid (0/1) */
int64_t b200sht_plan_query(const b200sht_plan* plan, int what);
/* device pointer to the fp32 table [mmax][lmax][kp] (for tests) */
const float* b200sht_plan_table(const b200sht_plan* plan);
/* copy the table into caller-owned device memory (mmax*lmax*kp floats) */
int b200sht_plan_copy_table(const b200sht_plan* plan, float* dst, void* stream);

/* ------------------------------------------------------------------------------ packed-format sizes */
int64_t b200sht_latspec_elems(const b200sht_plan* plan, int B, int C); /* floats in a latspec buffer */
int64_t b200sht_spec_elems(const b200sht_plan* plan, int B, int C);    /* floats in a spec buffer    */
int64_t b200sht_spec_elems_lm(int L, int M, int B, int C);

/* ------------------------------------------------------------------------------------ stage kernels */
/* Longitude analysis: real rows -> truncated half spectrum.
 *   X[m][p][r][k] = row_scale[k] * mode_scale[m] * sum_j x[r][k][j] exp(-2 pi i m j / nlon)
 * scale_mode 0: SHT forward     (row_scale = quad_w[k] * 2 pi / nlon, mode_scale = 1)
 * scale_mode 1: adjoint of irfft (row_scale = 1, mode_scale = 1 for m = 0 and Nyquist, 2 otherwise)
 * scale_mode | 2: TF32 precision: the output is rounded to the nearest TF32 value (it is the operand of a kind::tf32 GEMM) and,
 *                 for nlon = 8 * N2 <= 1520 and mmax <= 256, the transform itself runs on the tensor cores (radix-8 butterflies on
 *                 the CUDA cores x a [mmax/8 x nlon/16] DFT matrix as a kind::tf32 GEMM, csrc/dft.cu) */
int b200sht_fft_analysis(const b200sht_plan* plan, const void* x, int dtype, int B, int C,
                         float* latspec, int scale_mode, void* stream);
/* Longitude synthesis: truncated half spectrum -> real rows (+ optional per-channel bias, cast to dtype).
 * scale_mode 0: irfft(norm="forward") semantics (imaginary part of m=0 / Nyquist ignored)
 * scale_mode 1: adjoint of the scale_mode-0 analysis (row_scale = quad_w[k] 2 pi/nlon, modes m>0 halved)
 * scale_mode | 2: `latspec` is in the TILED layout written by b200sht_legendre_synthesis_tiled and the transform runs on the tensor
 *                 cores (TF32; radix-8 butterflies on the CUDA cores x a [mmax/8 x nlon/16] DFT matrix as a kind::tf32 GEMM, csrc/dft.cu).
 *                 Error unless b200sht_plan_query(plan, 8) == 1. */
int b200sht_fft_synthesis(const b200sht_plan* plan, const float* latspec, void* y, int dtype, int B, int C,
                          const float* bias, int scale_mode, void* stream);
/* Legendre analysis  spec[l][m][..] = sum_k P[m][l][k] latspec[m][..][k]   (l >= 32*floor(m/32)) */
int b200sht_legendre_analysis(const b200sht_plan* plan, const float* latspec, float* spec, int B, int C,
                              int precision, void* stream);
/* Legendre synthesis latspec[m][..][k] = sum_l P[m][l][k] spec[l][m][..] */
int b200sht_legendre_synthesis(const b200sht_plan* plan, const float* spec, float* latspec, int B, int C,
                               int precision, void* stream);
/* Legendre synthesis (TF32) into the TILED latspec layout the tensor-core longitude DFT consumes:
 *   latspec[r][k / 8][plane][m / 8][m % 8][k % 8]   (orders padded with zeros to a multiple of 8; same size as the standard layout)
 * i.e. the 16 KB that one 8-row tile of the DFT kernel reads are contiguous and arrive as 128-byte TMA rows.  Pair it with
 * b200sht_fft_synthesis(..., scale_mode | 2).  Requires b200sht_plan_query(plan, 8) == 1. */
int b200sht_legendre_synthesis_tiled(const b200sht_plan* plan, const float* spec, float* latspec, int B, int C, void* stream);
/* packed spec [L][M][2][B][cp] <-> torch complex64 [B*C][L][M] (exact zeros written for l < m).  These and the
 * mix / ComplexReLU entry points below depend only on the mode counts (L, M), not on a grid, so they take no plan. */
int b200sht_spec_unpack(int L, int M, const float* spec, void* coeffs, int B, int C, void* stream);
int b200sht_spec_pack(int L, int M, const void* coeffs, float* spec, int B, int C, void* stream);
/* same with an order offset (orders m_offset + m) and/or dense storage (every (l, m) entry stored: used for l/m-sharded spectra) */
int b200sht_spec_unpack_ex(int L, int M, int m_offset, int dense, const float* spec, void* coeffs, int B, int C, void* stream);
int b200sht_spec_pack_ex(int L, int M, int m_offset, int dense, const void* coeffs, float* spec, int B, int C, void* stream);
/* latspec [mmax][2][B*C][kp] <-> complex64 [B*C][nlat][mmax]: the layout the distributed lat<->lon transposes exchange */
int b200sht_latspec_unpack(const b200sht_plan* plan, const float* latspec, void* coeffs, int B, int C, void* stream);
int b200sht_latspec_pack(const b200sht_plan* plan, const void* coeffs, float* latspec, int B, int C, void* stream);

/* ------------------------------------------------------------------------- torch-harmonics boundary */
/* bytes of scratch the four calls below need for (B, C) */
int64_t b200sht_sht_workspace_bytes(const b200sht_plan* plan, int B, int C);
/* x [B*C][nlat][nlon] (dtype) -> coeffs complex64 [B*C][lmax][mmax] */
int b200sht_sht_forward(const b200sht_plan* plan, const void* x, int dtype, int B, int C, void* coeffs,
                        void* workspace, int precision, void* stream);
/* coeffs complex64 [B*C][lmax][mmax] -> y [B*C][nlat][nlon] (dtype) */
int b200sht_sht_inverse(const b200sht_plan* plan, const void* coeffs, void* y, int dtype, int B, int C,
                        void* workspace, int precision, void* stream);
/* gradient of sht_forward w.r.t. x given dL/dcoeffs (PyTorch complex-gradient convention) */
int b200sht_sht_forward_adjoint(const b200sht_plan* plan, const void* gcoeffs, void* gx, int dtype, int B, int C,
                                void* workspace, int precision, void* stream);
/* gradient of sht_inverse w.r.t. coeffs given dL/dy */
int b200sht_sht_inverse_adjoint(const b200sht_plan* plan, const void* gy, int dtype, int B, int C, void* gcoeffs,
                                void* workspace, int precision, void* stream);

/* -------------------------------------------------------------------------------------- channel mix */
/* weight re-layout: native torch parameter (complex64, shapes per b200sht_mix_op) -> packed
 * float [L][G][Ci/G][2][cop] (real and imaginary planes; cop = Co/G rounded up to 4; l-stride 0 for OP_SHARED).  Only the dense
 * operators (DHCONV, SHARED, LDEP) use a packed weight; the others read the native layout. */
int64_t b200sht_mix_weight_elems(int op, int L, int M, int G, int Ci, int Co);
int b200sht_mix_weight_pack(int op, const void* w_native, float* w_packed, int L, int G, int Ci, int Co, int precision, void* stream);
int b200sht_mix_weight_unpack(int op, const float* w_packed, void* w_native, int L, int G, int Ci, int Co, void* stream);

/* y[l][m][.][b][o] = sum_i x[l][m][.][b][i] * w[...]   on packed spec tensors (x: C = Ci, y: C = Co).
 * w: packed weight for dense ops, native complex64 for DIAGONAL / SEP_* ops.
 * cbias (OP_SHARED / OP_LDEP only, may be null): complex64 [Co] added to every mode (compl_muladd2d_fwd). */
int b200sht_mix_forward(int L, int M, int op, const float* x, const void* w, const void* cbias,
                        float* y, int B, int G, int Ci, int Co, int precision, void* stream);
/* 1 when b200sht_mix_forward / _backward run this shape on the tensor cores at `precision`, 0 when they run the fp32 CUDA-core
 * kernels: B200SHT_PREC_TF32 needs a dense operator, a batch that divides 32 and 16-byte aligned group slices; anything else is
 * served in fp32 -- the caller should then pack the weight with B200SHT_PREC_FP32 (no TF32 rounding) and may want to tell the user. */
int b200sht_mix_uses_tensor_cores(int op, int B, int G, int Ci, int Co, int precision);
/* gx = dL/dx (may be null), gw = dL/dw in the same format as w (may be null; overwritten, not accumulated),
 * gcbias complex64 [Co] (may be null). */
int b200sht_mix_backward(int L, int M, int op, const float* x, const void* w, const float* gy,
                         float* gx, void* gw, void* gcbias, int B, int G, int Ci, int Co, int precision, void* stream);

/* ------------------------------------------------------------------------------------- ComplexReLU */
/* mode 0 real, 1 cartesian, 2 modulus, 3 halfplane (activations.py:88-127) on a packed spec tensor.
 * bias: float [C] (modes 2,3; null -> 0).  In-place allowed (y == x). */
int b200sht_complex_relu_forward(int L, int M, int mode, const float* x, const float* bias,
                                 float negative_slope, float* y, int B, int C, void* stream);
int b200sht_complex_relu_backward(int L, int M, int mode, const float* x, const float* bias,
                                  float negative_slope, const float* gy, float* gx, float* gbias, int B, int C,
                                  void* stream);

/* ------------------------------------------------------------------------------ SpectralConv, one call */
typedef struct {
  int B, Cin, Cout, G;
  int op;          /* b200sht_mix_op */
  int dtype;       /* activation dtype of x / y / residual */
  int precision;   /* b200sht_precision */
} b200sht_conv_desc;
int64_t b200sht_spectral_conv_workspace_bytes(const b200sht_plan* fwd, const b200sht_plan* inv, const b200sht_conv_desc* d);
/* y = iSHT(W . SHT(x)) (+bias); residual (may be null) = iSHT(SHT(x)).  w: packed for dense ops, native otherwise.
 * spec_x_saved (may be null): packed spec buffer that receives SHT(x) for the backward pass. */
int b200sht_spectral_conv_forward(const b200sht_plan* fwd, const b200sht_plan* inv, const b200sht_conv_desc* d,
                                  const void* x, const void* w, const float* bias, void* y, void* residual,
                                  float* spec_x_saved, void* workspace, void* stream);
/* gy, gresidual (may be null) -> gx, gw (same format as w), gbias float [Cout] (may be null) */
int b200sht_spectral_conv_backward(const b200sht_plan* fwd, const b200sht_plan* inv, const b200sht_conv_desc* d,
                                   const void* gy, const void* gresidual, const float* spec_x_saved, const void* w,
                                   void* gx, void* gw, float* gbias, void* workspace, void* stream);

/* same, plus (both optional): gw_native receives the weight gradient re-laid-out to the parameter's native complex64 layout (dense operators)
 * and wgrad_ready_event (a cudaEvent_t) is recorded on `stream` as soon as the weight / bias gradients are final, i.e. BEFORE the two stages
 * that produce gx -- a data-parallel gradient all-reduce waiting on it from another stream overlaps them.  (The reference reaches the same
 * overlap through DDP gradient hooks, makani/mpu/mappings.py:398-406 init_gradient_reduction_hooks.) */
int b200sht_spectral_conv_backward_ex(const b200sht_plan* fwd, const b200sht_plan* inv, const b200sht_conv_desc* d,
                                      const void* gy, const void* gresidual, const float* spec_x_saved, const void* w,
                                      void* gx, void* gw, float* gbias, void* workspace, void* gw_native, void* wgrad_ready_event,
                                      void* stream);

/* sum over batch and latitude of latspec[m=0][re][b][c][k]: d(loss)/d(bias) when latspec = fft_analysis(gy, mode 1) */
int b200sht_bias_grad(const b200sht_plan* plan, const float* latspec, float* gbias, int B, int C, void* stream);

/* host-buffer convenience (end-to-end path for FFI users): copies x (pinned or pageable HOST memory) to the
 * device, runs b200sht_spectral_conv_forward, copies y back.  Synchronises the stream before returning. */
int b200sht_spectral_conv_forward_host(const b200sht_plan* fwd, const b200sht_plan* inv, const b200sht_conv_desc* d,
                                       const void* x_host, const void* w_device, const float* bias_device,
                                       void* y_host, void* stream);

/* ------------------------------------------------------------------- debug / CPU-testable entry points
 * These run the SAME __host__ __device__ code as the kernels on the host, so the FFT plan / butterflies / pair
 * splitting and the Legendre recurrence are unit-tested without a GPU.  All pointers are HOST pointers. */
/* direction 0: rows a, b (float[N]) -> half spectra Xa, Xb (float[2*mmax], interleaved), unscaled rfft.
 * direction 1: half spectra (float[2*mmax]) -> rows (float[N]), irfft(norm="forward") semantics. */
int b200sht_debug_fft_host(int N, int mmax, int direction, const float* in_a, const float* in_b, float* out_a, float* out_b);
/* the tensor-core DFT's factorisation (radix-8 stage, twiddles, index maps: the same __host__ __device__ code as the kernels) with the
 * GEMM summed in double on the host.  direction 0: in float[N] -> out float[2*mmax] (interleaved); direction 1: the reverse.
 * scale_mode / row_scale as for b200sht_fft_analysis / _synthesis (row_scale = the row's quadrature factor). */
int b200sht_debug_dft_host(int N, int mmax, int direction, int scale_mode, float row_scale, const float* in, float* out);
/* wait-time profile of the tensor-core DFT kernels (environment B200SHT_DFT_PROF=1): 16 counters of SM clocks, accumulated over all launches since the
 * last call and cleared by it (slots: see csrc/dft.cu).  All zeros when the profile is off.  Synchronises the device. */
int b200sht_debug_dft_profile(uint64_t* counters16);
/* ------------------------------------------------------------------ pointwise tail of the SFNO block */
/* Replaces torch.nn.InstanceNorm2d(num_features, eps, affine) (+ the nn.GELU that follows it) as built at makani/models/networks/sfnonet.py:618-620 and
 * applied at :385-406, and the bias + GELU of the 1x1-convolution stacks (makani/models/common/layers.py:537-760).  x, y, dy, dx: [B][C][hw] contiguous, float or
 * bf16 (dtype); gamma / beta / bias: float [C] (null: 1 / 0); stats: float [B*C][2] (mean, rstd), written by forward, read by backward; sums: float [B*C][2]
 * written by backward: per row sum g and sum g * xhat -- dbeta[c] / dgamma[c] are their sums over the batch; workspace: b200sht_pointwise_workspace_floats floats.
 * gelu != 0 fuses y = gelu(norm(x)) (exact erf GELU).  Statistics are biased (1 / hw), as InstanceNorm uses them. */
int64_t b200sht_pointwise_workspace_floats(int B, int C, int64_t hw);
int b200sht_instance_norm_forward(const void* x, void* y, const float* gamma, const float* beta, float* stats, float* workspace, int dtype, int B, int C,
                                  int64_t hw, float eps, int gelu, void* stream);
int b200sht_instance_norm_backward(const void* x, const void* dy, void* dx, const float* gamma, const float* beta, const float* stats, float* sums,
                                   float* workspace, int dtype, int B, int C, int64_t hw, int gelu, void* stream);
/* y = gelu(x + bias[c]);  dx = dy * gelu'(x + bias[c]), row_sums: float [B*C][2] with sum dx in [.][0] (dbias[c] = its sum over the batch; may be null) */
int b200sht_bias_gelu_forward(const void* x, const float* bias, void* y, int dtype, int B, int C, int64_t hw, void* stream);
int b200sht_bias_gelu_backward(const void* x, const float* bias, const void* dy, void* dx, float* row_sums, float* workspace, int dtype, int B, int C, int64_t hw,
                               void* stream);

/* Programmatic dependent launch between the tcgen05 kernels of a call sequence (prologue of kernel i+1 under the tail of kernel i; environment
 * B200SHT_PDL sets the initial value, default on).  Returns the previous setting.  Results do not depend on it. */
int b200sht_debug_set_pdl(int on);
/* Latitude chunks of the fused (longitude analysis -> Legendre analysis) pair inside b200sht_sht_forward / _inverse_adjoint and the
 * SpectralConv entry points at B200SHT_PREC_TF32: n > 1 forces n chunks, 1 switches chunking off, 0 restores the default (by size; the
 * environment variable B200SHT_LAT_CHUNKS sets the initial value).  Returns the previous setting.  Results are identical up to the
 * summation order of the Legendre sums. */
int b200sht_debug_set_lat_chunks(int n);
/* radices chosen for length N; returns the number of stages or a negative status */
int b200sht_debug_fft_plan(int N, int* radices, int max_radices);
/* table [mmax][lmax][nlat] (fp32) from cos(colatitude) cost[nlat] */
int b200sht_debug_table_host(int nlat, int lmax, int mmax, const double* cost, int csphase, float* table);

#ifdef __cplusplus
}
#endif
#endif /* B200SHT_H */
