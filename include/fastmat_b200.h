/*
 * fastmat_b200.h -- C-ABI of the B200-native structured-matrix apply path.
 *
 * This is the drop-in boundary for the hot path of EMS-TU-Ilmenau/fastmat: the `_forward/_backward[C]`
 * overrides of Fourier, Circulant, Toeplitz, Hadamard, Kron(Fourier...), Partial, Diag that
 * `Matrix.forward/backward` dispatches to (reference: fastmat/Matrix.pyx:1844-1910, :1937-2007, override
 * points declared at fastmat/Matrix.pxd:143-147).  One immutable *plan* object per matrix instance owns the
 * device-side constants (twiddles, spectrum, chirp, index vectors) the reference keeps as ndarray attributes;
 * `fmb_plan_apply` mirrors the cython-style `_forwardC(arrX, arrRes, typeX, typeRes)` contract
 * (fastmat/Matrix.pyx:1819-1829): the caller owns and pre-allocates the output, the callee writes it.
 *
 * Conventions
 *   - plain C, no exceptions: every function returns 0 (FMB_OK) or a negative fmb_status; the message of the
 *     last failure on the calling thread is available from fmb_last_error().  The Python layer maps
 *     FMB_ERR_VALUE -> ValueError and FMB_ERR_TYPE -> TypeError (same exception classes as the reference,
 *     fastmat/Matrix.pyx:1772-1782, fastmat/core/types.pyx:159-161).
 *   - x / y / workspace are DEVICE pointers; generator vectors handed to *_plan_create are HOST pointers and are
 *     copied (the reference copies its defining vectors too: fastmat/Diag.pyx:88, fastmat/Circulant.pyx:88).
 *   - arrays are 2-D (n, M) with explicit element strides (row stride, column stride), so both the fastmat-native
 *     column-major layout (row stride 1) and torch's row-major layout (column stride 1) are taken as they are;
 *     x is never modified (fastmat/inspect/test.py:334-338); x and y must not alias.
 *   - apply is asynchronous on the CUDA stream passed in (a cudaStream_t cast to void*; NULL = default stream),
 *     never synchronises the host, allocates nothing (scratch comes from the caller: fmb_plan_workspace_bytes)
 *     and is safe to call concurrently on one plan from several host threads / streams with distinct workspaces.
 *   - dtype codes are fastmat's ftype ids (fastmat/core/types.pxd:45-55).
 */
#ifndef FASTMAT_B200_H
#define FASTMAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmb_plan fmb_plan;

typedef enum fmb_dtype {          /* fastmat/core/types.pxd:45-55 */
    FMB_INT8 = 0, FMB_INT16 = 1, FMB_INT32 = 2, FMB_INT64 = 3,
    FMB_FLOAT32 = 4, FMB_FLOAT64 = 5, FMB_COMPLEX64 = 6, FMB_COMPLEX128 = 7
} fmb_dtype;

typedef enum fmb_status {
    FMB_OK = 0,
    FMB_ERR_VALUE = -1,           /* -> ValueError  (bad sizes / arguments)            */
    FMB_ERR_TYPE = -2,            /* -> TypeError   (unsupported dtype combination)    */
    FMB_ERR_CUDA = -3,            /* -> RuntimeError (CUDA runtime failure)            */
    FMB_ERR_NOTIMPL = -4,         /* -> NotImplementedError                            */
    FMB_ERR_WORKSPACE = -5        /* -> ValueError  (workspace too small)              */
} fmb_status;

typedef enum fmb_direction { FMB_FORWARD = 0, FMB_BACKWARD = 1 } fmb_direction;

typedef enum fmb_kind {
    FMB_KIND_FOURIER = 1, FMB_KIND_CIRCULANT = 2, FMB_KIND_TOEPLITZ = 3, FMB_KIND_HADAMARD = 4,
    FMB_KIND_DIAG = 5, FMB_KIND_PARTIAL = 6, FMB_KIND_KRON_FOURIER = 7
} fmb_kind;

typedef struct fmb_plan_info {
    int32_t kind;                 /* fmb_kind                                                               */
    int64_t num_rows, num_cols;   /* shape of the matrix the plan applies                                   */
    int64_t inner_size;           /* FFT length actually transformed (circulant embedding / Bluestein L)    */
    int64_t bluestein;            /* Fourier: the reference's `_numL` decision (0 = plain FFT)              */
    int32_t passes_fwd;           /* kernel launches per slab of columns for forward                        */
    int32_t slab_cols;            /* columns pushed through all passes together (L2-resident intermediate)  */
} fmb_plan_info;

/* ---- library ------------------------------------------------------------------------------------------ */
const char *fmb_last_error(void);
int fmb_version(void);
int fmb_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes);

/* ---- host-side planner: fastmat/core/cmath.pyx:88-153 (_findOptimalFFTSize), :156-214 (_getFFTComplexity).
 *      Bit-identical to the reference including its C-float rounding (it decides padded sizes).             */
int64_t fmb_find_optimal_fft_size(int64_t order, int max_stage);
float fmb_fft_complexity(int64_t n);

/* ---- host-side LFSR logic of LFSRCirculant (integer, exact): fastmat/LFSRCirculant.pyx:28-48 (lfsrGenStep /
 *      lfsrTapStep), :196-222 (order and period, with the constructor's ValueErrors), :277-314 (states, vecC),
 *      :316-395 (generator / tap address sequences of _core).  fmb_lfsr_period returns FMB_ERR_VALUE (< 0) with
 *      the reference's message in fmb_last_error() for an invalid register.  Output pointers may be NULL.      */
int fmb_lfsr_order(uint32_t polynomial);
int64_t fmb_lfsr_period(uint32_t polynomial, uint32_t start);
int fmb_lfsr_sequences(uint32_t polynomial, uint32_t start, int64_t n, uint32_t *gen_states, uint32_t *tap_states,
                       int8_t *vec_c);

/* ---- plan constructors (one per reference class on the path) -------------------------------------------- */

/* Fourier(order, optimize, maxStage): fastmat/Fourier.pyx:63-161.  forward = unnormalised DFT along axis 0
 * (:208-214), backward = conj(F conj(x)) (:235-238).  Non-smooth orders run a chirp-z (Bluestein) transform. */
int fmb_fourier_plan_create(fmb_plan **out, int64_t order, int optimize, int max_stage);

/* Circulant(c, optimize, maxStage), one level: fastmat/Circulant.pyx:63-137, 218-221.  `c_host` = first column,
 * n complex128 values.  forward y = ifft(fft(c) . fft(x)); backward uses conj(fft(c)).                       */
int fmb_circulant_plan_create(fmb_plan **out, const void *c_host, int64_t n, int optimize, int max_stage);

/* Toeplitz(vecC, vecR, optimize, maxStage), one level: fastmat/Toeplitz.pyx:79-322.  vecC = first column (n),
 * vecR = first row without element (0,0), stored reversed as in the reference (:730-732), m-1 values; both
 * complex128 on the host.  Zero padding / truncation is folded into the kernel's loads and stores.            */
int fmb_toeplitz_plan_create(fmb_plan **out, const void *vec_c_host, int64_t n, const void *vec_r_host,
                             int64_t m_minus_1, int optimize, int max_stage);

/* Hadamard(order): fastmat/Hadamard.pyx:102-131, _forwardC :164-230.  Natural-order WHT, unnormalised,
 * arithmetic in the array's own dtype (integers wrap), stage order as the reference -> bit-exact all dtypes. */
int fmb_hadamard_plan_create(fmb_plan **out, int order);

/* Diag(vecD): fastmat/Diag.pyx:73-167 -> core/cmath.pyx:958-1012.  y[n,m] = x[n,m] * d[n] (backward conj(d)).
 * d_host has n elements of `dtype`.                                                                          */
int fmb_diag_plan_create(fmb_plan **out, const void *d_host, int dtype, int64_t n);

/* Partial(mat, rows, cols) index plumbing: fastmat/Partial.pyx:268-294.  The plan is the exact gather/scatter
 * pair: forward  y[r,:] = x[idx[r],:]  (gather: num_sel x M from num_total x M),
 *       backward y = 0; y[idx[r],:] = x[r,:]  (scatter into num_total x M).   idx_host: num_sel int64 values. */
int fmb_partial_plan_create(fmb_plan **out, const int64_t *idx_host, int64_t num_sel, int64_t num_total);

/* Kron(Fourier(d0), ..., Fourier(dk-1)): fastmat/Kron.pyx:267-341 restricted to Fourier factors = the N-D DFT
 * of the row-major reshaped column.  2 <= ndims <= 3 in this release.                                        */
int fmb_kron_fourier_plan_create(fmb_plan **out, const int64_t *dims, int ndims);

/* ---- plan use ------------------------------------------------------------------------------------------ */
int fmb_plan_info_get(const fmb_plan *plan, fmb_plan_info *info);

/* Scratch bytes apply() needs for M columns of `dtype_out` (0 is a valid answer).                           */
int64_t fmb_plan_workspace_bytes(const fmb_plan *plan, int direction, int64_t M, int dtype_in, int dtype_out);

/* y = A x (FMB_FORWARD) or y = A^H x (FMB_BACKWARD).  x: (num_cols or num_rows, M); y likewise.
 * Mirrors _forwardC / _backwardC (fastmat/Matrix.pyx:1819-1829, :1912-1922).                                  */
int fmb_plan_apply(const fmb_plan *plan, int direction,
                   const void *x, int64_t x_row_stride, int64_t x_col_stride,
                   void *y, int64_t y_row_stride, int64_t y_col_stride,
                   int64_t M, int dtype_in, int dtype_out,
                   void *workspace, int64_t workspace_bytes, void *cuda_stream);

int fmb_plan_destroy(fmb_plan *plan);

/* ---- small exact helpers used by the class layer (conjugate, cast; fastmat/core/cmath.pyx:744-840) -------- */
int fmb_conjugate(const void *x, int64_t x_row_stride, int64_t x_col_stride,
                  void *y, int64_t y_row_stride, int64_t y_col_stride,
                  int64_t n, int64_t M, int dtype, void *cuda_stream);

/* y = (dtype_out) x, element-wise widening cast (int/float -> float/complex); the class layer uses it for the dtype
 * promotion of fastmat/Matrix.pyx:1799-1808 when an operator cannot read the input dtype directly. */
int fmb_cast(const void *x, int64_t x_row_stride, int64_t x_col_stride, int dtype_in,
             void *y, int64_t y_row_stride, int64_t y_col_stride, int dtype_out,
             int64_t n, int64_t M, void *cuda_stream);

/* Fused proximal-gradient update of ISTA / FISTA on `count` contiguous elements (replaces the numpy expression chain of
 * fastmat/algorithms/ISTA.py:150-158 + softThreshold :113-123):
 *     step = x - num_l * grad;  m = max(|step| - alpha, 0);  x_out = m / (m + alpha) * step
 * grad == NULL: step = x (plain soft threshold); step_out == NULL: the step is not stored.  dtype: float32/64, complex64/128. */
int fmb_ista_step(const void *x, const void *grad, void *step_out, void *x_out, int64_t count,
                  double num_l, double alpha, int dtype, void *cuda_stream);

/* Atom selection of OMP (fastmat/algorithms/OMP.pyx:196-199: np.argmax(np.abs(C^H r), axis=0)) in one sweep over a
 * column-major (rows x cols) device array (element (r, c) at x[c * col_stride + r]): out_index[c] = the row of the largest
 * magnitude of column c, the first one on ties (as np.argmax).  The magnitude array is never formed.  dtype: float32/64,
 * complex64/128; cols <= 65535; workspace: fmb_abs_argmax_workspace_bytes(cols) device bytes. */
int64_t fmb_abs_argmax_workspace_bytes(int64_t cols);
int fmb_abs_argmax(const void *x, int64_t rows, int64_t cols, int64_t col_stride, int dtype, int64_t *out_index,
                   void *workspace, int64_t workspace_bytes, void *cuda_stream);

/* Batched Gram-Schmidt step of OMP's incremental QR (replaces the explicit pseudo-inverse updates of
 * fastmat/algorithms/OMP.pyx:211-241): `batches` independent problems; problem l has k orthonormal rows
 * q[l * q_batch_stride + j * q_row_stride + (0..n-1)] and a vector v[l * v_batch_stride + (0..n-1)] (element strides):
 *     fmb_gs_project :  coef[l * coef_batch_stride + j] = sum_e conj(q[l, j, e]) * v[l, e]       (j < k)
 *     fmb_gs_subtract:  v[l, e] -= sum_j coef[l, j] * q[l, j, e]                                 (in place)
 * dtype: float32/64, complex64/128; k <= 256, batches <= 65535. */
int fmb_gs_project(const void *q, int64_t q_batch_stride, int64_t q_row_stride, int k, const void *v, int64_t v_batch_stride,
                   int64_t n, int64_t batches, void *coef, int64_t coef_batch_stride, int dtype, void *cuda_stream);
int fmb_gs_subtract(const void *q, int64_t q_batch_stride, int64_t q_row_stride, int k, void *v, int64_t v_batch_stride,
                    int64_t n, int64_t batches, const void *coef, int64_t coef_batch_stride, int dtype, void *cuda_stream);

/* Count of kernel launches issued by this library in the calling process (for bench.py's gpu_launches). */
int64_t fmb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FASTMAT_B200_H */
