/* capital_b200 -- C ABI of the B200-native CholInv / CholeskyQR2 hot path.
 *
 * Drop-in boundary for the factorization entry points of tbennun/capital (a header-only C++14
 * template library; it has no FFI of its own, so each entry point below names the reference
 * template it replaces, paths relative to the reference root).  Plain pointers and sizes only;
 * no torch / C++ types.  All matrices are FP64, column-major, in the reference's element-cyclic
 * layout (src/matrix/matrix.hpp:6-19, src/matrix/structure.h:13,37-39):
 *   global (row gy, col gx) lives on process (x = gx mod d, y = gy mod d) at local (col gx/d, row gy/d);
 *   `rect` local block: element (col i, row j) at i*ld + j;  packed `uppertri`: (i, j<=i) at i(i+1)/2 + j.
 *
 * Pointers passed to the compute entry points may be device pointers (resident HBM, the fast
 * path) or host pointers (the library stages them through pinned buffers: this is the
 * "reference-facing" call a C++ caller of the reference would make).  There is NO CPU fallback:
 * every entry point fails with CAPITAL_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CAPITAL_B200_H
#define CAPITAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

typedef struct capital_ctx capital_ctx;

typedef enum {
  CAPITAL_OK = 0,
  CAPITAL_ERR_INVALID = 1,     /* bad argument (the reference would assert: cholinv.hpp:9,81) */
  CAPITAL_ERR_CUDA = 2,        /* CUDA runtime / driver failure, or no sm_100 device */
  CAPITAL_ERR_NOT_SPD = 3,     /* non-positive pivot in a base case (reference drops LAPACKE info: lapack/interface.hpp:39) */
  CAPITAL_ERR_COMM = 4,        /* NCCL failure */
  CAPITAL_ERR_UNSUPPORTED = 5  /* grid / policy combination outside the hot path */
} capital_status_t;

/* topo::square / topo::rect public members (src/util/topology.h:62-64,140-142). */
typedef struct {
  int size, rank;          /* world */
  int c, d;                /* replication depth, face edge (square: c x d x d; rect: c x d x c) */
  int x, y, z;             /* process column, process row, layer */
  int layout, num_chunks;  /* kept for signature parity; layout 0 only, num_chunks ignored on NVSwitch */
} capital_grid_t;

/* cholesky::cholinv<...>::info user members (src/alg/cholesky/cholinv/cholinv.h:25-30). */
typedef struct {
  int64_t complete_inv;  /* 0: skip the top-level Rinv12 block (cholinv.hpp:147) */
  int64_t split;         /* recursion split shift (>0; 1 = halves) */
  int64_t bc_mult_dim;   /* base-case depth factor (cholinv.hpp:15-18) */
  char dir;              /* must be 'U' (cholinv.hpp:9) */
} capital_cholinv_args_t;

/* serialize policy of the outputs: policy::cholinv::Serialize -> packed uppertri, NoSerialize -> rect. */
typedef enum { CAPITAL_RECT = 0, CAPITAL_UPPERTRI_PACKED = 1 } capital_structure_t;

/* counters for bench/tests: how many of OUR kernels were launched since the last reset */
typedef struct {
  int64_t kernel_launches;
  int64_t gemm_launches;
  int64_t leaf_launches;
  int64_t h2d_bytes, d2h_bytes;
  double gemm_flops;  /* flops executed by the tensor-core GEMM kernel (2*m*n*k over computed tiles) */
} capital_counters_t;

/* ---- grid helpers -------------------------------------------------------------------------- */
/* topo::square(comm, c, layout, num_chunks) rank -> (x,y,z) map, topology.h:67-95 (layout 0). */
capital_status_t capital_grid_square(int size, int rank, int c, int layout, int num_chunks, capital_grid_t* out);
/* topo::rect(comm, c, layout, num_chunks), topology.h:16-51. */
capital_status_t capital_grid_rect(int size, int rank, int c, int layout, int num_chunks, capital_grid_t* out);
/* base-case global dimension derived from bc_mult_dim, cholinv.hpp:15-18. */
int64_t capital_cholinv_bc_dimension(int64_t local_dim, int c, int d, int64_t bc_mult_dim);

/* ---- context ------------------------------------------------------------------------------- */
/* One context per process / GPU.  `stream` is a cudaStream_t (NULL = library-owned stream). */
capital_status_t capital_create(capital_ctx** ctx, const capital_grid_t* grid, int device, void* stream);
/* Multi-GPU: join the NCCL clique.  `nccl_unique_id` = 128 bytes of ncclUniqueId produced by
 * capital_comm_unique_id on rank 0 and broadcast by the caller (replaces MPI_Comm_split in
 * topology.h:84-94).  Not needed when grid.size == 1. */
capital_status_t capital_comm_unique_id(void* out128);
capital_status_t capital_comm_init(capital_ctx* ctx, const void* nccl_unique_id);
/* Same, bootstrapped through a caller-supplied host allgather instead of NCCL (MPI_Allgather in an MPI program:
 *   int ag(void* user, const void* send, void* recv, int64_t bytes) { return MPI_Allgather(send, bytes, MPI_BYTE, recv, bytes, MPI_BYTE, *(MPI_Comm*)user); }
 * ).  The library only exchanges small blobs (IPC handles) through it, at init and when its peer-visible arena has to grow;
 * matrix data never goes through it.  Must return 0 on success; recv holds size * bytes. */
typedef int (*capital_allgather_fn)(void* user, const void* send, void* recv, int64_t bytes_per_rank);
capital_status_t capital_comm_init_host(capital_ctx* ctx, capital_allgather_fn allgather, void* user);
/* How this rank's streams wait for a flag written by a peer GPU: 0 = stream memory-op wait, 1 = memory-op wait followed by a flush
 * of outstanding remote writes (CU_STREAM_WAIT_VALUE_FLUSH; the default where the device supports it), 2 = one-warp kernel spinning
 * on ld.acquire.sys, -1 = the context has not joined a clique.  Override: env CAPITAL_PEER_WAIT = memop | flush | kernel. */
int capital_peer_wait_mode(const capital_ctx* ctx);
/* Switch the wait flavour between calls (measurement: bench.py times both on the same box).  CAPITAL_ERR_UNSUPPORTED when the device
 * cannot do it (mode 1 without flush support). */
capital_status_t capital_set_peer_wait_mode(capital_ctx* ctx, int mode);
void capital_destroy(capital_ctx* ctx);
const char* capital_last_error(const capital_ctx* ctx);
capital_status_t capital_get_counters(const capital_ctx* ctx, capital_counters_t* out);
capital_status_t capital_reset_counters(capital_ctx* ctx);
capital_status_t capital_synchronize(capital_ctx* ctx);
/* Rebind the context to another caller stream (cudaStream_t): later calls are enqueued on it, ordered after everything already
 * enqueued on the previous stream.  The Python mirror calls it whenever torch's current stream changed. */
capital_status_t capital_set_stream(capital_ctx* ctx, void* stream);
/* policy::cholinv::FlushIntermediates (cholinv/policy.h:85-156): release every work buffer the context holds (the next factor call
 * re-allocates: SaveIntermediates semantics -- keep them between calls -- are the default, cholinv/policy.h:20-83). */
capital_status_t capital_release_workspace(capital_ctx* ctx);
/* time (ms) between two library-recorded CUDA events bracketing the last factor call, on its stream */
capital_status_t capital_last_factor_ms(const capital_ctx* ctx, float* ms);

/* Per-launch CUDA-event timing of the dominant kernel (the 128x128 DMMA GEMM) on the stream it is launched on:
 * begin arms it; end synchronizes and returns the summed launch durations, the algorithmic flops of those
 * launches (structure exploited) and their count.  Used by bench.py for the roofline line. */
capital_status_t capital_profile_begin(capital_ctx* ctx);
/* enabled = 0 runs the recursion on a single stream (no deferred-update overlap): isolates per-kernel durations. */
capital_status_t capital_set_overlap(capital_ctx* ctx, int enabled);
capital_status_t capital_profile_end(capital_ctx* ctx, double* kernel_ms, double* kernel_flops, int64_t* launches);
/* Timeline of the schedule (profiling aid; the image has no nsys): between begin and end every launch of the library is bracketed by
 * CUDA events on its own stream.  end synchronizes the device and writes 8 doubles per launch: stream id (0 caller, 1 critical chain,
 * 2..4 deferred (recursion depth 0..2), 5..9 push streams, 10 copy-in, 11 copy-out), kind (1 big GEMM, 2 small GEMM, 3 cluster base case, 4 leaf, 5 flag wait,
 * 6 flag signal, 7 peer DMA, 8 layout kernel), start ms, end ms, three kind-specific numbers (GEMM: m, n, k), 0. */
capital_status_t capital_timeline_begin(capital_ctx* ctx);
capital_status_t capital_timeline_end(capital_ctx* ctx, double* out, int64_t cap_records, int64_t* n_records);
/* FP64 tensor-pipe ceiling of this device right now: a register-resident DMMA.8x8x4 loop on every SM (~50 ms), CUDA-event timed.
 * The denominator of bench.py's roofline fraction. */
capital_status_t capital_probe_dmma_f64(capital_ctx* ctx, double* tflops, double* ms);

/* Test facility, no device needed: records the synchronisation-relevant operations (flag waits / signals, fused products, events,
 * peer DMA, arena windows read / written) that rank `grid->rank` would enqueue for two consecutive cholinv::factor calls, 8 int64
 * per record (kind, stream, a .. f); tests replay the traces of all ranks of a grid to prove the flag protocol cannot deadlock. */
capital_status_t capital_dist_trace_cholinv(const capital_grid_t* grid, int64_t n_global, const capital_cholinv_args_t* args,
                                            int64_t* out, int64_t cap_records, int64_t* n_records);

/* ---- generators (device kernels; bit-exact with the reference's drand48-based ones) ---------- */
/* matrix::distribute_symmetric(x, y, d, d, key, diagonallyDominant) -- structure.hpp:69-103. */
capital_status_t capital_distribute_symmetric_f64(capital_ctx* ctx, double* A_local, int64_t n_global,
                                                  int diagonally_dominant);
/* matrix::distribute_random(x, y, c, d, key) -- structure.hpp:106-129 (rows over d, columns over c). */
capital_status_t capital_distribute_random_f64(capital_ctx* ctx, double* A_local, int64_t m_global,
                                               int64_t n_global, int64_t key);

/* ---- CholInv ------------------------------------------------------------------------------- */
/* cholesky::cholinv<SP,IP,BP>::factor(A, args, topo) -- cholinv.hpp:6-28 (+ invoke :87-165, base
 * case policy.h:160-224 semantics: zeros on local-diagonal slots of ranks with y > x).
 * A_local: rect local block, ld = ceil(n/d), never modified.  R_local / Rinv_local: caller-owned,
 * packed upper (L(L+1)/2) or rect (L*L, lower part zero), identical on all c layers. */
capital_status_t capital_cholinv_factor_f64(capital_ctx* ctx, const double* A_local, int64_t n_global,
                                            const capital_cholinv_args_t* args, capital_structure_t out_structure,
                                            double* R_local, double* Rinv_local);
/* cholesky::validate<Alg>::residual -- test/cholesky/validate.hpp:7-49:
 * sqrt(sum_upper (R^T R - A)^2) / sqrt(sum_upper A^2), computed on the device(s). */
capital_status_t capital_cholinv_residual_f64(capital_ctx* ctx, const double* A_local, int64_t n_global,
                                              capital_structure_t structure, const double* R_local, double* residual);

/* ---- CholeskyQR2 --------------------------------------------------------------------------- */
/* qr::cacqr<SP,IP>::factor(A, args, topo) -- cacqr.hpp:217-248; 1D path (c == 1): invoke_1d :172-193,
 * sweep_1d :5-29, Gram allreduce policy.h:78-85.  A_local: (m/d) x n rect.  Q_local same shape;
 * R_local: n x n packed upper (or rect), replicated on every rank.  num_iter: 1 = CQR, 2 = CQR2. */
capital_status_t capital_cacqr_factor_f64(capital_ctx* ctx, const double* A_local, int64_t m_global, int64_t n_global,
                                          int num_iter, const capital_cholinv_args_t* ci_args,
                                          capital_structure_t r_structure, double* Q_local, double* R_local);
/* qr::validate<Alg>::residual / orthogonality -- test/qr/validate.hpp:7-52. */
capital_status_t capital_cacqr_residual_f64(capital_ctx* ctx, const double* A_local, int64_t m_global, int64_t n_global,
                                            const double* Q_local, capital_structure_t r_structure,
                                            const double* R_local, double* residual, double* orthogonality);

/* ---- SUMMA ------------------------------------------------------------------------------------ */
/* matmult::summa::invoke(A, B, C, topo, gemm{Trans, NoTrans, alpha, beta}) -- summa.hpp:6-44 in the T*N form the validators
 * and syrk_internal use (test/cholesky/validate.hpp:35, summa.hpp:143-145):  C = alpha * A^T B + beta * C  with A (k x m),
 * B (k x n), C (m x n) all element-cyclic over the d x d face (local blocks k/d x m/d, k/d x n/d, m/d x n/d, column-major,
 * replicated over the c layers; d must divide m, n, k).  Host or device pointers. */
capital_status_t capital_summa_gemm_tn_f64(capital_ctx* ctx, int64_t m_global, int64_t n_global, int64_t k_global, double alpha,
                                           const double* A_local, const double* B_local, double beta, double* C_local);

/* ---- leaf-engine seam (the reference's designated swap point, blas/engine.h:7-8) ------------- */
/* blas::engine::_gemm (blas/interface.hpp:43-59) restricted to the T*N form the hot path executes
 * (summa.hpp:143-145): C[m x n] = alpha * A^T B + beta * C, A is k x m, B is k x n, all column-major
 * DEVICE pointers.  `flags` = OR of CAPITAL_GEMM_* below (structure hints that skip zero tiles). */
enum {
  CAPITAL_GEMM_A_UPPER = 1,  /* A[k,i] == 0 for k > i   (trmm Left/Upper/Trans, summa.hpp:64) */
  CAPITAL_GEMM_A_LOWER = 2,  /* A[k,i] == 0 for k < i */
  CAPITAL_GEMM_B_UPPER = 4,  /* B[k,j] == 0 for k > j */
  CAPITAL_GEMM_B_LOWER = 8,  /* B[k,j] == 0 for k < j */
  CAPITAL_GEMM_C_UPPER = 16  /* only tiles touching i <= j are computed/stored (syrk 'U') */
};
capital_status_t capital_blas_gemm_tn_f64(capital_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha,
                                          const double* A, int64_t lda, const double* B, int64_t ldb,
                                          double beta, double* C, int64_t ldc, int flags);
/* EXPERIMENTAL, OFF BY DEFAULT -- BASELINE config 5 ("FP32/TF32 Cholesky, mixed-precision trailing update with FP64 panel").
 * The reference has no float BLAS path (src/blas/interface.hpp:43-97 is double only): this is an extension of the blas::engine seam,
 * not a replacement of a reference entry point.  capital_blas_gemm_tn_tf32: the product of capital_blas_gemm_tn_f64 (FP64 operands
 * and result, only the CAPITAL_GEMM_C_UPPER flag) computed on the TF32 tensor cores (tcgen05.mma.kind::tf32, accumulator in TMEM);
 * passes = 1: operands rounded to TF32 (relative error ~ 5e-4 per product), passes = 3: operands split hi + lo (FP32-class).
 * capital_set_trailing_precision(ctx, 0 | 1 | 3): cholinv::factor runs its trailing updates A22 -= R12^T R12 (cholinv.hpp:131-134,
 * summa::syrk) in that mode; base cases, R12, the inverse and every other product stay FP64.  Single GPU and c = 1 grids.
 * capital_tf32_stats: launches / flops of the TF32 kernel since capital_create. */
capital_status_t capital_blas_gemm_tn_tf32(capital_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                                           const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int passes);
capital_status_t capital_set_trailing_precision(capital_ctx* ctx, int mode);
capital_status_t capital_tf32_stats(const capital_ctx* ctx, int64_t* launches, double* flops);

/* lapack::engine::_potrf('U') + _trtri('U','N') fused (lapack/interface.hpp:30-58; called back to
 * back at cholinv/policy.h:199-201): A (n x n, upper read) -> R, Rinv upper (lower zeroed). DEVICE pointers. */
capital_status_t capital_lapack_potrf_trtri_f64(capital_ctx* ctx, int64_t n, const double* A, int64_t lda,
                                                double* R, int64_t ldr, double* Rinv, int64_t ldri);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CAPITAL_B200_H */
