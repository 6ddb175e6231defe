// Single-device recursive CholInv on a dense block (the schedule of cholinv::invoke, cholinv.hpp:87-165, with
// every product expressed as C = alpha A^T B + beta C so that one DMMA kernel serves all four of them).
//
// Buffers (all column-major, same leading dimension allowed):
//   W   : in  -- SPD block, upper triangle read; destroyed (Schur complements; the dead lower-left blocks are
//               reused as the scratch for T^T, which replaces the reference's rect_table temporaries, cholinv.h:37-38)
//   R   : out -- upper factor, A = R^T R
//   Ri  : out -- R^{-1} (upper);   RiT : out -- (R^{-1})^T (lower), kept so that "Left/Upper/NoTrans" and
//               "Right/Upper/NoTrans" trmm (cholinv.hpp:150-154) are both A^T B products with K-contiguous operands.
// Ri and RiT must be zero on entry (the triangular products read whole diagonal tiles).
#include "common.cuh"

namespace {

int64_t split_point(int64_t n) {
  // halves (split = 1), rounded so that leaves stay LEAF_MAX-aligned
  int64_t s1 = n >> 1;
  if (n > 2 * LEAF_MAX) s1 = round_up(s1, LEAF_MAX);
  else s1 = round_up(s1, 2);  // even split points keep every window 16-byte aligned for TMA
  if (s1 >= n) s1 = n >> 1;
  return s1;
}

// Levels above the base case (n > bc) split by the reference's rule s1 = n >> split (cholinv.hpp:92,107); it fixes
// which Rinv block stays zero when complete_inv == 0.  Below it -- the reference's potrf/trtri base case
// (cholinv.hpp:93-104) -- the recursion continues with 64-aligned halves down to the shared-memory leaf.
capital_status_t rec(capital_ctx* ctx, cudaStream_t st, int64_t n, double* W, int64_t ldw, double* R, int64_t ldr, double* Ri,
                     int64_t ldri, double* RiT, int64_t ldrit, bool complete, int64_t bc, int split) {
  int64_t s1;
  if (n > bc && (n >> split) >= split && (n >> split) > 0 && (n > LEAF_MAX || !complete)) s1 = n >> split;
  else if (n <= LEAF_MAX) return leaf_cholinv(ctx, st, (int)n, W, ldw, R, ldr, Ri, ldri, RiT, ldrit);
  else if (n <= BASECASE_MAX && n % 64 == 0 && complete) return basecase_cholinv(ctx, st, (int)n, W, ldw, R, ldr, Ri, ldri, RiT, ldrit);
  else s1 = split_point(n);
  const int64_t s2 = n - s1;
  double* W12 = W + s1 * ldw;
  double* W21 = W + s1;  // dead lower-left block: scratch for T^T (s2 x s1)
  double* W22 = W + s1 * ldw + s1;
  double* R12 = R + s1 * ldr;
  double* R22 = R + s1 * ldr + s1;
  double* Ri12 = Ri + s1 * ldri;
  double* Ri22 = Ri + s1 * ldri + s1;
  double* RiT21 = RiT + s1;
  double* RiT22 = RiT + s1 * ldrit + s1;

  CAP_TRY(rec(ctx, st, s1, W, ldw, R, ldr, Ri, ldri, RiT, ldrit, true, bc, split));
  // "trsm" via the inverse (cholinv.hpp:116-122): R12 = Rinv11^T A12
  CAP_TRY(gemm_tn(ctx, st, s1, s2, s1, 1.0, Ri, ldri, W12, ldw, 0.0, R12, ldr, CAPITAL_GEMM_A_UPPER));
  // trailing update (cholinv.hpp:131-134): A22 -= R12^T R12, upper tiles only
  CAP_TRY(gemm_tn(ctx, st, s2, s2, s1, -1.0, R12, ldr, R12, ldr, 1.0, W22, ldw, CAPITAL_GEMM_C_UPPER));
  CAP_TRY(rec(ctx, st, s2, W22, ldw, R22, ldr, Ri22, ldri, RiT22, ldrit, true, bc, split));
  if (complete) {
    // inverse combine (cholinv.hpp:147-155): Rinv12 = -Rinv11 R12 Rinv22
    //   T^T = R12^T Rinv11^T      (B = RiT11, lower triangular)
    CAP_TRY(gemm_tn(ctx, st, s2, s1, s1, 1.0, R12, ldr, RiT, ldrit, 0.0, W21, ldw, CAPITAL_GEMM_B_LOWER));
    //   Rinv12 = -(T^T)^T Rinv22  (B = Ri22, upper triangular)
    CAP_TRY(gemm_tn(ctx, st, s1, s2, s2, -1.0, W21, ldw, Ri22, ldri, 0.0, Ri12, ldri, CAPITAL_GEMM_B_UPPER));
    CAP_TRY(transpose_block(ctx, st, s1, s2, Ri12, ldri, RiT21, ldrit, 1.0));
  }
  return CAPITAL_OK;
}

}  // namespace

capital_status_t cholinv_local(capital_ctx* ctx, cudaStream_t st, int64_t n, double* W, int64_t ldw, double* R, int64_t ldr, double* Ri,
                               int64_t ldri, double* RiT, int64_t ldrit, bool complete_top, int64_t bc, int split) {
  return rec(ctx, st, n, W, ldw, R, ldr, Ri, ldri, RiT, ldrit, complete_top, bc, split);
}
