// Bottom of the CholInv recursion: potrf('U') + trtri('U','N') of one diagonal block.  Replaces the LAPACKE_dpotrf /
// LAPACKE_dtrtri pair the reference calls on its gathered base-case block (cholinv/policy.h:199-201,
// lapack/interface.hpp:30-58); unlike the reference the pivot sign is checked and reported (CAPITAL_ERR_NOT_SPD).
//
// Two kernels share one device routine:
//   leaf_kernel      one CTA, nb <= 64, everything in shared memory.
//   basecase_kernel  one thread-block cluster (8 CTAs) for nb = 64 t (t <= 8): blocked right-looking Cholesky with
//                    64-wide panels -- diagonal block by the leaf routine, row panel and trailing update as 64x64x64
//                    DMMA tile products spread over the cluster, hardware cluster barriers between phases -- followed
//                    by the blocked triangular inverse.  One launch replaces ~60 latency-bound launches of the
//                    recursion below 512 (r01a profile: leaves + small GEMMs were 37% of the step).
// The critical path of a leaf is the pivot chain (64 dependent rsqrt + rank-1 updates), so the leaf keeps all
// 256 threads on a fixed 16x16 grid (no index division), scales the pivot row with two warps, and uses one
// rsqrt per pivot instead of a sqrt and a divide.
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace {
constexpr int LD = LEAF_MAX + 1;   // leaf arrays: conflict-free row and column walks
constexpr int TLD = 68;            // DMMA tiles: rows of 64 k-contiguous doubles, padded so that fragment loads
                                   // (row g, k q) of a half-warp touch 16 distinct bank pairs
constexpr int TILE_DOUBLES = 64 * TLD;
constexpr int BC_CLUSTER = 8;

// ---- leaf: factor + invert an nb x nb block held in shared memory ---------------------------------------------
// a : in  upper triangle of the SPD block (destroyed), out R^{-1} (upper, zeros below)
// r : out R (upper, zeros below)
// t : scratch
__device__ void leaf_factor_invert(int nb, double* __restrict__ a, double* __restrict__ r, double* __restrict__ t, int* info,
                                   int pivot_base) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int k = 0; k < nb; k++) {
    if (tid < 64 && tid >= k && tid < nb) {
      double d = a[k + k * LD];
      if (!(d > 0.0)) {
        if (tid == k) atomicCAS(info, 0, pivot_base + k + 1);
        d = 1.0;
      }
      const double rs = rsqrt(d);
      r[k + tid * LD] = (tid == k) ? d * rs : a[k + tid * LD] * rs;
    }
    __syncthreads();
#pragma unroll
    for (int ia = 0; ia < 4; ia++) {
      const int i = ty + 16 * ia;
      if (i > k && i < nb) {
        const double ri = r[k + i * LD];
#pragma unroll
        for (int jb = 0; jb < 4; jb++) {
          const int j = tx + 16 * jb;
          if (j >= i && j < nb) a[i + j * LD] -= ri * r[k + j * LD];
        }
      }
    }
    __syncthreads();
  }
  // zero the strictly lower part of r, start the inverse with the reciprocal diagonal
#pragma unroll
  for (int ia = 0; ia < 4; ia++) {
    const int i = ty + 16 * ia;
#pragma unroll
    for (int jb = 0; jb < 4; jb++) {
      const int j = tx + 16 * jb;
      if (i < nb && j < nb) {
        if (i > j) r[i + j * LD] = 0.0;
        a[i + j * LD] = (i == j) ? 1.0 / r[i + i * LD] : 0.0;
      }
    }
  }
  __syncthreads();
  // recursive doubling: X12 = -X11 R12 X22 for block sizes 1, 2, 4, ...
  int lg = 0;
  for (int bs = 1; bs < nb; bs <<= 1, lg++) {
    const int span = 2 * bs;
    const int npairs = (nb + span - 1) / span;
    const int per = bs * bs, total = npairs * per;
    for (int idx = tid; idx < total; idx += 256) {  // T = X11 R12
      const int p = idx >> (2 * lg), e = idx & (per - 1);
      const int li = e & (bs - 1), lj = e >> lg;
      const int o = p * span, i = o + li, j = o + bs + lj;
      if (j < nb) {
        double s = 0.0;
        const int kend = o + bs;
        for (int k = i; k < kend; k++) s += a[i + k * LD] * r[k + j * LD];
        t[i + j * LD] = s;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < total; idx += 256) {  // X12 = -T X22
      const int p = idx >> (2 * lg), e = idx & (per - 1);
      const int li = e & (bs - 1), lj = e >> lg;
      const int o = p * span, i = o + li, j = o + bs + lj;
      if (j < nb) {
        double s = 0.0;
        for (int k = o + bs; k <= j; k++) s += t[i + k * LD] * a[k + j * LD];
        a[i + j * LD] = -s;
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void leaf_load(int nb, const double* __restrict__ W, long long ldw, double* __restrict__ a) {
  for (int idx = threadIdx.x; idx < nb * nb; idx += 256) {
    const int i = idx % nb, j = idx / nb;
    a[i + j * LD] = (i <= j) ? __ldcg(W + i + (long long)j * ldw) : 0.0;
  }
}
__device__ __forceinline__ void leaf_store(int nb, const double* __restrict__ a, const double* __restrict__ r, double* __restrict__ R,
                                           long long ldr, double* __restrict__ Ri, long long ldri, double* __restrict__ RiT,
                                           long long ldrit) {
  for (int idx = threadIdx.x; idx < nb * nb; idx += 256) {
    const int i = idx % nb, j = idx / nb;
    R[i + (long long)j * ldr] = r[i + j * LD];
    Ri[i + (long long)j * ldri] = a[i + j * LD];
  }
  if (RiT != nullptr) {
    for (int idx = threadIdx.x; idx < nb * nb; idx += 256) {
      const int j = idx % nb, i = idx / nb;  // RiT(j, i) = Ri(i, j)
      RiT[j + (long long)i * ldrit] = a[i + j * LD];
    }
  }
}

__global__ void __launch_bounds__(256, 1)
    leaf_kernel(int nb, const double* __restrict__ W, long long ldw, double* __restrict__ R, long long ldr, double* __restrict__ Ri,
                long long ldri, double* __restrict__ RiT, long long ldrit, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* a = sm;
  double* r = sm + LEAF_MAX * LD;
  double* t = sm + 2 * LEAF_MAX * LD;
  leaf_load(nb, W, ldw, a);
  __syncthreads();
  leaf_factor_invert(nb, a, r, t, info, 0);
  leaf_store(nb, a, r, R, ldr, Ri, ldri, RiT, ldrit);
}

// ---- 64x64x64 DMMA tile products out of shared memory -------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// copy a 64 (k) x 64 (cols) global block (k contiguous) into a padded tile: dst[col * TLD + k]
__device__ __forceinline__ void tile_load(double* __restrict__ dst, const double* __restrict__ src, long long ld) {
  const int k = threadIdx.x & 63, c0 = threadIdx.x >> 6;
#pragma unroll
  for (int r = 0; r < 16; r++) {
    const int c = c0 + 4 * r;
    dst[c * TLD + k] = __ldcg(src + k + (long long)c * ld);
  }
}
// acc += As^T Bs for the 64x64 tile; warp w owns rows (w&1)*32.., cols (w>>1)*16..; As/Bs: [row][k] padded tiles.
// kmax: contraction length actually needed (multiple of 4, <= 64)
__device__ __forceinline__ void tile_mma(double (&acc)[4][2][2], const double* __restrict__ As, const double* __restrict__ Bs, int kmax) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const double* ap = As + ((w & 1) * 32 + g) * TLD + q;
  const double* bp = Bs + ((w >> 1) * 16 + g) * TLD + q;
#pragma unroll 4
  for (int k0 = 0; k0 < kmax; k0 += 4) {
    double af[4], bf[2];
#pragma unroll
    for (int i = 0; i < 4; i++) af[i] = ap[i * 8 * TLD + k0];
#pragma unroll
    for (int j = 0; j < 2; j++) bf[j] = bp[j * 8 * TLD + k0];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}
__device__ __forceinline__ void acc_zero(double (&acc)[4][2][2]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
}
// element (row, col) of the tile owned by this lane for fragment (i, j), value e
#define TILE_ROW(i) ((w & 1) * 32 + (i) * 8 + g)
#define TILE_COL(j, e) ((w >> 1) * 16 + (j) * 8 + 2 * q + (e))

// ---- cluster base case ----------------------------------------------------------------------------------------
// W (nb x nb, upper read, destroyed) -> R, Ri, RiT blocks (full nb x nb blocks written: zeros in the other triangle).
__global__ void __cluster_dims__(BC_CLUSTER, 1, 1) __launch_bounds__(256, 1)
    basecase_kernel(int nb, double* __restrict__ W, long long ldw, double* __restrict__ R, long long ldr, double* __restrict__ Ri,
                    long long ldri, double* __restrict__ RiT, long long ldrit, int* __restrict__ info, long long* __restrict__ dbg) {
  extern __shared__ double sm[];
  double* sA = sm;                     // tile / leaf array a
  double* sB = sm + TILE_DOUBLES;      // tile / leaf array r
  double* sT = sm + 2 * TILE_DOUBLES;  // tile / leaf scratch
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int T = nb >> 6;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  double acc[4][2][2];
  int dbgi = 0;
#define DBG_STAMP() do { if (dbg && rank == 0 && threadIdx.x == 0) dbg[dbgi++] = clock64(); } while (0)
  DBG_STAMP();

  // ---------------- Cholesky: right-looking over 64-wide block columns ----------------
  for (int jb = 0; jb < T; jb++) {
    const long long o = (long long)jb * 64;
    if (rank == (jb % BC_CLUSTER)) {  // diagonal block
      leaf_load(64, W + o + o * ldw, ldw, sA);
      __syncthreads();
      leaf_factor_invert(64, sA, sB, sT, info, jb * 64);
      leaf_store(64, sA, sB, R + o + o * ldr, ldr, Ri + o + o * ldri, ldri, RiT + o + o * ldrit, ldrit);
    }
    if (jb == 0) DBG_STAMP();
    __threadfence();
    cluster.sync();
    if (jb == 0) DBG_STAMP();
    // row panel: R(jb, j) = Rinv_jj^T W(jb, j), j > jb
    int work = 0;
    for (int j = jb + 1; j < T; j++, work++) {
      if (work % BC_CLUSTER != rank) continue;
      __syncthreads();
      tile_load(sA, Ri + o + o * ldri, ldri);                         // A[k][i] = Rinv_jj(k, i)
      tile_load(sB, W + o + (long long)j * 64 * ldw, ldw);            // B[k][c] = W(jb rows, j cols)
      __syncthreads();
      acc_zero(acc);
      tile_mma(acc, sA, sB, 64);
      double* out = R + o + (long long)j * 64 * ldr;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++)
#pragma unroll
          for (int e = 0; e < 2; e++) out[TILE_ROW(i) + (long long)TILE_COL(jj, e) * ldr] = acc[i][jj][e];
    }
    if (jb == 0) DBG_STAMP();
    __threadfence();
    cluster.sync();
    if (jb == 0) DBG_STAMP();
    // trailing update: W(i, j) -= R(jb, i)^T R(jb, j), jb < i <= j
    work = 0;
    for (int j = jb + 1; j < T; j++)
      for (int i = jb + 1; i <= j; i++, work++) {
        if (work % BC_CLUSTER != rank) continue;
        __syncthreads();
        tile_load(sA, R + o + (long long)i * 64 * ldr, ldr);
        tile_load(sB, R + o + (long long)j * 64 * ldr, ldr);
        __syncthreads();
        acc_zero(acc);
        tile_mma(acc, sA, sB, 64);
        double* out = W + (long long)i * 64 + (long long)j * 64 * ldw;
#pragma unroll
        for (int ii = 0; ii < 4; ii++)
#pragma unroll
          for (int jj = 0; jj < 2; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              double* p = out + TILE_ROW(ii) + (long long)TILE_COL(jj, e) * ldw;
              *p = __ldcg(p) - acc[ii][jj][e];
            }
      }
    if (jb == 0) DBG_STAMP();
    __threadfence();
    cluster.sync();
    if (jb == 0) DBG_STAMP();
  }
  DBG_STAMP();
  // zero the strictly-lower blocks of R (the leaf wrote the diagonal blocks completely)
  {
    int work = 0;
    for (int j = 0; j < T; j++)
      for (int i = j + 1; i < T; i++, work++) {
        if (work % BC_CLUSTER != rank) continue;
        for (int idx = threadIdx.x; idx < 4096; idx += 256)
          R[(long long)i * 64 + (idx & 63) + ((long long)j * 64 + (idx >> 6)) * ldr] = 0.0;
      }
  }
  // ---------------- inverse: block column j from the columns before it ----------------
  //   Rinv(i, j) = -[ sum_{k=i}^{j-1} Rinv(i, k) R(k, j) ] Rinv(j, j),  i < j
  for (int j = 1; j < T; j++) {
    for (int i = rank; i < j; i += BC_CLUSTER) {
      acc_zero(acc);
      for (int k = i; k < j; k++) {
        __syncthreads();
        tile_load(sA, RiT + (long long)k * 64 + (long long)i * 64 * ldrit, ldrit);  // A[kk][ii] = RiT(k blk, i blk) = Rinv(i, k)^T
        tile_load(sB, R + (long long)k * 64 + (long long)j * 64 * ldr, ldr);
        __syncthreads();
        tile_mma(acc, sA, sB, 64);
      }
      __syncthreads();
      // S (64 x 64, rows i, cols t) -> shared as the next A operand: A[row i][k = t]
#pragma unroll
      for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++)
#pragma unroll
          for (int e = 0; e < 2; e++) sT[TILE_ROW(ii) * TLD + TILE_COL(jj, e)] = acc[ii][jj][e];
      tile_load(sB, Ri + (long long)j * 64 + (long long)j * 64 * ldri, ldri);  // B[t][c] = Rinv_jj(t, c)
      __syncthreads();
      acc_zero(acc);
      tile_mma(acc, sT, sB, 64);
      double* o1 = Ri + (long long)i * 64 + (long long)j * 64 * ldri;
      double* o2 = RiT + (long long)j * 64 + (long long)i * 64 * ldrit;
#pragma unroll
      for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const double v = -acc[ii][jj][e];
            o1[TILE_ROW(ii) + (long long)TILE_COL(jj, e) * ldri] = v;
            o2[TILE_COL(jj, e) + (long long)TILE_ROW(ii) * ldrit] = v;
          }
    }
    __threadfence();
    cluster.sync();
  }
  DBG_STAMP();
}
}  // namespace

capital_status_t leaf_cholinv(capital_ctx* ctx, cudaStream_t st, int nb, const double* W, int64_t ldw, double* R, int64_t ldr, double* Ri,
                              int64_t ldri, double* RiT, int64_t ldrit) {
  if (nb <= 0) return CAPITAL_OK;
  if (nb > LEAF_MAX) return CAPITAL_ERR_INVALID;
  constexpr int smem = 3 * LEAF_MAX * LD * (int)sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    CAP_CUDA(cudaFuncSetAttribute(leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  leaf_kernel<<<1, 256, smem, st>>>(nb, W, ldw, R, ldr, Ri, ldri, RiT, ldrit, ctx->d_info);
  ctx->counters.kernel_launches++;
  ctx->counters.leaf_launches++;
  CAP_CUDA(cudaGetLastError());
  return CAPITAL_OK;
}

// nb must be a multiple of 64, 128 <= nb <= BASECASE_MAX, and RiT non-null.
capital_status_t basecase_cholinv(capital_ctx* ctx, cudaStream_t st, int nb, double* W, int64_t ldw, double* R, int64_t ldr, double* Ri,
                                  int64_t ldri, double* RiT, int64_t ldrit) {
  if (nb % 64 != 0 || nb < 64 || nb > BASECASE_MAX || RiT == nullptr) return CAPITAL_ERR_INVALID;
  constexpr int smem = 3 * TILE_DOUBLES * (int)sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    CAP_CUDA(cudaFuncSetAttribute(basecase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  long long* dbg = nullptr;
  if (getenv("CAPITAL_BC_DEBUG")) {
    CAP_TRY(ctx->workspace("bc_dbg", 64 * sizeof(long long), (void**)&dbg));
    CAP_CUDA(cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), st));
  }
  basecase_kernel<<<BC_CLUSTER, 256, smem, st>>>(nb, W, ldw, R, ldr, Ri, ldri, RiT, ldrit, ctx->d_info, dbg);
  if (dbg) {
    long long h[16];
    CAP_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    CAP_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[bc nb=%d] leaf0=%lld bar=%lld panel0=%lld bar=%lld trail0=%lld bar=%lld | chol_total=%lld inverse=%lld total=%lld cycles\n", nb,
            h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[0], h[8] - h[7], h[8] - h[0]);
  }
  ctx->counters.kernel_launches++;
  ctx->counters.leaf_launches++;
  CAP_CUDA(cudaGetLastError());
  return CAPITAL_OK;
}
