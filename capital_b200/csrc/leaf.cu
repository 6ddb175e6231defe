// Leaf of the CholInv recursion: potrf('U') + trtri('U','N') of one nb x nb block (nb <= 64) held in shared
// memory by a single CTA.  Replaces the LAPACKE_dpotrf / LAPACKE_dtrtri pair the reference calls on its
// gathered base-case block (cholinv/policy.h:199-201, lapack/interface.hpp:30-58); unlike the reference the
// pivot sign is checked and reported (CAPITAL_ERR_NOT_SPD) instead of being dropped.
//
// potrf: right-looking with deferred row scaling (one barrier per column): after step k the pivot row keeps
//        a[k,j] = r[k,k] * r[k,j]; trailing update a[i,j] -= a[k,i] a[k,j] / a[k,k]; R = D^{-1/2} a at the end.
// trtri: recursive doubling X12 = -X11 R12 X22 over block sizes 1,2,4,...: log2(nb) levels, all pairs of a
//        level processed concurrently by the whole CTA (two barriers per level).
#include "common.cuh"

namespace {
constexpr int LD = LEAF_MAX + 1;  // padded leading dimension: conflict-free row and column walks

__global__ void __launch_bounds__(256, 1)
    leaf_kernel(int nb, const double* __restrict__ W, long long ldw, double* __restrict__ R, long long ldr, double* __restrict__ Ri,
                long long ldri, double* __restrict__ RiT, long long ldrit, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* a = sm;              // working copy / later the inverse
  double* r = sm + LEAF_MAX * LD;      // R
  double* t = sm + 2 * LEAF_MAX * LD;  // temp products
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int idx = tid; idx < nb * nb; idx += nt) {
    const int i = idx % nb, j = idx / nb;
    a[i + j * LD] = (i <= j) ? W[i + (long long)j * ldw] : 0.0;
  }
  __syncthreads();

  for (int k = 0; k < nb; k++) {
    double d = a[k + k * LD];
    if (!(d > 0.0)) {
      if (tid == 0) atomicCAS(info, 0, k + 1);
      d = 1.0;
    }
    const double inv = 1.0 / d;
    const int m = nb - k - 1;
    for (int idx = tid; idx < m * m; idx += nt) {
      const int ii = idx % m, jj = idx / m;
      if (ii <= jj) {
        const int i = k + 1 + ii, j = k + 1 + jj;
        a[i + j * LD] -= a[k + i * LD] * a[k + j * LD] * inv;
      }
    }
    __syncthreads();
  }
  // R = D^{-1/2} a (upper), zeros below
  for (int idx = tid; idx < nb * nb; idx += nt) {
    const int i = idx % nb, j = idx / nb;
    double v = 0.0;
    if (i <= j) {
      double d = a[i + i * LD];
      if (!(d > 0.0)) d = 1.0;
      v = (i == j) ? sqrt(d) : a[i + j * LD] / sqrt(d);
    }
    r[i + j * LD] = v;
  }
  __syncthreads();
  // inverse: start with the diagonal, zeros elsewhere
  for (int idx = tid; idx < nb * nb; idx += nt) {
    const int i = idx % nb, j = idx / nb;
    a[i + j * LD] = (i == j) ? 1.0 / r[i + i * LD] : 0.0;
  }
  __syncthreads();
  for (int bs = 1; bs < nb; bs <<= 1) {
    const int span = 2 * bs;
    const int npairs = (nb + span - 1) / span;
    // T = X11 * R12  (bs x bs2), X11 upper triangular
    for (int idx = tid; idx < npairs * bs * bs; idx += nt) {
      const int p = idx / (bs * bs), e = idx % (bs * bs);
      const int li = e % bs, lj = e / bs;
      const int o = p * span;
      const int i = o + li, j = o + bs + lj;
      if (i < nb && j < nb) {
        double s = 0.0;
        const int kend = min(o + bs, nb);
        for (int k = i; k < kend; k++) s += a[i + k * LD] * r[k + j * LD];
        t[i + j * LD] = s;
      }
    }
    __syncthreads();
    // X12 = -T * X22, X22 upper triangular
    for (int idx = tid; idx < npairs * bs * bs; idx += nt) {
      const int p = idx / (bs * bs), e = idx % (bs * bs);
      const int li = e % bs, lj = e / bs;
      const int o = p * span;
      const int i = o + li, j = o + bs + lj;
      if (i < nb && j < nb) {
        double s = 0.0;
        for (int k = o + bs; k <= j; k++) s += t[i + k * LD] * a[k + j * LD];
        a[i + j * LD] = -s;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * nb; idx += nt) {
    const int i = idx % nb, j = idx / nb;
    R[i + (long long)j * ldr] = r[i + j * LD];
    Ri[i + (long long)j * ldri] = a[i + j * LD];
  }
  if (RiT != nullptr) {
    for (int idx = tid; idx < nb * nb; idx += nt) {
      const int j = idx % nb, i = idx / nb;  // RiT(j, i) = Ri(i, j)
      RiT[j + (long long)i * ldrit] = a[i + j * LD];
    }
  }
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
