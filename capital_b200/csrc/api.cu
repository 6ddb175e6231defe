// C ABI (include/capital_b200.h): context, grid helpers, generators, validators and the factor entry points.
// Host orchestration only -- every flop and every byte of layout work runs in the kernels of gemm_tn.cu,
// leaf.cu and layout.cu.  There is no CPU fallback: without a usable sm_100 device capital_create fails.
#include "common.cuh"
#include "dist.cuh"
#include "peer.cuh"
#include <math.h>
#include <stdlib.h>

capital_status_t capital_ctx::workspace(const std::string& name, size_t bytes, void** out) {
  capital_ctx* ctx = this;
  DeviceBuf& b = pool[name];
  if (b.bytes < bytes) {
    if (b.p) CAP_CUDA(cudaFree(b.p));
    b.p = nullptr; b.bytes = 0;
    CAP_CUDA(cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
  }
  *out = b.p;
  return CAPITAL_OK;
}

int capital_ctx::stream_id(cudaStream_t st) const {
  if (st == hi) return 1;
  if (st == side) return 2;
  if (st == side_deep[0]) return 3;
  if (st == side_deep[1]) return 4;
  if (peer) {
    const Peer* P = (const Peer*)peer;
    for (int q = 0; q < PEER_Q; q++) if (st == P->push[q]) return 5 + q;
  }
  if (st == copy_in) return 10;
  if (st == copy_out) return 11;
  return 0;
}
int capital_ctx::tl_begin(cudaStream_t st, int kind, double a, double b, double c) {
  if (!timeline) return -1;
  while (tl_pool.size() < tl_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    tl_pool.push_back(e);
  }
  TlRec r{tl_pool[tl_used], tl_pool[tl_used + 1], stream_id(st), kind, a, b, c};
  tl_used += 2;
  cudaEventRecord(r.e0, st);
  tl.push_back(r);
  return (int)tl.size() - 1;
}
void capital_ctx::tl_end(cudaStream_t st, int idx) {
  if (idx >= 0) cudaEventRecord(tl[idx].e1, st);
}

bool cap_is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Bring `count` doubles to the device (no-op for device pointers).
capital_status_t cap_stage_in(capital_ctx* ctx, const double* src, size_t count, const char* name, const double** out) {
  if (cap_is_device_ptr(src)) { *out = src; return CAPITAL_OK; }
  void* d;
  CAP_TRY(ctx->workspace(name, count * 8, &d));
  CAP_CUDA(cudaMemcpyAsync(d, src, count * 8, cudaMemcpyHostToDevice, ctx->stream));
  ctx->counters.h2d_bytes += (int64_t)count * 8;
  *out = (const double*)d;
  return CAPITAL_OK;
}
// Device output target for a caller pointer: the pointer itself if on device, else a workspace to be copied back.
capital_status_t cap_stage_out_begin(capital_ctx* ctx, double* dst, size_t count, const char* name, double** dev) {
  if (cap_is_device_ptr(dst)) { *dev = dst; return CAPITAL_OK; }
  void* d;
  CAP_TRY(ctx->workspace(name, count * 8, &d));
  *dev = (double*)d;
  return CAPITAL_OK;
}
capital_status_t cap_stage_out_end(capital_ctx* ctx, double* dst, size_t count, const double* dev) {
  if (dev == dst) return CAPITAL_OK;
  CAP_CUDA(cudaMemcpyAsync(dst, dev, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->counters.d2h_bytes += (int64_t)count * 8;
  return CAPITAL_OK;
}

capital_status_t cap_check_info(capital_ctx* ctx) {
  int info = 0;
  CAP_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CAP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info == -3) {
    ctx->set_error("the experimental TF32 trailing-update kernel gave up on an mbarrier wait (gemm_tf32.cu watchdog): results are invalid");
    return CAPITAL_ERR_CUDA;
  }
  if (info < 0) {
    ctx->set_error("a wait on a peer GPU timed out (a rank of the grid died, or the ranks did not make the same sequence of calls)");
    return CAPITAL_ERR_COMM;
  }
  if (info != 0) {
    ctx->set_error("matrix is not positive definite: non-positive pivot " + std::to_string(info) + " in a base-case block");
    return CAPITAL_ERR_NOT_SPD;
  }
  return CAPITAL_OK;
}

// ---- host-pointer streaming of cholinv::factor (single GPU) ----------------------------------------------------
namespace {
struct HostIO {
  capital_ctx* ctx = nullptr;
  int64_t L = 0, ld = 0;
  double *Rm = nullptr, *Ri = nullptr, *dR = nullptr, *dRinv = nullptr, *hR = nullptr, *hRinv = nullptr;
  bool packed = true;
  std::vector<std::pair<int64_t, cudaEvent_t>> chunks;  // (col_end, arrived)
  int64_t waited = 0;
  int64_t cols_out = 0, rinv_cols_out = 0;
  bool rinv_streams = false;  // Rinv columns right of the top split are final as soon as R's are (complete_inv == 0)
  cudaEvent_t e_out = nullptr;
  double *zR = nullptr, *zRinv = nullptr;  // experimental block-wise output: device aliases of the pinned host arrays
};
capital_status_t io_event(capital_ctx* ctx, cudaEvent_t* e) {
  if (ctx->io_used == ctx->io_pool.size()) {
    cudaEvent_t ev;
    CAP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->io_pool.push_back(ev);
  }
  *e = ctx->io_pool[ctx->io_used++];
  return CAPITAL_OK;
}
capital_status_t hostio_need_cols(void* user, cudaStream_t st, int64_t col_end) {
  HostIO* io = (HostIO*)user;
  capital_ctx* ctx = io->ctx;
  if (col_end <= io->waited) return CAPITAL_OK;
  for (auto& ch : io->chunks)
    if (ch.first >= col_end) {
      CAP_CUDA(cudaStreamWaitEvent(st, ch.second, 0));
      io->waited = ch.first;
      return CAPITAL_OK;
    }
  return CAPITAL_OK;
}
// columns [cols_out, col_end) of R are final (and of Rinv too when the top-level inverse block is skipped, complete_inv == 0:
// then Rinv's columns right of the top split only hold the right child's own inverse): pack them -- a contiguous range of
// the packed triangle -- and start their D2H while the rest of the factorization runs.
capital_status_t hostio_left_done(void* user, cudaStream_t st, int64_t col_end, int depth) {
  HostIO* io = (HostIO*)user;
  capital_ctx* ctx = io->ctx;
  const int64_t c0 = io->cols_out;
  if (col_end <= c0) return CAPITAL_OK;
  // Rinv: left of the top split always; the next range only when the top-level inverse block is skipped (deeper ranges still
  // miss the off-diagonal inverse blocks of the right-spine ancestors, computed after their right children)
  const bool rinv_too = depth == 0 || (depth == 1 && io->rinv_streams && io->rinv_cols_out == c0);
  const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = (size_t)col_end * (col_end + 1) / 2 - off;
  // packing is HBM-bound filler work: it goes to the low-priority stream (joined before cholinv_local returns), not to the chain
  // (host callers keep it on the chain: the D2H of the range should start right away, not behind queued deferred work)
  cudaStream_t ps = (ctx->side && !ctx->no_overlap && !io->hR && !io->hRinv) ? ctx->side : st;
  cudaEvent_t e;
  if (ps != st) {
    CAP_TRY(io_event(ctx, &e));
    CAP_CUDA(cudaEventRecord(e, st));
    CAP_CUDA(cudaStreamWaitEvent(ps, e, 0));
  }
  CAP_TRY(pack_upper(ctx, ps, io->L, io->Rm, io->ld, io->dR, 0, c0, col_end));
  if (rinv_too) CAP_TRY(pack_upper(ctx, ps, io->L, io->Ri, io->ld, io->dRinv, 0, c0, col_end));
  io->cols_out = col_end;
  if (rinv_too) io->rinv_cols_out = col_end;
  if (!io->hR && !io->hRinv) return CAPITAL_OK;  // device outputs: nothing to copy out
  CAP_TRY(io_event(ctx, &e));
  CAP_CUDA(cudaEventRecord(e, ps));
  CAP_CUDA(cudaStreamWaitEvent(ctx->copy_out, e, 0));
  if (io->hR) { CAP_CUDA(cudaMemcpyAsync(io->hR + off, io->dR + off, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_out)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
  if (io->hRinv && rinv_too) { CAP_CUDA(cudaMemcpyAsync(io->hRinv + off, io->dRinv + off, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_out)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
  CAP_TRY(io_event(ctx, &io->e_out));
  CAP_CUDA(cudaEventRecord(io->e_out, ctx->copy_out));
  return CAPITAL_OK;
}
int64_t hostio_cols_waited(void* user) { return ((HostIO*)user)->waited; }
// pack columns [done, col_end) of R (or Rinv) on the chain and queue their D2H on the copy-out stream (host outputs only)
capital_status_t hostio_emit(HostIO* io, cudaStream_t st, bool r_part, int64_t col_end) {
  capital_ctx* ctx = io->ctx;
  int64_t& done = r_part ? io->cols_out : io->rinv_cols_out;
  const int64_t c0 = done;
  if (col_end <= c0) return CAPITAL_OK;
  const double* src = r_part ? io->Rm : io->Ri;
  double* dev = r_part ? io->dR : io->dRinv;
  double* host = r_part ? io->hR : io->hRinv;
  const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = (size_t)col_end * (col_end + 1) / 2 - off;
  CAP_TRY(pack_upper(ctx, st, io->L, src, io->ld, dev, 0, c0, col_end));
  done = col_end;
  if (!host) return CAPITAL_OK;
  cudaEvent_t e;
  CAP_TRY(io_event(ctx, &e));
  CAP_CUDA(cudaEventRecord(e, st));
  CAP_CUDA(cudaStreamWaitEvent(ctx->copy_out, e, 0));
  CAP_CUDA(cudaMemcpyAsync(host + off, dev + off, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_out));
  ctx->counters.d2h_bytes += (int64_t)cnt * 8;
  CAP_TRY(io_event(ctx, &io->e_out));
  CAP_CUDA(cudaEventRecord(io->e_out, ctx->copy_out));
  return CAPITAL_OK;
}
// the top-level right child is done: R is final everywhere (and so is Rinv when the top-level inverse block is skipped); its last
// columns leave while the top-level inverse block is still being computed
capital_status_t hostio_right_done(void* user, cudaStream_t st) {
  HostIO* io = (HostIO*)user;
  CAP_TRY(hostio_emit(io, st, true, io->L));
  if (io->rinv_streams) CAP_TRY(hostio_emit(io, st, false, io->L));
  return CAPITAL_OK;
}
capital_status_t hostio_inv_cols(void* user, cudaStream_t st, int64_t col_end) { return hostio_emit((HostIO*)user, st, false, col_end); }
// experimental block-wise output: the block is stored straight from the rect work buffer into its slots of the pinned packed array
capital_status_t hostio_block_done(void* user, cudaStream_t st, int which, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
  HostIO* io = (HostIO*)user;
  capital_ctx* ctx = io->ctx;
  cudaEvent_t e;
  CAP_TRY(io_event(ctx, &e));
  CAP_CUDA(cudaEventRecord(e, st));
  CAP_CUDA(cudaStreamWaitEvent(ctx->zc_out, e, 0));
  CAP_TRY(emit_block_packed(ctx, ctx->zc_out, which ? io->Ri : io->Rm, io->ld, which ? io->zRinv : io->zR, r0, r1, c0, c1, ctx->zc_ctas));
  for (int64_t c = c0; c < c1; c++) {  // bytes that cross the link: the block clipped to the upper triangle
    const int64_t re = r1 < c + 1 ? r1 : c + 1;
    if (re > r0) ctx->counters.d2h_bytes += (re - r0) * 8;
  }
  return CAPITAL_OK;
}
// device-visible alias of a pinned (cudaHostAlloc / cudaHostRegister) host pointer, or nullptr
double* host_alias(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? (double*)at.devicePointer : nullptr;
}
}  // namespace

extern "C" {

capital_status_t capital_grid_square(int size, int rank, int c, int layout, int num_chunks, capital_grid_t* out) {
  if (!out || size <= 0 || rank < 0 || rank >= size || c <= 0 || layout != 0) return CAPITAL_ERR_INVALID;
  capital_grid_t g{};
  g.size = size; g.rank = rank; g.c = c; g.layout = layout; g.num_chunks = num_chunks;
  g.d = (int)nearbyint(ceil(sqrt((double)(size / c))));  // topology.h:77
  g.z = rank % c;
  g.y = rank / (g.d * c);
  g.x = (rank % (g.d * c)) / c;
  if ((int64_t)g.c * g.d * g.d != size) return CAPITAL_ERR_INVALID;
  *out = g;
  return CAPITAL_OK;
}
capital_status_t capital_grid_rect(int size, int rank, int c, int layout, int num_chunks, capital_grid_t* out) {
  if (!out || size <= 0 || rank < 0 || rank >= size || c <= 0 || layout != 0) return CAPITAL_ERR_INVALID;
  if (size % (c * c)) return CAPITAL_ERR_INVALID;
  capital_grid_t g{};
  g.size = size; g.rank = rank; g.c = c; g.layout = layout; g.num_chunks = num_chunks;
  g.d = size / (c * c);  // topology.h:46
  g.z = rank % c;
  g.y = rank / (c * c);
  g.x = (rank % (c * c)) / c;
  *out = g;
  return CAPITAL_OK;
}
int64_t capital_cholinv_bc_dimension(int64_t local_dim, int c, int d, int64_t bc_mult_dim) {
  int64_t bc = (int64_t)c * d;  // cholinv.hpp:15-18
  if (bc_mult_dim < 0) { for (int64_t i = 0; i < -bc_mult_dim; i++) bc *= 2; }
  else { for (int64_t i = 0; i < bc_mult_dim; i++) bc /= 2; }
  if (bc < 1) bc = 1;
  if (bc > local_dim) bc = local_dim;
  return (int64_t)d * (local_dim / bc);
}

// The deferred stream gets its own SM partition (a CUDA green context): all SMs but `reserve` of them.  Deferred GEMM tiles hold an
// SM for up to ~1 ms and a running CTA cannot be preempted, so without a partition the latency-critical kernels of the chain
// (8-CTA cluster base case, small products) wait that long for SMs although their stream has the higher priority: measured
// 8034 us vs 144 us for ten 10-us cluster kernels behind a saturating low-priority kernel (profiles/r02a_probe_greenctx.log).
// The chain's streams stay in the primary context and may use every SM.
static bool make_green_side_stream(capital_ctx* ctx, int reserve, int prio) {
  typedef CUresult (*fn_devget)(CUdevice*, int);
  typedef CUresult (*fn_getres)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*fn_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*fn_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*fn_gcreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*fn_gstream)(CUstream*, CUgreenCtx, unsigned int, int);
  void *p1 = nullptr, *p2 = nullptr, *p3 = nullptr, *p4 = nullptr, *p5 = nullptr, *p6 = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuDeviceGet", &p1, cudaEnableDefault, &q) != cudaSuccess || !p1) return false;
  if (cudaGetDriverEntryPoint("cuDeviceGetDevResource", &p2, cudaEnableDefault, &q) != cudaSuccess || !p2) return false;
  if (cudaGetDriverEntryPoint("cuDevSmResourceSplitByCount", &p3, cudaEnableDefault, &q) != cudaSuccess || !p3) return false;
  if (cudaGetDriverEntryPoint("cuDevResourceGenerateDesc", &p4, cudaEnableDefault, &q) != cudaSuccess || !p4) return false;
  if (cudaGetDriverEntryPoint("cuGreenCtxCreate", &p5, cudaEnableDefault, &q) != cudaSuccess || !p5) return false;
  if (cudaGetDriverEntryPoint("cuGreenCtxStreamCreate", &p6, cudaEnableDefault, &q) != cudaSuccess || !p6) return false;
  CUdevice dev;
  if (((fn_devget)p1)(&dev, ctx->device) != CUDA_SUCCESS) return false;
  CUdevResource all, rem;
  if (((fn_getres)p2)(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  unsigned nb = 0;
  if (((fn_split)p3)(nullptr, &nb, &all, nullptr, 0, 8) != CUDA_SUCCESS || nb < 2) return false;
  std::vector<CUdevResource> groups(nb);
  if (((fn_split)p3)(groups.data(), &nb, &all, &rem, 0, 8) != CUDA_SUCCESS) return false;
  unsigned skip = 0, got = 0;
  while (skip < nb - 1 && (int)got < reserve) got += groups[skip++].sm.smCount;
  std::vector<CUdevResource> far(groups.begin() + skip, groups.begin() + nb);
  if (rem.sm.smCount) far.push_back(rem);
  CUdevResourceDesc desc;
  if (((fn_desc)p4)(&desc, far.data(), (unsigned)far.size()) != CUDA_SUCCESS) return false;
  CUgreenCtx g;
  if (((fn_gcreate)p5)(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  CUstream gs, gd[2];
  if (((fn_gstream)p6)(&gs, g, CU_STREAM_NON_BLOCKING, prio) != CUDA_SUCCESS) return false;
  for (int i = 0; i < 2; i++)
    if (((fn_gstream)p6)(&gd[i], g, CU_STREAM_NON_BLOCKING, prio - 1 - i) != CUDA_SUCCESS) return false;
  ctx->side = (cudaStream_t)gs;
  ctx->side_deep[0] = (cudaStream_t)gd[0];
  ctx->side_deep[1] = (cudaStream_t)gd[1];
  ctx->green = (void*)g;
  ctx->side_sms = 0;
  for (auto& r : far) ctx->side_sms += (int)r.sm.smCount;
  return true;
}

capital_status_t capital_create(capital_ctx** out, const capital_grid_t* grid, int device, void* stream) {
  if (!out || !grid) return CAPITAL_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return CAPITAL_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CAPITAL_ERR_CUDA;
  if (prop.major != 10) return CAPITAL_ERR_CUDA;  // sm_100a binary only: no fallback path exists
  if (cudaSetDevice(device) != cudaSuccess) return CAPITAL_ERR_CUDA;
  capital_ctx* ctx = new capital_ctx();
  ctx->grid = *grid;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CAPITAL_ERR_CUDA; }
    ctx->own_stream = true;
  }
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // lo = numerically greatest = lowest priority
  int reserve = 8;  // SMs kept free of deferred work [env CAPITAL_GREEN_SMS; 0 = plain low-priority stream]
  if (const char* e = getenv("CAPITAL_GREEN_SMS")) reserve = atoi(e);
  bool ok = true;
  if (reserve <= 0 || !make_green_side_stream(ctx, reserve, prio_lo)) {
    ok = cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo) == cudaSuccess;
    for (int i = 0; i < 2; i++) ok = ok && cudaStreamCreateWithPriority(&ctx->side_deep[i], cudaStreamNonBlocking, prio_lo - 1 - i) == cudaSuccess;
  }
  ok = ok && cudaStreamCreateWithPriority(&ctx->hi, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreate(&ctx->ev_start) == cudaSuccess && cudaEventCreate(&ctx->ev_stop) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_info, sizeof(int)) == cudaSuccess && cudaMalloc(&ctx->d_scalars, 16 * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMemset(ctx->d_info, 0, sizeof(int)) == cudaSuccess;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  ok = ok && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn != nullptr;
  if (!ok) { capital_destroy(ctx); return CAPITAL_ERR_CUDA; }
  ctx->encode = (cuTensorMapEncodeTiled_fn)fn;
  if (gemm_tn_init(ctx) != CAPITAL_OK || leaf_init(ctx) != CAPITAL_OK) { capital_destroy(ctx); return CAPITAL_ERR_CUDA; }
  if (const char* e = getenv("CAPITAL_TF32_MIN_K")) ctx->tf32_min_k = atoll(e);
  if (const char* e = getenv("CAPITAL_KCHUNK")) ctx->kchunk = atoll(e);
  if (const char* e = getenv("CAPITAL_FAR_MIN")) ctx->far_min = atoll(e);
  if (const char* e = getenv("CAPITAL_SIDE_MIN")) ctx->side_min = atoll(e);
  if (const char* e = getenv("CAPITAL_ZC_OUT")) ctx->zc_mode = atoi(e);
  if (const char* e = getenv("CAPITAL_ZC_CTAS")) ctx->zc_ctas = atoi(e);
  if (const char* e = getenv("CAPITAL_ZC_DEPTH")) ctx->zc_depth = atoi(e);
  if (ctx->zc_mode && cudaStreamCreateWithPriority(&ctx->zc_out, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { capital_destroy(ctx); return CAPITAL_ERR_CUDA; }
  *out = ctx;
  return CAPITAL_OK;
}

void capital_destroy(capital_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  dist_destroy(ctx);
  for (auto& kv : ctx->pool) if (kv.second.p) cudaFree(kv.second.p);
  if (ctx->d_info) cudaFree(ctx->d_info);
  if (ctx->d_scalars) cudaFree(ctx->d_scalars);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->tl_pool) cudaEventDestroy(e);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  for (int i = 0; i < 2; i++) if (ctx->side_deep[i]) cudaStreamDestroy(ctx->side_deep[i]);
  if (ctx->green) {
    typedef CUresult (*fn_gdestroy)(CUgreenCtx);
    void* pd = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &pd, cudaEnableDefault, &q) == cudaSuccess && pd) ((fn_gdestroy)pd)((CUgreenCtx)ctx->green);
  }
  if (ctx->hi) cudaStreamDestroy(ctx->hi);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  if (ctx->zc_out) cudaStreamDestroy(ctx->zc_out);
  for (cudaEvent_t e : ctx->dep_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->io_pool) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* capital_last_error(const capital_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
capital_status_t capital_get_counters(const capital_ctx* ctx, capital_counters_t* out) {
  if (!ctx || !out) return CAPITAL_ERR_INVALID;
  *out = ctx->counters;
  return CAPITAL_OK;
}
capital_status_t capital_reset_counters(capital_ctx* ctx) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  ctx->counters = capital_counters_t{};
  return CAPITAL_OK;
}
capital_status_t capital_synchronize(capital_ctx* ctx) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaStreamSynchronize(ctx->stream));
  return CAPITAL_OK;
}
capital_status_t capital_set_stream(capital_ctx* ctx, void* stream) {
  if (!ctx || !stream) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t ns = (cudaStream_t)stream;
  if (ns == ctx->stream) return CAPITAL_OK;
  // order the new stream after the work already enqueued on the old one (workspaces are shared between calls)
  CAP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  CAP_CUDA(cudaStreamWaitEvent(ns, ctx->ev_fork, 0));
  if (ctx->own_stream) { CAP_CUDA(cudaStreamSynchronize(ctx->stream)); CAP_CUDA(cudaStreamDestroy(ctx->stream)); ctx->own_stream = false; }
  ctx->stream = ns;
  return CAPITAL_OK;
}
capital_status_t capital_release_workspace(capital_ctx* ctx) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  CAP_CUDA(cudaDeviceSynchronize());  // side / copy streams may still reference the buffers
  CAP_TRY(dist_release_peer_maps(ctx));
  for (auto& kv : ctx->pool) if (kv.second.p) CAP_CUDA(cudaFree(kv.second.p));
  ctx->pool.clear();
  if (ctx->pinned) { CAP_CUDA(cudaFreeHost(ctx->pinned)); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
  return CAPITAL_OK;
}
capital_status_t capital_last_factor_ms(const capital_ctx* ctx_, float* ms) {
  capital_ctx* ctx = const_cast<capital_ctx*>(ctx_);
  if (!ctx || !ms) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
  return CAPITAL_OK;
}

capital_status_t capital_timeline_begin(capital_ctx* ctx) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  ctx->timeline = true; ctx->tl.clear(); ctx->tl_used = 0;
  return CAPITAL_OK;
}
capital_status_t capital_timeline_end(capital_ctx* ctx, double* out, int64_t cap_records, int64_t* n_records) {
  if (!ctx || !n_records) return CAPITAL_ERR_INVALID;
  ctx->timeline = false;
  CAP_CUDA(cudaSetDevice(ctx->device));
  CAP_CUDA(cudaDeviceSynchronize());
  *n_records = (int64_t)ctx->tl.size();
  if (out && !ctx->tl.empty()) {
    // time origin: the earliest start
    cudaEvent_t base = ctx->tl[0].e0;
    for (auto& r : ctx->tl) {
      float d = 0;
      if (cudaEventElapsedTime(&d, base, r.e0) == cudaSuccess && d < 0) base = r.e0;
    }
    for (int64_t i = 0; i < *n_records && i < cap_records; i++) {
      const auto& r = ctx->tl[i];
      float t0 = 0, t1 = 0;
      CAP_CUDA(cudaEventElapsedTime(&t0, base, r.e0));
      CAP_CUDA(cudaEventElapsedTime(&t1, base, r.e1));
      double* o = out + i * 8;
      o[0] = r.sid; o[1] = r.kind; o[2] = t0; o[3] = t1; o[4] = r.a; o[5] = r.b; o[6] = r.c; o[7] = 0;
    }
  }
  return CAPITAL_OK;
}
capital_status_t capital_probe_dmma_f64(capital_ctx* ctx, double* tflops, double* ms) {
  if (!ctx || !tflops || !ms) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  return gemm_probe_dmma(ctx, tflops, ms);
}
capital_status_t capital_set_overlap(capital_ctx* ctx, int enabled) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  ctx->no_overlap = !enabled;
  return CAPITAL_OK;
}
capital_status_t capital_profile_begin(capital_ctx* ctx) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  ctx->profiling = true; ctx->prof_used = 0; ctx->prof_recs.clear();
  return CAPITAL_OK;
}
capital_status_t capital_profile_end(capital_ctx* ctx, double* kernel_ms, double* kernel_flops, int64_t* launches) {
  if (!ctx || !kernel_ms || !kernel_flops || !launches) return CAPITAL_ERR_INVALID;
  ctx->profiling = false;
  CAP_CUDA(cudaStreamSynchronize(ctx->stream));
  double ms = 0, fl = 0;
  for (auto& r : ctx->prof_recs) {
    float t = 0;
    CAP_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t; fl += r.flops;
  }
  *kernel_ms = ms; *kernel_flops = fl; *launches = (int64_t)ctx->prof_recs.size();
  ctx->prof_recs.clear(); ctx->prof_used = 0;
  return CAPITAL_OK;
}

// ---- generators -------------------------------------------------------------------------------
capital_status_t capital_distribute_symmetric_f64(capital_ctx* ctx, double* A_local, int64_t n, int diag_dom) {
  if (!ctx || !A_local || n <= 0) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  const int d = ctx->grid.d;
  const int64_t L = ceil_div(n, d);
  double* dev;
  CAP_TRY(cap_stage_out_begin(ctx, A_local, (size_t)L * L, "gen_out", &dev));
  CAP_TRY(gen_symmetric(ctx, ctx->stream, dev, L, L, L, n, ctx->grid.x, ctx->grid.y, d, diag_dom));
  CAP_TRY(cap_stage_out_end(ctx, A_local, (size_t)L * L, dev));
  CAP_CUDA(cudaStreamSynchronize(ctx->stream));
  return CAPITAL_OK;
}
capital_status_t capital_distribute_random_f64(capital_ctx* ctx, double* A_local, int64_t m, int64_t n, int64_t key) {
  if (!ctx || !A_local || n <= 0 || m <= 0) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  const int c = ctx->grid.c, d = ctx->grid.d, x = ctx->grid.x, y = ctx->grid.y;
  const int64_t lr = ceil_div(m, d), lc = ceil_div(n, c);
  // structure.hpp:110-111: the stream is consumed over the un-padded local extent only
  const int64_t pad_c = ((n % c != 0) && ((lc - 1) * c + x >= n)) ? lc - 1 : lc;
  const int64_t pad_r = ((m % d != 0) && ((lr - 1) * d + y >= m)) ? lr - 1 : lr;
  double* dev;
  CAP_TRY(cap_stage_out_begin(ctx, A_local, (size_t)lr * lc, "gen_out", &dev));
  CAP_TRY(gen_random(ctx, ctx->stream, dev, lr, lr, lc, pad_r, pad_c, key));
  CAP_TRY(cap_stage_out_end(ctx, A_local, (size_t)lr * lc, dev));
  CAP_CUDA(cudaStreamSynchronize(ctx->stream));
  return CAPITAL_OK;
}

// ---- CholInv ----------------------------------------------------------------------------------
capital_status_t capital_cholinv_factor_f64(capital_ctx* ctx, const double* A_local, int64_t n, const capital_cholinv_args_t* args,
                                            capital_structure_t ostruct, double* R_local, double* Rinv_local) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  if (!A_local || !args || !R_local || !Rinv_local || n <= 0 || args->split <= 0 || args->dir != 'U') {
    ctx->set_error("cholinv::factor: invalid arguments (split > 0 and dir == 'U' are required, cholinv.hpp:9)");
    return CAPITAL_ERR_INVALID;
  }
  CAP_CUDA(cudaSetDevice(ctx->device));
  const capital_grid_t& g = ctx->grid;
  if (g.size > 1) return dist_cholinv_factor(ctx, A_local, n, args, ostruct, R_local, Rinv_local);

  const int64_t L = n, ld = round_up(L, 16);
  const size_t out_count = ostruct == CAPITAL_UPPERTRI_PACKED ? (size_t)L * (L + 1) / 2 : (size_t)L * L;
  cudaStream_t st = ctx->stream;
  ctx->io_used = 0;
  CAP_CUDA(cudaEventRecord(ctx->ev_start, st));
  double *W, *Rm, *Ri, *RiT, *dR, *dRinv;
  CAP_TRY(ctx->workspace("W", (size_t)ld * L * 8, (void**)&W));
  CAP_TRY(ctx->workspace("Rm", (size_t)ld * L * 8, (void**)&Rm));
  CAP_TRY(ctx->workspace("Ri", (size_t)ld * L * 8, (void**)&Ri));
  CAP_TRY(ctx->workspace("RiT", (size_t)ld * L * 8, (void**)&RiT));
  // experimental block-wise output: both outputs pinned host arrays, packed, large enough to matter
  double *zR = nullptr, *zRinv = nullptr;
  if (ctx->zc_mode && ctx->zc_out && ostruct == CAPITAL_UPPERTRI_PACKED && L >= 2048 && !cap_is_device_ptr(R_local) && !cap_is_device_ptr(Rinv_local)) {
    zR = host_alias(R_local); zRinv = host_alias(Rinv_local);
    if (!zR || !zRinv) zR = zRinv = nullptr;
  }
  if (zR) { dR = dRinv = nullptr; }
  else {
    CAP_TRY(cap_stage_out_begin(ctx, R_local, out_count, "R_out", &dR));
    CAP_TRY(cap_stage_out_begin(ctx, Rinv_local, out_count, "Rinv_out", &dRinv));
  }
  CAP_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), st));
  const int64_t bc = capital_cholinv_bc_dimension(L, g.c, g.d, args->bc_mult_dim);
  if (ostruct == CAPITAL_RECT) {
    CAP_CUDA(cudaMemsetAsync(Ri, 0, (size_t)ld * L * 8, st));  // rect outputs expose everything
  } else {
    CAP_TRY(zero_band(ctx, st, L, Ri, ld));
    // the skipped top-level block of Rinv (cholinv.hpp:147) must read as zeros in the packed output -- it only exists when the top
    // node splits (same predicate as cholinv_local: a top-level base case returns the full inverse)
    if (args->complete_inv == 0 && cholinv_node_splits(L, bc, (int)args->split)) {
      const int64_t s1 = L >> args->split;
      if (s1 > 0 && s1 < L) CAP_TRY(zero_block(ctx, st, s1, L - s1, Ri + s1 * ld, ld));
    }
  }
  CAP_TRY(zero_band(ctx, st, L, RiT, ld));
  if (ostruct == CAPITAL_RECT) CAP_CUDA(cudaMemsetAsync(Rm, 0, (size_t)ld * L * 8, st));

  // Host-pointer callers: A streams in by column chunks on a copy stream while the recursion already works on the leading
  // columns (it consumes W left to right); the finished left half of R / Rinv streams out while the right half computes.
  HostIO io;
  io.ctx = ctx; io.L = L; io.ld = ld; io.Rm = Rm; io.Ri = Ri; io.dR = dR; io.dRinv = dRinv;
  io.hR = (dR != R_local) ? R_local : nullptr; io.hRinv = (dRinv != Rinv_local) ? Rinv_local : nullptr;
  io.packed = ostruct == CAPITAL_UPPERTRI_PACKED;
  io.rinv_streams = args->complete_inv == 0 && cholinv_node_splits(L, bc, (int)args->split);
  CholinvHooks hooks{&io, nullptr, nullptr};
  if (!cap_is_device_ptr(A_local)) {
    cudaEvent_t e0;
    CAP_TRY(io_event(ctx, &e0));
    CAP_CUDA(cudaEventRecord(e0, st));  // W must be free (previous users on st) before the copies land
    CAP_CUDA(cudaStreamWaitEvent(ctx->copy_in, e0, 0));
    const int64_t chunk = round_up(ceil_div(L, 16), 64);
    for (int64_t c0 = 0; c0 < L; c0 += chunk) {
      const int64_t nc = (c0 + chunk <= L) ? chunk : L - c0;
      // only the upper triangle of A is read (serialize<uppertri>(A -> R), cholinv.hpp:13): rows [0, c0 + nc) of this chunk
      const int64_t rows = c0 + nc;
      CAP_CUDA(cudaMemcpy2DAsync(W + c0 * ld, (size_t)ld * 8, A_local + c0 * L, (size_t)L * 8, (size_t)rows * 8, (size_t)nc,
                                 cudaMemcpyHostToDevice, ctx->copy_in));
      ctx->counters.h2d_bytes += rows * nc * 8;
      cudaEvent_t e;
      CAP_TRY(io_event(ctx, &e));
      CAP_CUDA(cudaEventRecord(e, ctx->copy_in));
      io.chunks.push_back({c0 + nc, e});
    }
    hooks.need_cols = hostio_need_cols;
  } else {
    CAP_TRY(copy_block(ctx, st, L, L, A_local, L, W, ld));  // serialize(A -> R), cholinv.hpp:13
  }
  io.zR = zR; io.zRinv = zRinv;
  if (zR) hooks.block_done = hostio_block_done;
  else if (io.packed && L >= 2048) hooks.left_done = hostio_left_done;  // finished column ranges are packed (and copied out) early
  if (hooks.need_cols) hooks.cols_waited = hostio_cols_waited;
  if (hooks.left_done && (io.hR || io.hRinv)) {  // host outputs: the tail of R and the top-level inverse block stream out too
    hooks.right_done = hostio_right_done;
    if (io.hRinv) hooks.inv_cols = hostio_inv_cols;
  }
  CAP_TRY(cholinv_local(ctx, st, L, W, ld, Rm, ld, Ri, ld, RiT, ld, args->complete_inv != 0, bc, (int)args->split, &hooks));
  if (zR) {  // every block has been handed to the output stream inside the recursion
    cudaEvent_t e;
    CAP_TRY(io_event(ctx, &e));
    CAP_CUDA(cudaEventRecord(e, ctx->zc_out));
    CAP_CUDA(cudaStreamWaitEvent(st, e, 0));
  } else if (ostruct == CAPITAL_UPPERTRI_PACKED) {
    {  // columns [0, c0) are already packed (and on their way to the host)
      const int64_t c0 = io.cols_out;
      const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = out_count - off;
      CAP_TRY(pack_upper(ctx, st, L, Rm, ld, dR, 0, c0, L));
      if (io.hR && cnt) { CAP_CUDA(cudaMemcpyAsync(io.hR + off, dR + off, cnt * 8, cudaMemcpyDeviceToHost, st)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
    }
    {
      const int64_t c0 = io.rinv_cols_out;
      const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = out_count - off;
      CAP_TRY(pack_upper(ctx, st, L, Ri, ld, dRinv, 0, c0, L));
      if (io.hRinv && cnt) { CAP_CUDA(cudaMemcpyAsync(io.hRinv + off, dRinv + off, cnt * 8, cudaMemcpyDeviceToHost, st)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
    }
    if (io.e_out) CAP_CUDA(cudaStreamWaitEvent(st, io.e_out, 0));  // the early D2H of the left half
  } else {
    CAP_TRY(triu_copy(ctx, st, L, Rm, ld, dR, L, 0));
    CAP_TRY(triu_copy(ctx, st, L, Ri, ld, dRinv, L, 0));
    CAP_TRY(cap_stage_out_end(ctx, R_local, out_count, dR));
    CAP_TRY(cap_stage_out_end(ctx, Rinv_local, out_count, dRinv));
  }
  CAP_CUDA(cudaEventRecord(ctx->ev_stop, st));
  return cap_check_info(ctx);
}

capital_status_t capital_cholinv_residual_f64(capital_ctx* ctx, const double* A_local, int64_t n, capital_structure_t structure,
                                              const double* R_local, double* residual) {
  if (!ctx || !A_local || !R_local || !residual || n <= 0) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  const capital_grid_t& g = ctx->grid;
  if (g.size > 1) return dist_cholinv_residual(ctx, A_local, n, structure, R_local, residual);
  const int64_t L = n, ld = round_up(L, 16);
  cudaStream_t st = ctx->stream;
  const size_t r_count = structure == CAPITAL_UPPERTRI_PACKED ? (size_t)L * (L + 1) / 2 : (size_t)L * L;
  const double *dA, *dRin;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)L * L, "A_in", &dA));
  CAP_TRY(cap_stage_in(ctx, R_local, r_count, "R_in", &dRin));
  double *E, *Rr;
  CAP_TRY(ctx->workspace("W", (size_t)ld * L * 8, (void**)&E));
  CAP_TRY(ctx->workspace("Rm", (size_t)ld * L * 8, (void**)&Rr));
  if (structure == CAPITAL_UPPERTRI_PACKED) CAP_TRY(unpack_upper(ctx, st, L, dRin, Rr, ld));
  else CAP_TRY(triu_copy(ctx, st, L, dRin, L, Rr, ld, 0));  // util::remove_triangle, validate.hpp:11
  CAP_TRY(copy_block(ctx, st, L, L, dA, L, E, ld));
  CAP_CUDA(cudaMemsetAsync(ctx->d_scalars, 0, 2 * sizeof(double), st));
  CAP_TRY(sumsq_block(ctx, st, L, L, E, ld, 1, 0, 0, 1, ctx->d_scalars + 1));  // control: sum_upper A^2
  // E = R^T R - A on the upper tiles (validate.hpp:35 with gemm(T,N,1,-1))
  CAP_TRY(gemm_tn(ctx, st, L, L, L, 1.0, Rr, ld, Rr, ld, -1.0, E, ld,
                  CAPITAL_GEMM_A_UPPER | CAPITAL_GEMM_B_UPPER | CAPITAL_GEMM_C_UPPER));
  CAP_TRY(sumsq_block(ctx, st, L, L, E, ld, 1, 0, 0, 1, ctx->d_scalars));
  double h[2];
  CAP_CUDA(cudaMemcpyAsync(h, ctx->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CAP_CUDA(cudaStreamSynchronize(st));
  *residual = sqrt(h[0]) / sqrt(h[1]);  // util.hpp:51
  return CAPITAL_OK;
}

// ---- CholeskyQR2 ------------------------------------------------------------------------------
capital_status_t capital_cacqr_factor_f64(capital_ctx* ctx, const double* A_local, int64_t m, int64_t n, int num_iter,
                                          const capital_cholinv_args_t* ci_args, capital_structure_t rstruct, double* Q_local,
                                          double* R_local) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  if (!A_local || !Q_local || !R_local || m <= 0 || n <= 0 || num_iter < 1 || num_iter > 2) {
    ctx->set_error("cacqr::factor: invalid arguments");
    return CAPITAL_ERR_INVALID;
  }
  CAP_CUDA(cudaSetDevice(ctx->device));
  return dist_cacqr_factor(ctx, A_local, m, n, num_iter, ci_args, rstruct, Q_local, R_local);
}
capital_status_t capital_cacqr_residual_f64(capital_ctx* ctx, const double* A_local, int64_t m, int64_t n, const double* Q_local,
                                            capital_structure_t rstruct, const double* R_local, double* residual,
                                            double* orthogonality) {
  if (!ctx || !A_local || !Q_local || !R_local || !residual || !orthogonality) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  return dist_cacqr_residual(ctx, A_local, m, n, Q_local, rstruct, R_local, residual, orthogonality);
}

// ---- SUMMA -----------------------------------------------------------------------------------
capital_status_t capital_summa_gemm_tn_f64(capital_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A_local,
                                           const double* B_local, double beta, double* C_local) {
  if (!ctx || !A_local || !B_local || !C_local || m <= 0 || n <= 0 || k <= 0) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  return dist_summa_gemm_tn(ctx, m, n, k, alpha, A_local, B_local, beta, C_local);
}

// ---- leaf-engine seam -------------------------------------------------------------------------
capital_status_t capital_blas_gemm_tn_f64(capital_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                                          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags) {
  if (!ctx || !A || !B || !C) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  if (!cap_is_device_ptr(A) || !cap_is_device_ptr(B) || !cap_is_device_ptr(C)) {
    ctx->set_error("capital_blas_gemm_tn_f64 takes device pointers");
    return CAPITAL_ERR_INVALID;
  }
  return gemm_tn(ctx, ctx->stream, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags);
}
// EXPERIMENTAL (BASELINE config 5): the same product on the TF32 tensor cores (tcgen05 + TMEM), FP64 in and out
capital_status_t capital_blas_gemm_tn_tf32(capital_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                                           const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int passes) {
  if (!ctx || !A || !B || !C) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  if (!cap_is_device_ptr(A) || !cap_is_device_ptr(B) || !cap_is_device_ptr(C)) {
    ctx->set_error("capital_blas_gemm_tn_tf32 takes device pointers");
    return CAPITAL_ERR_INVALID;
  }
  CAP_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream));
  CAP_TRY(gemm_tn_tf32(ctx, ctx->stream, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags, passes));
  return cap_check_info(ctx);
}
capital_status_t capital_set_trailing_precision(capital_ctx* ctx, int mode) {
  if (!ctx) return CAPITAL_ERR_INVALID;
  if (mode != 0 && mode != 1 && mode != 3) { ctx->set_error("trailing precision: 0 (FP64), 1 (TF32) or 3 (3 x TF32, split operands)"); return CAPITAL_ERR_INVALID; }
  ctx->trailing_mode = mode;
  return CAPITAL_OK;
}
capital_status_t capital_tf32_stats(const capital_ctx* ctx, int64_t* launches, double* flops) {
  if (!ctx || !launches || !flops) return CAPITAL_ERR_INVALID;
  *launches = ctx->tf32_launches; *flops = ctx->tf32_flops;
  return CAPITAL_OK;
}
capital_status_t capital_lapack_potrf_trtri_f64(capital_ctx* ctx, int64_t n, const double* A, int64_t lda, double* R, int64_t ldr,
                                                double* Rinv, int64_t ldri) {
  if (!ctx || !A || !R || !Rinv || n <= 0 || lda < n || ldr < n || ldri < n) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  if (!cap_is_device_ptr(A) || !cap_is_device_ptr(R) || !cap_is_device_ptr(Rinv)) return CAPITAL_ERR_INVALID;
  if ((ldr & 1) || (ldri & 1)) { ctx->set_error("potrf_trtri: ldr, ldri must be even"); return CAPITAL_ERR_INVALID; }
  cudaStream_t st = ctx->stream;
  const int64_t ld = round_up(n, 16);
  double *W, *RiT;
  CAP_TRY(ctx->workspace("bcW", (size_t)ld * n * 8, (void**)&W));
  CAP_TRY(ctx->workspace("bcRiT", (size_t)ld * n * 8, (void**)&RiT));
  CAP_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), st));
  CAP_TRY(copy_block(ctx, st, n, n, A, lda, W, ld));
  CAP_TRY(zero_block(ctx, st, n, n, R, ldr));
  CAP_TRY(zero_block(ctx, st, n, n, Rinv, ldri));
  CAP_CUDA(cudaMemsetAsync(RiT, 0, (size_t)ld * n * 8, st));
  CAP_TRY(cholinv_local(ctx, st, n, W, ld, R, ldr, Rinv, ldri, RiT, ld, true, n, 1));
  return cap_check_info(ctx);
}

}  // extern "C"
