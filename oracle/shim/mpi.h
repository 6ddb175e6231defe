/* TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product path.
 *
 * mini-MPI: a fork + shared-memory stand-in for <mpi.h>, just large enough to run the reference's
 * cholinv / cacqr / summa / validators on the host cores of one box (no MPI implementation exists in
 * this image). MPI_Init_thread forks MINIMPI_NP-1 children (env, default 1); ranks talk through
 * per-(src,dst) single-slot mailboxes in a MAP_SHARED region. Collectives are root loops over the
 * communicator members; non-blocking collectives execute eagerly (legal because every rank issues
 * its collectives in the same program order); MPI_Wait is a no-op.
 *
 * Covers exactly the calls the reference makes on this path:
 *   Init_thread Finalize Comm_rank Comm_size Comm_split Comm_dup Comm_free Barrier Wtime
 *   Bcast Ibcast Reduce Allreduce Iallreduce Allgather Gather Scatter Iscatter Alltoall
 *   Send Recv Sendrecv_replace Wait  (+ PMPI_Barrier / PMPI_Allreduce used by the benches).
 * MPI_Datatype values are the element size in bytes: the reference stores them in a
 * `constexpr static size_t` (src/util/shared.h:43,48), so they must be integral constants.
 */
#ifndef CAPITAL_ORACLE_MINIMPI_H
#define CAPITAL_ORACLE_MINIMPI_H
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <time.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <sched.h>
#include <algorithm>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int dummy; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_IN_PLACE ((void*)(intptr_t)-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_THREAD_SINGLE 0
#define MPI_SUCCESS 0

namespace minimpi {
enum { MAXP = 64, SLOT = 1 << 20 };
struct mailbox { volatile int full; volatile int64_t bytes; char pad[48]; char data[SLOT]; };
struct shared_t { volatile int bar_count; volatile int bar_sense; char pad[56]; mailbox box[1]; };
struct comm_t { int n; int me; std::vector<int> world; bool live; };
struct state_t {
  int np = 1, rank = 0; shared_t* sh = nullptr; std::vector<comm_t> comms; std::vector<pid_t> kids;
};
inline state_t& st() { static state_t s; return s; }
inline mailbox& box(int src, int dst) { return st().sh->box[(size_t)src * st().np + dst]; }
inline void spin() {
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}
inline void p2p_send(int dst, const void* buf, int64_t bytes) {
  const char* p = (const char*)buf; mailbox& b = box(st().rank, dst);
  do {
    int64_t n = std::min<int64_t>(bytes, SLOT);
    while (b.full) spin();
    memcpy(b.data, p, (size_t)n); b.bytes = n; __sync_synchronize(); b.full = 1;
    p += n; bytes -= n;
  } while (bytes > 0);
}
inline void p2p_recv(int src, void* buf, int64_t bytes) {
  char* p = (char*)buf; mailbox& b = box(src, st().rank);
  do {
    while (!b.full) spin();
    __sync_synchronize(); int64_t n = b.bytes; memcpy(p, b.data, (size_t)n); __sync_synchronize(); b.full = 0;
    p += n; bytes -= n;
  } while (bytes > 0);
}
inline void world_barrier() {
  shared_t* sh = st().sh; if (st().np == 1) return;
  int sense = sh->bar_sense;
  if (__sync_add_and_fetch(&sh->bar_count, 1) == st().np) { sh->bar_count = 0; __sync_synchronize(); sh->bar_sense = !sense; }
  else while (sh->bar_sense == sense) spin();
}
inline void reduce_into(void* acc, const void* in, int64_t count, MPI_Datatype dt, MPI_Op op) {
  if (dt == 8) { double* a = (double*)acc; const double* b = (const double*)in;
    for (int64_t i = 0; i < count; i++) a[i] = (op == MPI_SUM) ? a[i] + b[i] : std::max(a[i], b[i]); }
  else { float* a = (float*)acc; const float* b = (const float*)in;
    for (int64_t i = 0; i < count; i++) a[i] = (op == MPI_SUM) ? a[i] + b[i] : std::max(a[i], b[i]); }
}
}  // namespace minimpi

inline int MPI_Init_thread(int*, char***, int, int* provided) {
  using namespace minimpi; state_t& s = st();
  const char* e = getenv("MINIMPI_NP"); s.np = e ? atoi(e) : 1; if (s.np < 1 || s.np > MAXP) s.np = 1;
  size_t bytes = sizeof(shared_t) + sizeof(mailbox) * (size_t)s.np * s.np;
  s.sh = (shared_t*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (s.sh == MAP_FAILED) { perror("minimpi mmap"); exit(1); }
  s.sh->bar_count = 0; s.sh->bar_sense = 0;
  for (int i = 0; i < s.np * s.np; i++) s.sh->box[i].full = 0;
  s.rank = 0; fflush(stdout); fflush(stderr);
  for (int r = 1; r < s.np; r++) { pid_t p = fork(); if (p == 0) { s.rank = r; s.kids.clear(); break; } s.kids.push_back(p); }
  comm_t w; w.n = s.np; w.me = s.rank; w.live = true; for (int i = 0; i < s.np; i++) w.world.push_back(i);
  s.comms.clear(); s.comms.push_back(w);
  if (provided) *provided = 0; return 0;
}
inline int MPI_Init(int* a, char*** b) { return MPI_Init_thread(a, b, 0, nullptr); }
inline int MPI_Finalize() {
  using namespace minimpi; world_barrier(); fflush(stdout);
  if (st().rank != 0) _exit(0);
  for (pid_t p : st().kids) { int status; waitpid(p, &status, 0); }
  return 0;
}
inline double MPI_Wtime() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
inline int MPI_Comm_rank(MPI_Comm c, int* r) { *r = minimpi::st().comms[c].me; return 0; }
inline int MPI_Comm_size(MPI_Comm c, int* n) { *n = minimpi::st().comms[c].n; return 0; }
inline int MPI_Comm_free(MPI_Comm* c) { if (*c != 0) minimpi::st().comms[*c].live = false; return 0; }

inline int MPI_Bcast(void* buf, int64_t count, MPI_Datatype dt, int root, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c]; int64_t bytes = count * dt;
  if (cm.n == 1 || bytes == 0) return 0;
  if (cm.me == root) { for (int i = 0; i < cm.n; i++) if (i != root) p2p_send(cm.world[i], buf, bytes); }
  else p2p_recv(cm.world[root], buf, bytes);
  return 0;
}
inline int MPI_Reduce(const void* send, void* recv, int64_t count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c]; int64_t bytes = count * dt;
  if (cm.me == root) {
    if (send != MPI_IN_PLACE && send != recv) memcpy(recv, send, (size_t)bytes);
    if (cm.n > 1 && bytes > 0) { std::vector<char> tmp((size_t)bytes);
      for (int i = 0; i < cm.n; i++) if (i != root) { p2p_recv(cm.world[i], tmp.data(), bytes); reduce_into(recv, tmp.data(), count, dt, op); } }
  } else if (bytes > 0) p2p_send(cm.world[root], send == MPI_IN_PLACE ? recv : send, bytes);
  return 0;
}
inline int MPI_Allreduce(const void* send, void* recv, int64_t count, MPI_Datatype dt, MPI_Op op, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c];
  if (cm.me == 0) MPI_Reduce(send, recv, count, dt, op, 0, c);
  else MPI_Reduce(send == MPI_IN_PLACE ? recv : send, recv, count, dt, op, 0, c);
  return MPI_Bcast(recv, count, dt, 0, c);
}
inline int MPI_Gather(const void* send, int64_t scount, MPI_Datatype sdt, void* recv, int64_t rcount, MPI_Datatype rdt, int root, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c]; int64_t bytes = scount * sdt;
  if (cm.me == root) { for (int i = 0; i < cm.n; i++) { char* dst = (char*)recv + (size_t)i * rcount * rdt;
      if (i == root) { if (send != MPI_IN_PLACE) memcpy(dst, send, (size_t)bytes); } else if (bytes > 0) p2p_recv(cm.world[i], dst, bytes); } }
  else if (bytes > 0) p2p_send(cm.world[root], send, bytes);
  return 0;
}
inline int MPI_Scatter(const void* send, int64_t scount, MPI_Datatype sdt, void* recv, int64_t rcount, MPI_Datatype rdt, int root, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c]; int64_t bytes = rcount * rdt;
  if (cm.me == root) { for (int i = 0; i < cm.n; i++) { const char* src = (const char*)send + (size_t)i * scount * sdt;
      if (i == root) { if (recv != MPI_IN_PLACE) memcpy(recv, src, (size_t)bytes); } else if (bytes > 0) p2p_send(cm.world[i], src, bytes); } }
  else if (bytes > 0) p2p_recv(cm.world[root], recv, bytes);
  return 0;
}
inline int MPI_Allgather(const void* send, int64_t scount, MPI_Datatype sdt, void* recv, int64_t rcount, MPI_Datatype rdt, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c];
  MPI_Gather(send, scount, sdt, recv, rcount, rdt, 0, c);
  return MPI_Bcast(recv, rcount * cm.n, rdt, 0, c);
}
inline int MPI_Alltoall(const void* send, int64_t scount, MPI_Datatype sdt, void* recv, int64_t rcount, MPI_Datatype rdt, MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c];
  for (int r = 0; r < cm.n; r++) MPI_Gather((const char*)send + (size_t)r * scount * sdt, scount, sdt, recv, rcount, rdt, r, c);
  return 0;
}
inline int MPI_Send(const void* buf, int64_t count, MPI_Datatype dt, int dst, int, MPI_Comm c) {
  using namespace minimpi; p2p_send(st().comms[c].world[dst], buf, count * dt); return 0; }
inline int MPI_Recv(void* buf, int64_t count, MPI_Datatype dt, int src, int, MPI_Comm c, MPI_Status*) {
  using namespace minimpi; p2p_recv(st().comms[c].world[src], buf, count * dt); return 0; }
inline int MPI_Sendrecv_replace(void* buf, int64_t count, MPI_Datatype dt, int dst, int, int src, int, MPI_Comm c, MPI_Status*) {
  using namespace minimpi; comm_t& cm = st().comms[c]; int64_t bytes = count * dt;
  if (dst == cm.me && src == cm.me) return 0;
  if (bytes == 0) return 0;
  std::vector<char> tmp((size_t)bytes);
  if (cm.me < dst) { p2p_send(cm.world[dst], buf, bytes); p2p_recv(cm.world[src], tmp.data(), bytes); }
  else { p2p_recv(cm.world[src], tmp.data(), bytes); p2p_send(cm.world[dst], buf, bytes); }
  memcpy(buf, tmp.data(), (size_t)bytes); return 0;
}
inline int MPI_Barrier(MPI_Comm c) {
  using namespace minimpi; comm_t& cm = st().comms[c];
  if (c == 0) { world_barrier(); return 0; }
  char t = 0; std::vector<char> g((size_t)cm.n);  // gather then bcast one byte
  MPI_Gather(&t, 1, 1, g.data(), 1, 1, 0, c);
  return MPI_Bcast(&t, 1, 1, 0, c);
}
inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* out) {
  using namespace minimpi; comm_t cm = st().comms[c];
  std::vector<int> mine = {color, key, cm.me}; std::vector<int> all((size_t)3 * cm.n);
  MPI_Allgather(mine.data(), 3, 4, all.data(), 3, 4, c);
  std::vector<std::pair<std::pair<int,int>,int>> members;  // ((key, parent rank), world rank)
  for (int i = 0; i < cm.n; i++) if (all[3*i] == color) members.push_back({{all[3*i+1], all[3*i+2]}, cm.world[all[3*i+2]]});
  std::sort(members.begin(), members.end());
  comm_t nc; nc.n = (int)members.size(); nc.live = true; nc.me = -1;
  for (int i = 0; i < nc.n; i++) { nc.world.push_back(members[i].second); if (members[i].second == st().rank) nc.me = i; }
  st().comms.push_back(nc); *out = (int)st().comms.size() - 1; return 0;
}
inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* out) {
  using namespace minimpi; comm_t nc = st().comms[c]; st().comms.push_back(nc); *out = (int)st().comms.size() - 1; return 0; }
inline int MPI_Wait(MPI_Request*, MPI_Status*) { return 0; }
inline int MPI_Ibcast(void* b, int64_t n, MPI_Datatype dt, int root, MPI_Comm c, MPI_Request*) { return MPI_Bcast(b, n, dt, root, c); }
inline int MPI_Iallreduce(const void* s, void* r, int64_t n, MPI_Datatype dt, MPI_Op op, MPI_Comm c, MPI_Request*) { return MPI_Allreduce(s, r, n, dt, op, c); }
inline int MPI_Iscatter(const void* s, int64_t sc, MPI_Datatype sdt, void* r, int64_t rc, MPI_Datatype rdt, int root, MPI_Comm c, MPI_Request*) {
  return MPI_Scatter(s, sc, sdt, r, rc, rdt, root, c); }
#define PMPI_Barrier MPI_Barrier
#define PMPI_Allreduce MPI_Allreduce
#endif
