/* mpi.h -- minimal MPI stand-in for building the UNMODIFIED lanl/branson
 * reference as a parity oracle (test infrastructure only; never linked into the
 * product library).
 *
 * No MPI implementation exists in this image (no mpicxx / libmpi), yet every
 * reference header includes <mpi.h> (reference src/config.h.in:16).  This shim
 * provides exactly the symbols the reference's replicated-mode path executes:
 *
 *   - 1 rank (default): collectives are copies.
 *   - N ranks (env BRANSON_SHIM_NRANKS=N): MPI_Init forks N-1 children that
 *     share an anonymous mmap region; Allreduce is a RANK-ORDERED sum (rank 0
 *     first), which is how the n-rank oracle is *defined* (real MPI leaves the
 *     order unspecified); Isend/Irecv go through per-pair FIFO mailboxes
 *     (enough for replicate_mesh, reference src/decompose_mesh.h:705-736).
 *
 * RMA, Iallreduce, blocking Send/Recv and friends are only reached from the
 * domain-decomposed path that the reference itself disables
 * (src/particle_pass_transport.h:141-142); they abort here.
 *
 * Must be included from exactly one translation unit per binary (it defines
 * static state), which matches the reference's single-TU build.
 */
#ifndef BRANSON_ORACLE_MPI_SHIM_H
#define BRANSON_ORACLE_MPI_SHIM_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_Win;
typedef long MPI_Aint;
struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  int shim_count_bytes;
};

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LAND 4
/* basic datatypes: handle = index in the size table below */
#define MPI_DOUBLE 1
#define MPI_UNSIGNED 2
#define MPI_INT 3
#define MPI_UNSIGNED_LONG 4
#define MPI_UNSIGNED_CHAR 5
#define MPI_C_BOOL 6
#define MPI_FLOAT 7
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_INFO_NULL 0
#define MPI_PROC_NULL (-2)
#define MPI_ANY_SOURCE (-1)
#define MPI_MODE_NOCHECK 0
#define MPI_COMM_TYPE_SHARED 0
#define MPI_MAX_PROCESSOR_NAME 64

namespace mpishim {

enum { MAX_TYPES = 64, MAX_REQS = 4096, MAX_RANKS = 16 };

struct Mailbox { /* single-producer single-consumer byte FIFO in shared memory */
  volatile uint64_t head; /* bytes consumed */
  volatile uint64_t tail; /* bytes produced */
};

struct Shared {
  pthread_barrier_t barrier;
  int n_ranks;
  size_t slot_bytes;  /* per-rank collective slot */
  size_t mbox_bytes;  /* per-pair mailbox payload capacity */
};

struct Req {
  int kind; /* 0 free/complete, 2 recv pending */
  void *buf;
  size_t bytes;
  int src;
};

static int g_rank = 0;
static int g_n = 1;
static Shared *g_sh = nullptr;
static char *g_slots = nullptr;   /* n * slot_bytes */
static char *g_mboxes = nullptr;  /* n*n * (sizeof(Mailbox)+mbox_bytes) */
static size_t g_type_size[MAX_TYPES] = {0, 8, 4, 4, 8, 1, 1, 4};
static int g_n_types = 8;
static Req g_reqs[MAX_REQS];
static pid_t g_children[MAX_RANKS];

static inline void die(const char *what) {
  std::fprintf(stderr, "[mpi shim] %s is not supported by the oracle shim\n", what);
  std::abort();
}
static inline size_t tsize(MPI_Datatype t) {
  if (t <= 0 || t >= g_n_types) die("unknown datatype");
  return g_type_size[t];
}
static inline void barrier() {
  if (g_n > 1) pthread_barrier_wait(&g_sh->barrier);
}
static inline Mailbox *mbox(int src, int dst) {
  size_t stride = sizeof(Mailbox) + g_sh->mbox_bytes;
  return (Mailbox *)(g_mboxes + ((size_t)src * g_n + dst) * stride);
}
static inline void mbox_put(int dst, const void *data, size_t bytes) {
  Mailbox *m = mbox(g_rank, dst);
  char *payload = (char *)(m + 1);
  size_t cap = g_sh->mbox_bytes;
  const char *p = (const char *)data;
  size_t left = bytes;
  while (left) {
    while (m->tail - m->head >= cap) usleep(50);
    size_t off = m->tail % cap;
    size_t room = cap - (m->tail - m->head);
    size_t n = left < room ? left : room;
    if (n > cap - off) n = cap - off;
    std::memcpy(payload + off, p, n);
    __sync_synchronize();
    m->tail += n;
    p += n;
    left -= n;
  }
}
static inline void mbox_get(int src, void *data, size_t bytes) {
  Mailbox *m = mbox(src, g_rank);
  char *payload = (char *)(m + 1);
  size_t cap = g_sh->mbox_bytes;
  char *p = (char *)data;
  size_t left = bytes;
  while (left) {
    while (m->tail == m->head) usleep(50);
    __sync_synchronize();
    size_t off = m->head % cap;
    size_t avail = m->tail - m->head;
    size_t n = left < avail ? left : avail;
    if (n > cap - off) n = cap - off;
    std::memcpy(p, payload + off, n);
    __sync_synchronize();
    m->head += n;
    p += n;
    left -= n;
  }
}

template <typename T>
static inline void reduce_typed(void *out, int count, MPI_Op op) {
  T *o = (T *)out;
  for (int i = 0; i < count; ++i) {
    T acc = ((const T *)(g_slots))[i];
    for (int r = 1; r < g_n; ++r) {
      T v = ((const T *)(g_slots + (size_t)r * g_sh->slot_bytes))[i];
      if (op == MPI_SUM) acc = acc + v;
      else if (op == MPI_MAX) acc = (v > acc) ? v : acc;
      else if (op == MPI_MIN) acc = (v < acc) ? v : acc;
      else die("reduction op");
    }
    o[i] = acc;
  }
}
} // namespace mpishim

static inline int MPI_Init(int *, char ***) {
  using namespace mpishim;
  const char *e = std::getenv("BRANSON_SHIM_NRANKS");
  int n = e ? std::atoi(e) : 1;
  if (n < 1 || n > MAX_RANKS) die("BRANSON_SHIM_NRANKS out of range");
  g_n = n;
  g_rank = 0;
  if (n == 1) return 0;
  const char *sb = std::getenv("BRANSON_SHIM_SLOT_MB");
  const char *mb = std::getenv("BRANSON_SHIM_MBOX_MB");
  size_t slot_bytes = (size_t)(sb ? std::atoi(sb) : 64) << 20;
  size_t mbox_bytes = (size_t)(mb ? std::atoi(mb) : 128) << 20;
  size_t mstride = sizeof(Mailbox) + mbox_bytes;
  size_t total = 4096 + (size_t)n * slot_bytes + (size_t)n * n * mstride;
  void *base = mmap(nullptr, total, PROT_READ | PROT_WRITE,
                    MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (base == MAP_FAILED) die("mmap of shared region");
  g_sh = (Shared *)base;
  g_slots = (char *)base + 4096;
  g_mboxes = g_slots + (size_t)n * slot_bytes;
  g_sh->n_ranks = n;
  g_sh->slot_bytes = slot_bytes;
  g_sh->mbox_bytes = mbox_bytes;
  pthread_barrierattr_t a;
  pthread_barrierattr_init(&a);
  pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&g_sh->barrier, &a, n);
  std::fflush(stdout);
  for (int r = 1; r < n; ++r) {
    pid_t p = fork();
    if (p < 0) die("fork");
    if (p == 0) {
      g_rank = r;
      /* children stay quiet unless asked otherwise */
      if (!std::getenv("BRANSON_SHIM_CHILD_STDOUT")) {
        if (!std::freopen("/dev/null", "w", stdout)) die("freopen");
      }
      break;
    }
    g_children[r] = p;
  }
  return 0;
}
static inline int MPI_Finalize() {
  using namespace mpishim;
  barrier();
  if (g_n > 1) {
    if (g_rank != 0) {
      std::fflush(stdout);
      _exit(0);
    }
    for (int r = 1; r < g_n; ++r) {
      int st = 0;
      waitpid(g_children[r], &st, 0);
    }
  }
  return 0;
}
static inline int MPI_Abort(MPI_Comm, int code) {
  std::fflush(stdout);
  std::_Exit(code ? code : 1);
  return 0;
}
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = mpishim::g_rank; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = mpishim::g_n; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm *o) { *o = c; return 0; }
static inline int MPI_Comm_split_type(MPI_Comm c, int, int, MPI_Info, MPI_Comm *o) { *o = c; return 0; }
static inline int MPI_Barrier(MPI_Comm) { mpishim::barrier(); return 0; }
static inline int MPI_Get_processor_name(char *name, int *len) {
  std::strcpy(name, "oracle-shim");
  *len = (int)std::strlen(name);
  return 0;
}

static inline int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype t, MPI_Op op, MPI_Comm) {
  using namespace mpishim;
  size_t bytes = (size_t)count * tsize(t);
  if (g_n == 1) {
    if (send != MPI_IN_PLACE && send != recv) std::memcpy(recv, send, bytes);
    return 0;
  }
  if (bytes > g_sh->slot_bytes) die("Allreduce larger than BRANSON_SHIM_SLOT_MB");
  const void *src = (send == MPI_IN_PLACE) ? recv : send;
  std::memcpy(g_slots + (size_t)g_rank * g_sh->slot_bytes, src, bytes);
  barrier();
  switch (t) {
  case MPI_DOUBLE: reduce_typed<double>(recv, count, op); break;
  case MPI_UNSIGNED: reduce_typed<unsigned>(recv, count, op); break;
  case MPI_INT: reduce_typed<int>(recv, count, op); break;
  case MPI_UNSIGNED_LONG: reduce_typed<unsigned long>(recv, count, op); break;
  default: die("Allreduce datatype");
  }
  barrier();
  return 0;
}
static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm) {
  using namespace mpishim;
  if (g_n == 1) return 0;
  size_t bytes = (size_t)count * tsize(t);
  if (bytes > g_sh->slot_bytes) die("Bcast larger than BRANSON_SHIM_SLOT_MB");
  if (g_rank == root) std::memcpy(g_slots, buf, bytes);
  barrier();
  if (g_rank != root) std::memcpy(buf, g_slots, bytes);
  barrier();
  return 0;
}

static inline int MPI_Type_create_struct(int n, const int *blocklen, const MPI_Aint *disp,
                                         const MPI_Datatype *types, MPI_Datatype *out) {
  using namespace mpishim;
  size_t extent = 0, align = 1;
  for (int i = 0; i < n; ++i) {
    size_t s = tsize(types[i]);
    size_t end = (size_t)disp[i] + (size_t)blocklen[i] * s;
    if (end > extent) extent = end;
    if (s > align) align = s;
  }
  extent = (extent + align - 1) / align * align;
  if (g_n_types >= MAX_TYPES) die("too many derived datatypes");
  g_type_size[g_n_types] = extent;
  *out = g_n_types++;
  return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_size(MPI_Datatype t, int *s) { *s = (int)mpishim::tsize(t); return 0; }
static inline int MPI_Type_dup(MPI_Datatype t, MPI_Datatype *o) { *o = t; return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }

static inline int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dst, int, MPI_Comm, MPI_Request *req) {
  using namespace mpishim;
  if (g_n == 1) die("Isend at 1 rank");
  mbox_put(dst, buf, (size_t)count * tsize(t));
  *req = -1; /* already complete */
  return 0;
}
static inline int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int, MPI_Comm, MPI_Request *req) {
  using namespace mpishim;
  if (g_n == 1) die("Irecv at 1 rank");
  if (src < 0) die("Irecv from MPI_ANY_SOURCE");
  for (int i = 0; i < MAX_REQS; ++i) {
    if (g_reqs[i].kind == 0) {
      g_reqs[i].kind = 2;
      g_reqs[i].buf = buf;
      g_reqs[i].bytes = (size_t)count * tsize(t);
      g_reqs[i].src = src;
      *req = i;
      return 0;
    }
  }
  die("request table full");
  return 1;
}
static inline int MPI_Wait(MPI_Request *req, MPI_Status *) {
  using namespace mpishim;
  if (*req >= 0 && g_reqs[*req].kind == 2) {
    mbox_get(g_reqs[*req].src, g_reqs[*req].buf, g_reqs[*req].bytes);
    g_reqs[*req].kind = 0;
  }
  *req = -1;
  return 0;
}
static inline int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *) {
  for (int i = 0; i < n; ++i) MPI_Wait(&reqs[i], MPI_STATUS_IGNORE);
  return 0;
}

/* ---- never executed in replicated mode: abort loudly if reached ---- */
static inline int MPI_Test(MPI_Request *, int *, MPI_Status *) { mpishim::die("MPI_Test"); return 1; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { mpishim::die("MPI_Send"); return 1; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { mpishim::die("MPI_Recv"); return 1; }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { mpishim::die("MPI_Get_count"); return 1; }
static inline int MPI_Iallreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm, MPI_Request *) { mpishim::die("MPI_Iallreduce"); return 1; }
static inline int MPI_Info_create(MPI_Info *i) { *i = 0; return 0; }
static inline int MPI_Info_set(MPI_Info, const char *, const char *) { return 0; }
static inline int MPI_Win_allocate(MPI_Aint, int, MPI_Info, MPI_Comm, void *, MPI_Win *) { mpishim::die("MPI_Win_allocate"); return 1; }
static inline int MPI_Win_free(MPI_Win *) { mpishim::die("MPI_Win_free"); return 1; }
static inline int MPI_Win_lock_all(int, MPI_Win) { mpishim::die("MPI_Win_lock_all"); return 1; }
static inline int MPI_Win_unlock_all(MPI_Win) { mpishim::die("MPI_Win_unlock_all"); return 1; }
static inline int MPI_Win_flush_all(MPI_Win) { mpishim::die("MPI_Win_flush_all"); return 1; }
static inline int MPI_Win_sync(MPI_Win) { mpishim::die("MPI_Win_sync"); return 1; }
static inline int MPI_Put(const void *, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win) { mpishim::die("MPI_Put"); return 1; }
static inline int MPI_Rget(void *, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win, MPI_Request *) { mpishim::die("MPI_Rget"); return 1; }
static inline int MPI_Raccumulate(const void *, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Op, MPI_Win, MPI_Request *) { mpishim::die("MPI_Raccumulate"); return 1; }

#endif /* BRANSON_ORACLE_MPI_SHIM_H */
