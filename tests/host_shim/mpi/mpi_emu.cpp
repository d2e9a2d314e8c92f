// TEST INFRASTRUCTURE.  The MPI interface the reference is written against (its own src/STUBS/mpi.h: MPI_Comm,
// MPI_Datatype, MPI_Request as ints, a one-field MPI_Status), implemented for SEVERAL processes of one host, so
// that the unmodified reference objects — and the /cuda plugin linked with them — run on more than one rank in
// an image that has no MPI (plugin/Makefile: spk_cuda_mpi, spk_ref_mpi; tests/test_emulated_library.py).
// Linked in place of the reference's single-process stub (oracle/_ref/obj/mpi_stub.o).
//
// Transport: one small file per message in the directory SPK_MPI_DIR (a tmpfs), named by communicator, source,
// destination, tag and a per-(communicator, source, destination, tag) sequence number, written under a temporary
// name and renamed; a receive polls for the next number.  Sends are buffered (they never wait for the receiver),
// collectives are built from point-to-point messages through rank 0 of the communicator.  Slow and simple: the
// lattices of the tests are small.  Ranks: SPK_MPI_RANK / SPK_MPI_SIZE.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "mpi.h"  // the reference's src/STUBS/mpi.h

namespace {

int g_rank = 0, g_size = 1, g_init = 0;
std::string g_dir;
constexpr int WAIT_LIMIT_S = 120;
constexpr int TAG_COLL = -7000;  // internal tags of the collectives

struct Comm { std::vector<int> world; int me; };  // members as world ranks, my index
std::vector<Comm> comms;
std::map<std::tuple<int, int, int, int>, long> send_seq, recv_seq;  // (comm, src, dst, tag) -> next number
struct Request { void *buf; int bytes, src, tag, comm; bool active; };
std::vector<Request> requests;
std::map<const MPI_Status *, int> status_bytes;

size_t type_size(int t) {
  switch (t) {
    case MPI_INT: return sizeof(int);
    case MPI_FLOAT: return sizeof(float);
    case MPI_DOUBLE: return sizeof(double);
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_LONG_LONG: return sizeof(long long);
    case MPI_DOUBLE_INT: return sizeof(double) + sizeof(int) + 4;  // struct {double; int;} padded to 16
    default: fprintf(stderr, "mpi_emu: unknown datatype %d\n", t); abort();
  }
}

void die(const char *what) {
  fprintf(stderr, "mpi_emu: rank %d: %s\n", g_rank, what);
  fflush(NULL);
  if (!g_dir.empty()) { const int fd = open((g_dir + "/ABORT").c_str(), O_CREAT | O_WRONLY, 0600); if (fd >= 0) close(fd); }
  _exit(1);
}

std::string msg_name(int comm, int src, int dst, int tag, long seq) {
  char b[160];
  snprintf(b, sizeof b, "/m_%d_%d_%d_%d_%ld", comm, src, dst, tag, seq);
  return g_dir + b;
}

// src, dst: WORLD ranks
void put(int comm, int dst, int tag, const void *buf, size_t bytes) {
  const long seq = send_seq[std::make_tuple(comm, g_rank, dst, tag)]++;
  const std::string fin = msg_name(comm, g_rank, dst, tag, seq), tmp = fin + ".tmp";
  const int fd = open(tmp.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0600);
  if (fd < 0) die("cannot create a message file");
  size_t off = 0;
  while (off < bytes) {
    const ssize_t w = write(fd, (const char *)buf + off, bytes - off);
    if (w <= 0) die("write of a message failed");
    off += (size_t)w;
  }
  close(fd);
  if (rename(tmp.c_str(), fin.c_str()) != 0) die("rename of a message failed");
}

bool try_get(int comm, int src, int tag, void *buf, size_t cap, size_t *bytes) {
  const auto key = std::make_tuple(comm, src, g_rank, tag);
  const std::string f = msg_name(comm, src, g_rank, tag, recv_seq[key]);
  const int fd = open(f.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  fstat(fd, &st);
  if ((size_t)st.st_size > cap) die("message longer than the receive buffer");
  size_t off = 0;
  while (off < (size_t)st.st_size) {
    const ssize_t r = read(fd, (char *)buf + off, (size_t)st.st_size - off);
    if (r <= 0) die("read of a message failed");
    off += (size_t)r;
  }
  close(fd);
  unlink(f.c_str());
  recv_seq[key]++;
  *bytes = off;
  return true;
}

// src: WORLD rank or -1 (any member of the communicator); returns the source's world rank
int get(int comm, int src, int tag, void *buf, size_t cap, size_t *bytes) {
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  const std::string abort_file = g_dir + "/ABORT";
  for (long spin = 0;; spin++) {
    if (src >= 0) {
      if (try_get(comm, src, tag, buf, cap, bytes)) return src;
    } else {
      for (int s : comms[(size_t)comm].world)
        if (try_get(comm, s, tag, buf, cap, bytes)) return s;
    }
    if (spin > 50) usleep(spin > 2000 ? 500 : 50);
    if ((spin & 255) == 255) {
      if (access(abort_file.c_str(), F_OK) == 0) { fflush(NULL); _exit(1); }
      timespec t;
      clock_gettime(CLOCK_MONOTONIC, &t);
      if (t.tv_sec - t0.tv_sec > WAIT_LIMIT_S) die("a receive waited 120 s: the ranks do not make matching calls");
    }
  }
}

Comm &C(int comm) {
  if (comm < 0 || (size_t)comm >= comms.size()) die("invalid communicator");
  return comms[(size_t)comm];
}

template <class T> void reduce_t(T *acc, const T *in, int n, int op) {
  for (int i = 0; i < n; i++) {
    if (op == MPI_SUM) acc[i] = acc[i] + in[i];
    else if (op == MPI_MAX) acc[i] = in[i] > acc[i] ? in[i] : acc[i];
    else if (op == MPI_MIN) acc[i] = in[i] < acc[i] ? in[i] : acc[i];
    else if (op == MPI_LOR) acc[i] = (T)((acc[i] != 0) || (in[i] != 0));
    else die("reduction operation not supported for this type");
  }
}
struct DoubleInt { double d; int i; };
void reduce_any(void *acc, const void *in, int n, int type, int op) {
  switch (type) {
    case MPI_INT: reduce_t((int *)acc, (const int *)in, n, op); break;
    case MPI_FLOAT: reduce_t((float *)acc, (const float *)in, n, op); break;
    case MPI_DOUBLE: reduce_t((double *)acc, (const double *)in, n, op); break;
    case MPI_CHAR: case MPI_BYTE: reduce_t((char *)acc, (const char *)in, n, op); break;
    case MPI_LONG_LONG: reduce_t((long long *)acc, (const long long *)in, n, op); break;
    case MPI_DOUBLE_INT: {
      DoubleInt *a = (DoubleInt *)acc;
      const DoubleInt *b = (const DoubleInt *)in;
      for (int i = 0; i < n; i++) {
        const bool take = op == MPI_MAXLOC ? (b[i].d > a[i].d || (b[i].d == a[i].d && b[i].i < a[i].i))
                                           : (b[i].d < a[i].d || (b[i].d == a[i].d && b[i].i < a[i].i));
        if (op != MPI_MAXLOC && op != MPI_MINLOC) die("MPI_DOUBLE_INT needs MAXLOC / MINLOC");
        if (take) a[i] = b[i];
      }
      break;
    }
    default: die("unknown datatype in a reduction");
  }
}

// everybody's `bytes` bytes gathered at comm rank 0, in comm rank order; returns them there (empty elsewhere)
std::vector<char> gather0(int comm, const void *mine, size_t bytes, int tag) {
  Comm &c = C(comm);
  std::vector<char> all;
  if (c.me == 0) {
    all.resize(bytes * c.world.size());
    memcpy(all.data(), mine, bytes);
    for (size_t r = 1; r < c.world.size(); r++) {
      size_t got = 0;
      get(comm, c.world[r], tag, all.data() + r * bytes, bytes, &got);
      if (got != bytes) die("collective called with different sizes on different ranks");
    }
  } else {
    put(comm, c.world[0], tag, mine, bytes);
  }
  return all;
}
void bcast0(int comm, void *buf, size_t bytes, int tag) {
  Comm &c = C(comm);
  if (c.me == 0) {
    for (size_t r = 1; r < c.world.size(); r++) put(comm, c.world[r], tag, buf, bytes);
  } else {
    size_t got = 0;
    get(comm, c.world[0], tag, buf, bytes, &got);
    if (got != bytes) die("broadcast with different sizes on different ranks");
  }
}

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) {
  const char *r = getenv("SPK_MPI_RANK"), *s = getenv("SPK_MPI_SIZE"), *d = getenv("SPK_MPI_DIR");
  g_rank = r ? atoi(r) : 0;
  g_size = s ? atoi(s) : 1;
  g_dir = d ? d : "";
  if (g_size > 1 && g_dir.empty()) { fprintf(stderr, "mpi_emu: SPK_MPI_DIR is not set\n"); _exit(1); }
  Comm w;
  for (int i = 0; i < g_size; i++) w.world.push_back(i);
  w.me = g_rank;
  comms.clear();
  comms.push_back(w);
  g_init = 1;
  return MPI_SUCCESS;
}
int MPI_Initialized(int *flag) { *flag = g_init; return MPI_SUCCESS; }
void MPI_Get_processor_name(char *name, int *len) { strcpy(name, "localhost"); *len = 9; }
int MPI_Comm_rank(MPI_Comm comm, int *me) { *me = C(comm).me; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *n) { *n = (int)C(comm).world.size(); return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) {
  fflush(NULL);
  if (!g_dir.empty()) { const int fd = open((g_dir + "/ABORT").c_str(), O_CREAT | O_WRONLY, 0600); if (fd >= 0) close(fd); }
  _exit(code ? code : 1);
}
int MPI_Finalize() { if (g_size > 1) MPI_Barrier(MPI_COMM_WORLD); g_init = 0; return MPI_SUCCESS; }
double MPI_Wtime() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1.0e-9 * (double)t.tv_nsec;
}
int MPI_Type_size(int type, int *size) { *size = (int)type_size(type); return MPI_SUCCESS; }

int MPI_Send(void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
  put(comm, C(comm).world[(size_t)dest], tag, buf, (size_t)count * type_size(type));
  return MPI_SUCCESS;
}
int MPI_Rsend(void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
  return MPI_Send(buf, count, type, dest, tag, comm);
}
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status *status) {
  Comm &c = C(comm);
  size_t got = 0;
  const int w = get(comm, source == MPI_ANY_SOURCE ? -1 : c.world[(size_t)source], tag, buf, (size_t)count * type_size(type), &got);
  if (status) {
    status->MPI_SOURCE = (int)(std::find(c.world.begin(), c.world.end(), w) - c.world.begin());
    status_bytes[status] = (int)got;
  }
  return MPI_SUCCESS;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *request) {
  Request r = {buf, (int)((size_t)count * type_size(type)), source, tag, comm, true};
  for (size_t i = 0; i < requests.size(); i++)
    if (!requests[i].active) { requests[i] = r; *request = (int)i; return MPI_SUCCESS; }
  requests.push_back(r);
  *request = (int)requests.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *request, MPI_Status *status) {
  Request &r = requests[(size_t)*request];
  if (!r.active) return MPI_SUCCESS;
  Comm &c = C(r.comm);
  size_t got = 0;
  const int w = get(r.comm, r.src == MPI_ANY_SOURCE ? -1 : c.world[(size_t)r.src], r.tag, r.buf, (size_t)r.bytes, &got);
  if (status) {
    status->MPI_SOURCE = (int)(std::find(c.world.begin(), c.world.end(), w) - c.world.begin());
    status_bytes[status] = (int)got;
  }
  r.active = false;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request *request, MPI_Status *status) {
  for (int i = 0; i < n; i++) MPI_Wait(&request[i], status ? &status[i] : nullptr);
  return MPI_SUCCESS;
}
int MPI_Waitany(int count, MPI_Request *request, int *index, MPI_Status *status) {
  for (;;) {
    bool any = false;
    for (int i = 0; i < count; i++) {
      Request &r = requests[(size_t)request[i]];
      if (!r.active) continue;
      any = true;
      Comm &c = C(r.comm);
      size_t got = 0;
      bool ok = false;
      int w = -1;
      if (r.src == MPI_ANY_SOURCE) {
        for (int s : c.world)
          if (try_get(r.comm, s, r.tag, r.buf, (size_t)r.bytes, &got)) { ok = true; w = s; break; }
      } else {
        w = c.world[(size_t)r.src];
        ok = try_get(r.comm, w, r.tag, r.buf, (size_t)r.bytes, &got);
      }
      if (ok) {
        if (status) {
          status->MPI_SOURCE = (int)(std::find(c.world.begin(), c.world.end(), w) - c.world.begin());
          status_bytes[status] = (int)got;
        }
        r.active = false;
        *index = i;
        return MPI_SUCCESS;
      }
    }
    if (!any) { *index = -1; return MPI_SUCCESS; }
    usleep(50);
  }
}
int MPI_Sendrecv(void *sbuf, int scount, MPI_Datatype stype, int dest, int stag, void *rbuf, int rcount, MPI_Datatype rtype,
                 int source, int rtag, MPI_Comm comm, MPI_Status *status) {
  MPI_Send(sbuf, scount, stype, dest, stag, comm);
  return MPI_Recv(rbuf, rcount, rtype, source, rtag, comm, status);
}
int MPI_Get_count(MPI_Status *status, MPI_Datatype type, int *count) {
  auto it = status_bytes.find(status);
  *count = it == status_bytes.end() ? 0 : (int)((size_t)it->second / type_size(type));
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) {
  char z = 0;
  gather0(comm, &z, 1, TAG_COLL - 1);
  bcast0(comm, &z, 1, TAG_COLL - 1);
  return MPI_SUCCESS;
}
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  Comm &c = C(comm);
  const size_t bytes = (size_t)count * type_size(type);
  if (c.me == root) {
    for (size_t r = 0; r < c.world.size(); r++)
      if ((int)r != root) put(comm, c.world[r], TAG_COLL - 2, buf, bytes);
  } else {
    size_t got = 0;
    get(comm, c.world[(size_t)root], TAG_COLL - 2, buf, bytes, &got);
  }
  return MPI_SUCCESS;
}
int MPI_Allreduce(void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  Comm &c = C(comm);
  const size_t bytes = (size_t)count * type_size(type);
  std::vector<char> all = gather0(comm, sendbuf, bytes, TAG_COLL - 3);
  std::vector<char> acc(bytes ? bytes : 1);
  if (c.me == 0) {
    memcpy(acc.data(), all.data(), bytes);
    for (size_t r = 1; r < c.world.size(); r++) reduce_any(acc.data(), all.data() + r * bytes, count, type, op);
  }
  bcast0(comm, acc.data(), bytes, TAG_COLL - 3);
  memcpy(recvbuf, acc.data(), bytes);
  return MPI_SUCCESS;
}
int MPI_Reduce(void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm) {
  Comm &c = C(comm);
  const size_t bytes = (size_t)count * type_size(type);
  std::vector<char> tmp(bytes ? bytes : 1);
  MPI_Allreduce(sendbuf, tmp.data(), count, type, op, comm);
  if (c.me == root) memcpy(recvbuf, tmp.data(), bytes);
  return MPI_SUCCESS;
}
int MPI_Scan(void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  Comm &c = C(comm);
  const size_t bytes = (size_t)count * type_size(type);
  std::vector<char> all = gather0(comm, sendbuf, bytes, TAG_COLL - 4);
  if (c.me == 0) {
    std::vector<char> acc(all.begin(), all.begin() + (long)bytes);
    memcpy(recvbuf, acc.data(), bytes);
    for (size_t r = 1; r < c.world.size(); r++) {
      reduce_any(acc.data(), all.data() + r * bytes, count, type, op);
      put(comm, c.world[r], TAG_COLL - 5, acc.data(), bytes);
    }
  } else {
    size_t got = 0;
    get(comm, c.world[0], TAG_COLL - 5, recvbuf, bytes, &got);
  }
  return MPI_SUCCESS;
}
int MPI_Allgatherv(void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int *recvcounts, int *displs,
                   MPI_Datatype recvtype, MPI_Comm comm) {
  Comm &c = C(comm);
  const size_t es = type_size(recvtype);
  // every rank sends its block to every other rank (buffered), then collects
  for (size_t r = 0; r < c.world.size(); r++)
    if ((int)r != c.me) put(comm, c.world[r], TAG_COLL - 6, sendbuf, (size_t)sendcount * type_size(sendtype));
  memcpy((char *)recvbuf + (size_t)displs[c.me] * es, sendbuf, (size_t)sendcount * type_size(sendtype));
  for (size_t r = 0; r < c.world.size(); r++)
    if ((int)r != c.me) {
      size_t got = 0;
      get(comm, c.world[r], TAG_COLL - 6, (char *)recvbuf + (size_t)displs[r] * es, (size_t)recvcounts[r] * es, &got);
    }
  return MPI_SUCCESS;
}
int MPI_Allgather(void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount, MPI_Datatype recvtype,
                  MPI_Comm comm) {
  Comm &c = C(comm);
  std::vector<int> counts(c.world.size(), recvcount), displs(c.world.size());
  for (size_t r = 0; r < c.world.size(); r++) displs[r] = (int)r * recvcount;
  return MPI_Allgatherv(sendbuf, sendcount, sendtype, recvbuf, counts.data(), displs.data(), recvtype, comm);
}
int MPI_Gatherv(void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int *recvcounts, int *displs,
                MPI_Datatype recvtype, int root, MPI_Comm comm) {
  Comm &c = C(comm);
  if (c.me != root) {
    put(comm, c.world[(size_t)root], TAG_COLL - 7, sendbuf, (size_t)sendcount * type_size(sendtype));
    return MPI_SUCCESS;
  }
  const size_t es = type_size(recvtype);
  memcpy((char *)recvbuf + (size_t)displs[c.me] * es, sendbuf, (size_t)sendcount * type_size(sendtype));
  for (size_t r = 0; r < c.world.size(); r++)
    if ((int)r != root) {
      size_t got = 0;
      get(comm, c.world[r], TAG_COLL - 7, (char *)recvbuf + (size_t)displs[r] * es, (size_t)recvcounts[r] * es, &got);
    }
  return MPI_SUCCESS;
}
int MPI_Gather(void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount, MPI_Datatype recvtype, int root,
               MPI_Comm comm) {
  Comm &c = C(comm);
  std::vector<int> counts(c.world.size(), recvcount), displs(c.world.size());
  for (size_t r = 0; r < c.world.size(); r++) displs[r] = (int)r * recvcount;
  return MPI_Gatherv(sendbuf, sendcount, sendtype, recvbuf, counts.data(), displs.data(), recvtype, root, comm);
}
int MPI_Reduce_scatter(void *sendbuf, void *recvbuf, int *recvcounts, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  Comm &c = C(comm);
  int total = 0, before = 0;
  for (size_t r = 0; r < c.world.size(); r++) {
    if ((int)r < c.me) before += recvcounts[r];
    total += recvcounts[r];
  }
  std::vector<char> tmp((size_t)total * type_size(type) + 1);
  MPI_Allreduce(sendbuf, tmp.data(), total, type, op, comm);
  memcpy(recvbuf, tmp.data() + (size_t)before * type_size(type), (size_t)recvcounts[c.me] * type_size(type));
  return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *sendbuf, const int *sendcounts, const int *sdispls, MPI_Datatype sendtype, void *recvbuf,
                  const int *recvcounts, const int *rdispls, MPI_Datatype recvtype, MPI_Comm comm) {
  Comm &c = C(comm);
  const size_t ss = type_size(sendtype), rs = type_size(recvtype);
  for (size_t r = 0; r < c.world.size(); r++) {
    if ((int)r == c.me) memcpy((char *)recvbuf + (size_t)rdispls[r] * rs, (const char *)sendbuf + (size_t)sdispls[r] * ss, (size_t)sendcounts[r] * ss);
    else put(comm, c.world[r], TAG_COLL - 8, (const char *)sendbuf + (size_t)sdispls[r] * ss, (size_t)sendcounts[r] * ss);
  }
  for (size_t r = 0; r < c.world.size(); r++)
    if ((int)r != c.me) {
      size_t got = 0;
      get(comm, c.world[r], TAG_COLL - 8, (char *)recvbuf + (size_t)rdispls[r] * rs, (size_t)recvcounts[r] * rs, &got);
    }
  return MPI_SUCCESS;
}

// communicators are created by collective calls, in the same order on every member: the index is the same everywhere
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *out) {
  Comm &c = C(comm);
  int mine[3] = {color, key, g_rank};
  std::vector<int> all(3 * c.world.size());
  MPI_Allgather(mine, 3, MPI_INT, all.data(), 3, MPI_INT, comm);
  std::vector<std::pair<int, int>> members;  // (key, world rank), stable in the old order
  for (size_t r = 0; r < c.world.size(); r++)
    if (all[3 * r] == color) members.push_back(std::make_pair(all[3 * r + 1], all[3 * r + 2]));
  std::stable_sort(members.begin(), members.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
  Comm n;
  for (size_t i = 0; i < members.size(); i++) {
    n.world.push_back(members[i].second);
    if (members[i].second == g_rank) n.me = (int)i;
  }
  comms.push_back(n);
  *out = (int)comms.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) {
  Comm n = C(comm);
  comms.push_back(n);
  *out = (int)comms.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *) { return MPI_SUCCESS; }

int MPI_Cart_create(MPI_Comm comm, int, int *, int *, int, MPI_Comm *out) { return MPI_Comm_dup(comm, out); }
int MPI_Cart_get(MPI_Comm, int, int *, int *, int *) { die("MPI_Cart_get is not emulated"); return 0; }
int MPI_Cart_shift(MPI_Comm, int, int, int *, int *) { die("MPI_Cart_shift is not emulated"); return 0; }
int MPI_Cart_rank(MPI_Comm, int *, int *) { die("MPI_Cart_rank is not emulated"); return 0; }

}  // extern "C"
