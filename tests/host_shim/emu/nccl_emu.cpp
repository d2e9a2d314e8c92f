// TEST INFRASTRUCTURE.  The ten NCCL entry points libspparks_cuda binds with dlopen (spk_cuda.cu:
// NcclApi), implemented over a shared-memory file for ranks that are processes of one host.  Used
// through SPK_CUDA_NCCL_LIB by the multi-rank tests of the emulated library
// (tests/test_emulated_library.py); the emulated streams are synchronous, so every call completes
// before it returns.  Collectives meet at an inter-process barrier; point-to-point operations of
// a group are matched per pair of ranks in the order they were issued, as NCCL does.
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <vector>

#include "nccl.h"

typedef struct emuStream_st *cudaStream_t;

namespace {
constexpr size_t SLOT_BYTES = 64ull << 20;  // per rank: payload of one collective or of one group's sends (sparse file)
constexpr int MAX_P2P = 16;
constexpr int WAIT_LIMIT_S = 120;

struct Header {
  std::atomic<uint32_t> arrived, generation;
};
struct P2pEntry { int32_t peer; uint32_t pad; uint64_t offset, bytes; };
struct SlotHeader {
  uint32_t nsend;
  uint32_t pad;
  P2pEntry send[MAX_P2P];
};
constexpr size_t SLOT_DATA = 4096;  // payload starts here inside a slot

size_t type_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    default: return 8;
  }
}
}  // namespace

struct ncclComm {
  int rank, nranks;
  unsigned char *map;
  size_t map_bytes;
  bool grouping = false;
  struct Pending { bool send; void *buf; size_t bytes; int peer; };
  std::vector<Pending> ops;
  Header *hdr() { return reinterpret_cast<Header *>(map); }
  unsigned char *slot(int r) { return map + 4096 + (size_t)r * SLOT_BYTES; }
  void barrier() {
    Header *h = hdr();
    const uint32_t gen = h->generation.load();
    if (h->arrived.fetch_add(1) + 1 == (uint32_t)nranks) {
      h->arrived.store(0);
      h->generation.fetch_add(1);
      return;
    }
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (long spins = 0; h->generation.load() == gen; spins++) {
      sched_yield();
      if ((spins & 1023) == 1023) {
        timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        if (t.tv_sec - t0.tv_sec > WAIT_LIMIT_S) {
          fprintf(stderr, "nccl_emu: rank %d waited %d s at a collective: the ranks do not call the same sequence\n", rank, WAIT_LIMIT_S);
          abort();
        }
      }
    }
  }
};

namespace {
thread_local ncclComm *group_comm = nullptr;
thread_local bool in_group = false;

ncclResult_t run_p2p(ncclComm *c) {
  // my sends into my slot, in issue order
  SlotHeader *mine = reinterpret_cast<SlotHeader *>(c->slot(c->rank));
  size_t off = SLOT_DATA;
  mine->nsend = 0;
  for (const auto &op : c->ops)
    if (op.send) {
      if (mine->nsend == MAX_P2P || off + op.bytes > SLOT_BYTES) return ncclInternalError;
      memcpy(c->slot(c->rank) + off, op.buf, op.bytes);
      mine->send[mine->nsend++] = P2pEntry{op.peer, 0, off, op.bytes};
      off += (op.bytes + 63) & ~(size_t)63;
    }
  c->barrier();
  // my receives: the k-th receive from peer p takes p's k-th send addressed to me
  std::vector<int> taken((size_t)c->nranks, 0);
  ncclResult_t rc = ncclSuccess;
  for (const auto &op : c->ops)
    if (!op.send) {
      const SlotHeader *theirs = reinterpret_cast<const SlotHeader *>(c->slot(op.peer));
      int seen = 0;
      bool found = false;
      for (uint32_t k = 0; k < theirs->nsend && !found; k++)
        if (theirs->send[k].peer == c->rank && seen++ == taken[(size_t)op.peer]) {
          if (theirs->send[k].bytes != op.bytes) rc = ncclInternalError;
          else memcpy(op.buf, c->slot(op.peer) + theirs->send[k].offset, op.bytes);
          found = true;
        }
      if (!found) rc = ncclInternalError;
      taken[(size_t)op.peer]++;
    }
  c->barrier();  // nobody refills a slot before everybody has read
  c->ops.clear();
  return rc;
}
}  // namespace

template <class T> static void reduce_into(T *out, const unsigned char *in, size_t count, ncclRedOp_t op, bool first) {
  for (size_t i = 0; i < count; i++) {
    T v;
    memcpy(&v, in + i * sizeof(T), sizeof(T));
    if (first) out[i] = v;
    else if (op == ncclSum) out[i] = out[i] + v;
    else if (op == ncclMax) out[i] = v > out[i] ? v : out[i];
    else if (op == ncclMin) out[i] = v < out[i] ? v : out[i];
    else out[i] = out[i] * v;
  }
}
extern "C" {

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "nccl_emu: error (unmatched or oversized operation)"; }

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof *id);
  uint64_t token = 0;
  const int fd = open("/dev/urandom", O_RDONLY);
  if (fd < 0 || read(fd, &token, sizeof token) != (ssize_t)sizeof token) return ncclSystemError;
  close(fd);
  const char *dir = getenv("SPK_EMU_ARENA_DIR");
  snprintf(id->internal, sizeof id->internal, "%s/spk_emu_nccl_%016llx", dir && dir[0] ? dir : "/tmp", (unsigned long long)token);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return ncclInternalError;
  id.internal[sizeof id.internal - 1] = 0;
  const size_t bytes = 4096 + (size_t)nranks * SLOT_BYTES;
  const int fd = open(id.internal, O_RDWR | O_CREAT, 0600);
  if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) return ncclSystemError;  // a new file reads as zeros: barrier state starts at 0
  void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return ncclSystemError;
  ncclComm *c = new ncclComm;
  c->rank = rank;
  c->nranks = nranks;
  c->map = (unsigned char *)m;
  c->map_bytes = bytes;
  c->barrier();  // everybody has the file mapped
  if (rank == 0) unlink(id.internal);
  *comm = c;
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) {
  if (!c) return ncclSuccess;
  munmap(c->map, c->map_bytes);
  delete c;
  return ncclSuccess;
}

ncclResult_t ncclGroupStart() {
  in_group = true;
  group_comm = nullptr;
  return ncclSuccess;
}
ncclResult_t ncclGroupEnd() {
  in_group = false;
  ncclComm *c = group_comm;
  group_comm = nullptr;
  return c ? run_p2p(c) : ncclSuccess;
}
static ncclResult_t p2p(bool send, void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c) {
  if (peer < 0 || peer >= c->nranks) return ncclInternalError;
  c->ops.push_back(ncclComm::Pending{send, buf, count * type_size(t), peer});
  if (in_group) {
    group_comm = c;
    return ncclSuccess;
  }
  return run_p2p(c);
}
ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  return p2p(true, const_cast<void *>(buf), count, t, peer, c);
}
ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  return p2p(false, buf, count, t, peer, c);
}

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t) {
  const size_t bytes = count * type_size(t);
  if (SLOT_DATA + bytes > SLOT_BYTES) return ncclInternalError;
  memcpy(c->slot(c->rank) + SLOT_DATA, send, bytes);
  c->barrier();
  for (int r = 0; r < c->nranks; r++) memcpy((unsigned char *)recv + (size_t)r * bytes, c->slot(r) + SLOT_DATA, bytes);
  c->barrier();
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c, cudaStream_t) {
  const size_t bytes = count * type_size(t);
  if (SLOT_DATA + bytes > SLOT_BYTES || op == ncclAvg) return ncclInternalError;
  memcpy(c->slot(c->rank) + SLOT_DATA, send, bytes);
  c->barrier();
  std::vector<unsigned char> acc(bytes ? bytes : 1);
  for (int r = 0; r < c->nranks; r++) {
    const unsigned char *in = c->slot(r) + SLOT_DATA;
    switch (t) {
      case ncclInt8: reduce_into(reinterpret_cast<int8_t *>(acc.data()), in, count, op, r == 0); break;
      case ncclUint8: reduce_into(reinterpret_cast<uint8_t *>(acc.data()), in, count, op, r == 0); break;
      case ncclInt32: reduce_into(reinterpret_cast<int32_t *>(acc.data()), in, count, op, r == 0); break;
      case ncclUint32: reduce_into(reinterpret_cast<uint32_t *>(acc.data()), in, count, op, r == 0); break;
      case ncclInt64: reduce_into(reinterpret_cast<int64_t *>(acc.data()), in, count, op, r == 0); break;
      case ncclUint64: reduce_into(reinterpret_cast<uint64_t *>(acc.data()), in, count, op, r == 0); break;
      case ncclFloat32: reduce_into(reinterpret_cast<float *>(acc.data()), in, count, op, r == 0); break;
      case ncclFloat64: reduce_into(reinterpret_cast<double *>(acc.data()), in, count, op, r == 0); break;
      default: return ncclInternalError;
    }
  }
  c->barrier();
  memcpy(recv, acc.data(), bytes);
  return ncclSuccess;
}

}  // extern "C"
