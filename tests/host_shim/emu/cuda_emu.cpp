// TEST INFRASTRUCTURE.  Host implementation of cuda_emu.h / cuda_runtime.h / cuda.h (see cuda_emu.h).
#include <cxxabi.h>
#include <dlfcn.h>
#include <pthread.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cuda.h"
#include "cuda_runtime.h"

namespace emu {

thread_local Tls *tlsp = nullptr;
dim3 bdim, gdim;

namespace {
constexpr size_t SMEM_MAX = 232448;     // 227 KB: the opt-in maximum of dynamic shared memory per CTA on sm_100
constexpr size_t FIBER_STACK = 256 << 10;
constexpr long STALL_LIMIT = 4000000;   // yields of waiting fibers without any progress in the CTA: a deadlock

size_t smem_limit = 48 * 1024;  // raised by cudaFuncSetAttribute(MaxDynamicSharedMemorySize)
cudaError_t last_error = cudaSuccess;

// ---- fibers: a context is the stack pointer of a suspended fiber, its callee-saved registers on its stack ----
#if defined(__x86_64__)
extern "C" void spk_emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.hidden spk_emu_switch
.globl spk_emu_switch
.type spk_emu_switch,@function
spk_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size spk_emu_switch,.-spk_emu_switch
)");
struct Context {
  void *sp = nullptr;
  void prepare(void *stack, size_t bytes, void (*entry)()) {
    void **top = reinterpret_cast<void **>(((uintptr_t)stack + bytes) & ~(uintptr_t)15);
    *--top = nullptr;                            // where `entry` would return to (it never does); rsp % 16 == 8 on entry
    *--top = reinterpret_cast<void *>(entry);    // the `ret` of the first switch jumps here
    for (int i = 0; i < 6; i++) *--top = nullptr;  // rbp rbx r12 r13 r14 r15
    sp = top;
  }
};
inline void switch_context(Context &from, Context &to) { spk_emu_switch(&from.sp, to.sp); }
#else  // portable, slower (two signal-mask system calls per switch)
struct Context {
  ucontext_t uc;
  void prepare(void *stack, size_t bytes, void (*entry)()) {
    getcontext(&uc);
    uc.uc_stack.ss_sp = stack;
    uc.uc_stack.ss_size = bytes;
    uc.uc_link = nullptr;
    makecontext(&uc, entry, 0);
  }
};
inline void switch_context(Context &from, Context &to) { swapcontext(&from.uc, &to.uc); }
#endif

struct PendingCopy {
  unsigned char *dst;
  unsigned char desc[128];
  int coord[5];
  int rank;
  uint64_t *bar;
  uint64_t bytes;
};
int order_mode() {
  static const int m = [] {
    const char *e = getenv("SPK_EMU_ORDER");
    return !e ? 0 : !strncmp(e, "reverse", 7) ? 1 : !strncmp(e, "shuffle", 7) ? 2 : 0;
  }();
  return m;
}
uint64_t order_seed() {
  static const uint64_t sd = [] {
    const char *e = getenv("SPK_EMU_ORDER");
    const char *c = e ? strchr(e, ':') : nullptr;
    return c ? strtoull(c + 1, nullptr, 10) : 12345ull;
  }();
  return sd;
}
bool tma_lazy() {
  static const bool on = [] { const char *e = getenv("SPK_EMU_TMA"); return e && !strcmp(e, "lazy"); }();
  return on;
}

struct Fiber {
  Context ctx;
  Tls tls;
  bool done = false;
  void *stack = nullptr;
};
struct WarpState {
  int lanes = 32, alive = 32, count = 0;
  unsigned gen = 0;
  uint64_t slot[32];
};
// one worker = one OS thread = one CTA at a time
struct Worker {
  Context main_ctx;
  std::vector<Fiber> fibers;
  int n = 0, alive = 0, cur = 0;
  int bar_count = 0;
  unsigned bar_gen = 0;
  WarpState warps[32];
  long stalled = 0;
  unsigned char *smem = nullptr;
  const std::function<void()> *body = nullptr;
  std::vector<int> perm;                 // fiber slot -> CUDA thread index
  std::vector<struct PendingCopy> pending;  // SPK_EMU_TMA=lazy: bulk tensor copies issued and not yet performed
};
thread_local Worker *W = nullptr;

void deadlock(const char *what) {
  fprintf(stderr, "cuda_emu: CTA (%u,%u,%u): no thread can make progress (%s): divergent barrier or a wait that is never satisfied\n",
          tlsp->bid.x, tlsp->bid.y, tlsp->bid.z, what);
  abort();
}
void progress() { W->stalled = 0; }

// switch to the next fiber of the CTA that has not finished (round robin)
void yield_fiber(const char *what) {
  Worker *w = W;
  if (++w->stalled > STALL_LIMIT) deadlock(what);
  const int from = w->cur;
  int to = from;
  do { to = to + 1 == w->n ? 0 : to + 1; } while (w->fibers[to].done && to != from);
  if (to == from) return;
  w->cur = to;
  tlsp = &w->fibers[to].tls;
  switch_context(w->fibers[from].ctx, w->fibers[to].ctx);
}

void barrier_release_if_complete(Worker *w) {
  if (w->alive > 0 && w->bar_count == w->alive) {  // threads that have exited count as arrived
    w->bar_count = 0;
    w->bar_gen++;
    progress();
  }
}
void warp_release_if_complete(WarpState &ws) {
  if (ws.alive > 0 && ws.count == ws.alive) {
    ws.count = 0;
    ws.gen++;
    progress();
  }
}

void fiber_main() {
  Worker *w = W;
  (*w->body)();
  // the thread has exited: it no longer takes part in barriers
  w = W;
  Fiber &f = w->fibers[w->cur];
  f.done = true;
  w->alive--;
  progress();
  WarpState &ws = w->warps[f.tls.warp];
  ws.alive--;
  barrier_release_if_complete(w);
  warp_release_if_complete(ws);
  if (w->alive == 0) {
    switch_context(f.ctx, w->main_ctx);
  } else {
    int to = w->cur;
    do { to = to + 1 == w->n ? 0 : to + 1; } while (w->fibers[to].done);
    w->cur = to;
    tlsp = &w->fibers[to].tls;
    switch_context(f.ctx, w->fibers[to].ctx);
  }
  abort();  // a finished fiber is never resumed
}

void run_cta(Worker *w, dim3 block, dim3 bid, const std::function<void()> *body) {
  const int n = (int)(block.x * block.y * block.z);
  if ((int)w->fibers.size() < n) {
    const size_t old = w->fibers.size();
    w->fibers.resize((size_t)n);
    for (size_t i = old; i < (size_t)n; i++) {
      void *st = mmap(nullptr, FIBER_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
      if (st == MAP_FAILED) { fprintf(stderr, "cuda_emu: cannot allocate a fiber stack\n"); abort(); }
      w->fibers[i].stack = st;
    }
  }
  w->n = w->alive = n;
  w->cur = 0;
  w->bar_count = 0;
  w->stalled = 0;
  w->body = body;
  const int nwarps = (n + 31) / 32;
  for (int q = 0; q < nwarps; q++) {
    WarpState &ws = w->warps[q];
    ws.lanes = ws.alive = std::min(32, n - 32 * q);
    ws.count = 0;
  }
  // SPK_EMU_ORDER=reverse|shuffle[:seed]: which CUDA thread runs first inside a CTA (fibers only switch at the
  // blocking points, so the order decides who gets to a shared-memory location first between two barriers: a
  // kernel that misses a barrier computes different results under a different order)
  w->perm.resize((size_t)n);
  for (int i = 0; i < n; i++) w->perm[(size_t)i] = order_mode() == 1 ? n - 1 - i : i;
  if (order_mode() == 2) {
    uint64_t x = order_seed() ^ (0x9E3779B97F4A7C15ull * (uint64_t)(1 + bid.x + 7919u * bid.y + 104729u * bid.z));
    for (int i = n - 1; i > 0; i--) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      std::swap(w->perm[(size_t)i], w->perm[(size_t)(x % (uint64_t)(i + 1))]);
    }
  }
  for (int slot = 0; slot < n; slot++) {
    Fiber &f = w->fibers[(size_t)slot];
    const int i = w->perm[(size_t)slot];
    f.done = false;
    f.tls.lin = i;
    f.tls.lane = i & 31;
    f.tls.warp = i >> 5;
    f.tls.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
    f.tls.bid = bid;
    f.ctx.prepare(f.stack, FIBER_STACK, fiber_main);
  }
  w->pending.clear();
  tlsp = &w->fibers[0].tls;
  switch_context(w->main_ctx, w->fibers[0].ctx);  // returns when the last thread of the CTA has exited
  if (!w->pending.empty()) {
    fprintf(stderr, "cuda_emu: CTA (%u,%u,%u) exited with %zu bulk tensor copies nobody waited for\n", bid.x, bid.y, bid.z, w->pending.size());
    abort();
  }
  tlsp = nullptr;
}

// ---- worker pool -----------------------------------------------------------------------------------
struct Pool {
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::vector<std::thread> threads;
  unsigned long job = 0;  // generation of the current launch
  int running = 0;
  dim3 grid, block;
  const std::function<void()> *body = nullptr;
  std::atomic<long long> next{0};
  long long total = 0;
  bool stop = false;
};
Pool &pool = *new Pool;  // never destroyed: the workers are detached and live until the process ends

void worker_loop() {
  Worker me;
  if (posix_memalign((void **)&me.smem, 1024, SMEM_MAX) != 0) abort();
  W = &me;
  unsigned long seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(pool.mu);
      pool.cv_job.wait(lk, [&] { return pool.stop || pool.job != seen; });
      if (pool.stop) return;
      seen = pool.job;
    }
    memset(me.smem, 0xCB, SMEM_MAX);  // shared memory is not initialised
    for (;;) {
      const long long c = pool.next.fetch_add(1);
      if (c >= pool.total) break;
      const dim3 bid((unsigned)(c % pool.grid.x), (unsigned)((c / pool.grid.x) % pool.grid.y),
                     (unsigned)(c / ((long long)pool.grid.x * pool.grid.y)));
      run_cta(&me, pool.block, bid, pool.body);
    }
    {
      std::lock_guard<std::mutex> lk(pool.mu);
      if (--pool.running == 0) pool.cv_done.notify_all();
    }
  }
}

std::map<std::string, long> launch_log;  // kernel (as written at the launch site) -> launches, under pool.mu
const bool trace = getenv("SPK_EMU_TRACE") != nullptr;

// the instantiation behind a launch: the demangled symbol of the kernel function (template arguments as
// values), or the text of the launch site if the symbol cannot be resolved
std::string kernel_name(const char *label, const void *fn) {
  Dl_info info;
  if (fn && dladdr(fn, &info) && info.dli_sname && info.dli_saddr == fn) {
    int status = 0;
    char *d = abi::__cxa_demangle(info.dli_sname, nullptr, nullptr, &status);
    if (status == 0 && d) {
      std::string s(d);
      free(d);
      const size_t paren = s.find('(');  // drop the parameter list and the return type
      if (paren != std::string::npos) s.resize(paren);
      if (s.rfind("void ", 0) == 0) s.erase(0, 5);
      return s;
    }
  }
  return label;
}

int pool_size() {
  static const int n = [] {
    const char *e = getenv("SPK_EMU_WORKERS");
    int v = e ? atoi(e) : 0;
    if (v <= 0) v = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(v, 64));
  }();
  return n;
}
}  // namespace

void syncthreads() {
  Worker *w = W;
  const unsigned gen = w->bar_gen;
  w->bar_count++;
  barrier_release_if_complete(w);
  while (W->bar_gen == gen) yield_fiber("__syncthreads");
}

void warp_gather(uint64_t mine, uint64_t (&all)[32]) {
  WarpState &ws = W->warps[tlsp->warp];
  const int lane = tlsp->lane;
  // phase 1: everybody deposits; phase 2: everybody has read (nobody deposits the next value earlier)
  ws.slot[lane] = mine;
  unsigned gen = ws.gen;
  ws.count++;
  warp_release_if_complete(ws);
  while (ws.gen == gen) yield_fiber("warp collective");
  for (int l = 0; l < 32; l++) all[l] = l < ws.lanes ? ws.slot[l] : 0;
  gen = ws.gen;
  ws.count++;
  warp_release_if_complete(ws);
  while (ws.gen == gen) yield_fiber("warp collective");
}
uint64_t warp_exchange(uint64_t mine, int src_lane) {
  uint64_t all[32];
  warp_gather(mine, all);
  return all[src_lane];
}

unsigned char *dynamic_smem() { return W->smem; }

void launch(const char *label, const void *fn, dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  const long long n = (long long)block.x * block.y * block.z;
  if (n <= 0 || n > 1024 || block.z > 64 || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535 ||
      smem_bytes > smem_limit) {
    fprintf(stderr, "cuda_emu: %s: invalid launch configuration grid (%u,%u,%u) block (%u,%u,%u) smem %zu (limit %zu)\n", label,
            grid.x, grid.y, grid.z, block.x, block.y, block.z, smem_bytes, smem_limit);
    last_error = cudaErrorInvalidValue;
    return;
  }
  const std::string kernel = kernel_name(label, fn);
  // SPK_EMU_JITTER=usec: a random pause before every launch (scale usec), different in every process: ranks that are
  // processes drift apart, so an ordering between ranks that only held by accident breaks
  static const long jitter_us = [] { const char *e = getenv("SPK_EMU_JITTER"); return e ? atol(e) : 0L; }();
  if (jitter_us > 0) {
    static uint64_t x = 0x9E3779B97F4A7C15ull * (uint64_t)(getpid() + 1);
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    // mostly short, one launch in sixteen long: a rank that stalls now and then, as a busy GPU does
    const uint64_t span = (x >> 40) % 16 == 0 ? 8 * (uint64_t)jitter_us : (uint64_t)jitter_us / 8 + 1;
    std::this_thread::sleep_for(std::chrono::microseconds((long)(x % span)));
  }
  std::unique_lock<std::mutex> lk(pool.mu);
  launch_log[kernel]++;
  if (trace) fprintf(stderr, "cuda_emu: %s grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", kernel.c_str(), grid.x, grid.y, grid.z, block.x,
                     block.y, block.z, smem_bytes);
  if (pool.threads.empty())
    for (int i = 0; i < pool_size(); i++) {
      pool.threads.emplace_back(worker_loop);
      pool.threads.back().detach();
    }
  bdim = block;
  gdim = grid;
  pool.grid = grid;
  pool.block = block;
  pool.body = &body;
  pool.total = (long long)grid.x * grid.y * grid.z;
  pool.next = 0;
  pool.running = (int)pool.threads.size();
  pool.job++;
  pool.cv_job.notify_all();
  pool.cv_done.wait(lk, [&] { return pool.running == 0; });
}

uint64_t globaltimer() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

// ---- mbarrier (arrival count 1) -------------------------------------------------------------------
// word: bits 0..30 number of completed phases, bit 31 "the pending arrival of this phase has
// happened" (mbarrier.arrive.expect_tx), bits 32..63 outstanding transaction bytes (signed).
// The phase completes when the arrival has happened and the transaction count is back to zero.
// The word lives in the CTA's shared memory and only that CTA's fibers (one OS thread) touch it.
namespace {
void mb_update(uint64_t *bar, int64_t tx_delta, bool arrive) {
  uint32_t lo = (uint32_t)*bar;
  int32_t tx = (int32_t)(*bar >> 32);
  tx += (int32_t)tx_delta;
  if (arrive) {
    if (lo & 0x80000000u) {
      fprintf(stderr, "cuda_emu: second arrival on an mbarrier of count 1 within one phase\n");
      abort();
    }
    lo |= 0x80000000u;
  }
  if ((lo & 0x80000000u) && tx == 0) {
    lo = ((lo & 0x7FFFFFFFu) + 1u) & 0x7FFFFFFFu;
    progress();
  }
  *bar = ((uint64_t)(uint32_t)tx << 32) | lo;
}
}  // namespace
void mbar_init(uint64_t *bar, int count) {
  if (count != 1) {
    fprintf(stderr, "cuda_emu: mbarrier arrival count %d not supported\n", count);
    abort();
  }
  *bar = 0;
}
void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { mb_update(bar, (int64_t)bytes, true); }
void tma_flush(uint64_t *bar);
void mbar_wait(uint64_t *bar, uint32_t parity) {
  // SPK_EMU_TMA=lazy: the copies that signal this barrier land now, at the latest moment the kernel allows
  if (tma_lazy()) tma_flush(bar);
  // try_wait.parity(p) succeeds once the phase of parity p has completed: the current phase has the other parity
  while ((((uint32_t)*(volatile uint64_t *)bar) & 1u) == (parity & 1u)) yield_fiber("mbarrier wait");
}

// ---- tensor maps and the bulk tensor copy ---------------------------------------------------------
namespace {
struct TmapDesc {
  uint64_t magic;
  unsigned char *base;
  uint32_t rank, elem;
  uint64_t dims[5];
  uint64_t strides[5];  // bytes; strides[0] = elem
  uint32_t box[5];
};
static_assert(sizeof(TmapDesc) <= sizeof(CUtensorMap), "descriptor must fit the opaque tensor map");
constexpr uint64_t TMAP_MAGIC = 0x544D41505F454D55ull;

// cuTensorMapEncodeTiled with the argument checks of the driver API reference
CUresult encode_tiled(CUtensorMap *map, CUtensorMapDataType dtype, cuuint32_t rank, void *base, const cuuint64_t *dims,
                      const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *estr, CUtensorMapInterleave il,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  static const uint32_t esize[] = {1, 2, 4, 4, 8, 8};
  if (!map || rank < 1 || rank > 5 || (int)dtype < 0 || (int)dtype > 5 || il != CU_TENSOR_MAP_INTERLEAVE_NONE ||
      sw != CU_TENSOR_MAP_SWIZZLE_NONE)
    return CUDA_ERROR_INVALID_VALUE;
  const uint32_t elem = esize[(int)dtype];
  if (!base || ((uintptr_t)base & 15u)) return CUDA_ERROR_INVALID_VALUE;  // globalAddress: 16-byte aligned
  TmapDesc d;
  memset(&d, 0, sizeof d);
  d.magic = TMAP_MAGIC;
  d.base = (unsigned char *)base;
  d.rank = rank;
  d.elem = elem;
  for (uint32_t i = 0; i < rank; i++) {
    if (dims[i] == 0 || dims[i] > (1ull << 32)) return CUDA_ERROR_INVALID_VALUE;  // globalDim: non-zero, <= 2^32
    if (box[i] == 0 || box[i] > 256) return CUDA_ERROR_INVALID_VALUE;              // boxDim: non-zero, <= 256
    if (estr[i] == 0 || estr[i] > 8) return CUDA_ERROR_INVALID_VALUE;
    if (estr[i] != 1) return CUDA_ERROR_INVALID_VALUE;  // not emulated
    d.dims[i] = dims[i];
    d.box[i] = box[i];
    if (i == 0) d.strides[0] = elem;
    else {
      const uint64_t s = strides[i - 1];
      if (s == 0 || (s & 15u) || s >= (1ull << 40)) return CUDA_ERROR_INVALID_VALUE;  // globalStrides: multiple of 16, < 2^40
      d.strides[i] = s;
    }
  }
  if (((uint64_t)box[0] * elem) & 15u) return CUDA_ERROR_INVALID_VALUE;  // inner box extent: multiple of 16 bytes
  memset(map, 0, sizeof *map);
  memcpy(map, &d, sizeof d);
  return CUDA_SUCCESS;
}
}  // namespace

namespace {
void tma_copy(const TmapDesc &d, unsigned char *out, const int *coord, uint64_t total) {
  uint32_t idx[5] = {0, 0, 0, 0, 0};
  for (uint64_t e = 0; e < total / d.elem; e++) {
    bool inside = true;
    uint64_t off = 0;
    for (int i = 0; i < (int)d.rank; i++) {
      const long long c = (long long)coord[i] + idx[i];
      if (c < 0 || c >= (long long)d.dims[i]) inside = false;
      else off += (uint64_t)c * d.strides[i];
    }
    if (inside) memcpy(out + e * d.elem, d.base + off, d.elem);
    else memset(out + e * d.elem, 0, d.elem);  // out-of-bound elements are filled with zeros
    for (int i = 0; i < (int)d.rank; i++) {
      if (++idx[i] < d.box[i]) break;
      idx[i] = 0;
    }
  }
}
}  // namespace

void tma_flush(uint64_t *bar) {
  Worker *w = W;
  for (size_t i = 0; i < w->pending.size();) {
    if (w->pending[i].bar != bar) { i++; continue; }
    const PendingCopy pc = w->pending[i];
    w->pending.erase(w->pending.begin() + (long)i);
    TmapDesc d;
    memcpy(&d, pc.desc, sizeof d);
    tma_copy(d, pc.dst, pc.coord, pc.bytes);
    mb_update(pc.bar, -(int64_t)pc.bytes, false);  // complete_tx
  }
}

void tma_load(void *dst, const void *tensor_map, uint64_t *bar, const int *coord, int rank) {
  TmapDesc d;
  memcpy(&d, tensor_map, sizeof d);
  if (d.magic != TMAP_MAGIC || (int)d.rank != rank) {
    fprintf(stderr, "cuda_emu: bulk tensor copy with an invalid tensor map\n");
    abort();
  }
  unsigned char *out = (unsigned char *)dst;
  uint64_t total = d.elem;
  for (int i = 0; i < rank; i++) total *= d.box[i];
  if (((uintptr_t)out & 127u) || out < W->smem || out + total > W->smem + SMEM_MAX) {  // shared-memory destination: 128-byte aligned
    fprintf(stderr, "cuda_emu: bulk tensor copy destination misaligned or outside shared memory\n");
    abort();
  }
  if (tma_lazy()) {
    // the copy is in flight: its destination holds neither the old nor the new contents for anybody who does not wait
    memset(out, 0xCB, total);
    PendingCopy pc;
    pc.dst = out;
    memset(pc.desc, 0, sizeof pc.desc);
    memcpy(pc.desc, &d, sizeof d);
    for (int i = 0; i < 5; i++) pc.coord[i] = i < rank ? coord[i] : 0;
    pc.rank = rank;
    pc.bar = bar;
    pc.bytes = total;
    W->pending.push_back(pc);
    return;
  }
  tma_copy(d, out, coord, total);
  mb_update(bar, -(int64_t)total, false);  // complete_tx
}

}  // namespace emu

void __nanosleep(unsigned) {  // inside a spin on memory another process writes: not a stall of this CTA
  if (emu::W) emu::W->stalled = 0;
  std::this_thread::yield();
}

// ---- runtime API ------------------------------------------------------------------------------------
struct emuStream_st { int dummy; };
struct emuEvent_st { std::chrono::steady_clock::time_point t; bool recorded = false; };

const char *cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotReady: return "not ready";
    case cudaErrorNotSupported: return "operation not supported (cuda_emu)";
    default: return "unknown error";
  }
}
cudaError_t cudaGetLastError() {
  const cudaError_t e = emu::last_error;
  emu::last_error = cudaSuccess;
  return e;
}
// ---- "device" memory --------------------------------------------------------------------------------
// Default: the heap.  With SPK_EMU_ARENA_DIR set (multi-rank tests: one process per rank) every
// allocation comes out of one file-backed shared mapping per process, so that cudaIpcGetMemHandle /
// cudaIpcOpenMemHandle can hand an allocation to another process the way CUDA IPC does: the peer
// maps the same pages, stores through the mapping are visible to the owner.
namespace {
constexpr size_t ARENA_BYTES = 8ull << 30;  // sparse: only touched pages exist
struct Arena {
  std::mutex mu;
  unsigned char *base = nullptr;
  std::string path;
  std::map<size_t, size_t> used, free_;  // offset -> bytes
  std::map<int, unsigned char *> peers;  // pid -> mapping of that process's arena
  std::string dir;
};
Arena &arena = *new Arena;
struct IpcPayload { uint64_t magic; int32_t pid; uint32_t pad; uint64_t offset, bytes; };
static_assert(sizeof(IpcPayload) <= sizeof(cudaIpcMemHandle_t), "");
constexpr uint64_t IPC_MAGIC = 0x4950435F454D5531ull;

std::string arena_path(const std::string &dir, int pid) { return dir + "/spk_emu_arena_" + std::to_string(pid); }
void arena_unlink() {
  if (!arena.path.empty()) unlink(arena.path.c_str());
}
bool arena_enabled() {
  static const bool on = [] {
    const char *d = getenv("SPK_EMU_ARENA_DIR");
    if (!d || !d[0]) return false;
    arena.dir = d;
    arena.path = arena_path(arena.dir, (int)getpid());
    const int fd = open(arena.path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)ARENA_BYTES) != 0) { perror("cuda_emu: arena"); abort(); }
    void *m = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { perror("cuda_emu: arena mmap"); abort(); }
    arena.base = (unsigned char *)m;
    arena.free_[0] = ARENA_BYTES;
    atexit(arena_unlink);
    return true;
  }();
  return on;
}
void *arena_alloc(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  std::lock_guard<std::mutex> lk(arena.mu);
  for (auto it = arena.free_.begin(); it != arena.free_.end(); ++it)
    if (it->second >= bytes) {
      const size_t off = it->first, left = it->second - bytes;
      arena.free_.erase(it);
      if (left) arena.free_[off + bytes] = left;
      arena.used[off] = bytes;
      return arena.base + off;
    }
  return nullptr;
}
bool arena_free(void *p) {
  std::lock_guard<std::mutex> lk(arena.mu);
  if (!arena.base || (unsigned char *)p < arena.base || (unsigned char *)p >= arena.base + ARENA_BYTES) return false;
  const size_t off = (size_t)((unsigned char *)p - arena.base);
  auto it = arena.used.find(off);
  if (it == arena.used.end()) { fprintf(stderr, "cuda_emu: cudaFree of a pointer that is not an allocation\n"); abort(); }
  size_t bytes = it->second, start = off;
  arena.used.erase(it);
  auto nx = arena.free_.find(start + bytes);  // coalesce with the free neighbours
  if (nx != arena.free_.end()) { bytes += nx->second; arena.free_.erase(nx); }
  auto pv = arena.free_.lower_bound(start);
  if (pv != arena.free_.begin()) {
    --pv;
    if (pv->first + pv->second == start) { start = pv->first; bytes += pv->second; arena.free_.erase(pv); }
  }
  arena.free_[start] = bytes;
  return true;
}
int device_count() {
  static const int n = [] { const char *e = getenv("SPK_EMU_DEVICES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
  return n;
}
}  // namespace

cudaError_t cudaGetDeviceCount(int *n) { *n = device_count(); return cudaSuccess; }
cudaError_t cudaSetDevice(int dev) { return dev >= 0 && dev < device_count() ? cudaSuccess : cudaErrorInvalidValue; }
// SPK_EMU_GUARD=after|before: every allocation gets an inaccessible page right behind its last
// (16-byte rounded) byte, or right before its first one, so that a kernel reading or writing outside an
// allocation faults at the access instead of silently using a neighbour's bytes (a memcheck for the CPU tier)
namespace {
struct Guarded { void *map; size_t map_bytes; };
std::map<void *, Guarded> guarded;
std::mutex guarded_mu;
int guard_mode() {
  static const int m = [] {
    const char *e = getenv("SPK_EMU_GUARD");
    return !e ? 0 : !strcmp(e, "after") ? 1 : !strcmp(e, "before") ? 2 : 0;
  }();
  return m;
}
void *guarded_alloc(size_t bytes) {
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  const size_t need = (bytes + 15) & ~(size_t)15;
  const size_t body = (need + page - 1) / page * page;
  unsigned char *m = (unsigned char *)mmap(nullptr, body + 2 * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) return nullptr;
  mprotect(m, page, PROT_NONE);
  mprotect(m + page + body, page, PROT_NONE);
  void *p = guard_mode() == 1 ? m + page + body - need : m + page;
  std::lock_guard<std::mutex> lk(guarded_mu);
  guarded[p] = Guarded{m, body + 2 * page};
  return p;
}
bool guarded_free(void *p) {
  std::lock_guard<std::mutex> lk(guarded_mu);
  auto it = guarded.find(p);
  if (it == guarded.end()) return false;
  munmap(it->second.map, it->second.map_bytes);
  guarded.erase(it);
  return true;
}
}  // namespace

cudaError_t emuMalloc(void **p, size_t bytes) {
  void *q = nullptr;
  if (arena_enabled()) q = arena_alloc(bytes ? bytes : 1);
  else if (guard_mode()) q = guarded_alloc(bytes ? bytes : 1);
  else if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) q = nullptr;
  if (!q) { *p = nullptr; return cudaErrorMemoryAllocation; }
  memset(q, 0xCD, bytes);  // device memory is not initialised
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
  if (!p) return cudaSuccess;
  if (arena_enabled() && arena_free(p)) return cudaSuccess;
  if (guard_mode()) {
    if (!guarded_free(p)) { fprintf(stderr, "cuda_emu: cudaFree of a pointer that is not an allocation\n"); abort(); }
    return cudaSuccess;
  }
  free(p);
  return cudaSuccess;
}
cudaError_t emuMallocHost(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
  memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int value, size_t bytes, cudaStream_t) { memset(p, value, bytes); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int value, size_t bytes) { memset(p, value, bytes); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emuStream_st; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent_st; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emuEvent_st; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->t = std::chrono::steady_clock::now();
  e->recorded = true;
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  if (!a->recorded || !b->recorded) return cudaErrorInvalidValue;
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  if (!arena_enabled()) return cudaErrorNotSupported;
  std::lock_guard<std::mutex> lk(arena.mu);
  const size_t off = (size_t)((unsigned char *)p - arena.base);
  auto it = arena.used.find(off);
  if ((unsigned char *)p < arena.base || it == arena.used.end()) return cudaErrorInvalidValue;  // not the base of an allocation
  IpcPayload pl = {IPC_MAGIC, (int32_t)getpid(), 0, off, it->second};
  memset(h, 0, sizeof *h);
  memcpy(h, &pl, sizeof pl);
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
  if (!arena_enabled()) return cudaErrorNotSupported;
  IpcPayload pl;
  memcpy(&pl, &h, sizeof pl);
  if (pl.magic != IPC_MAGIC) return cudaErrorInvalidValue;
  if (pl.pid == (int32_t)getpid()) return cudaErrorInvalidValue;  // CUDA IPC cannot open a handle in the process that made it
  std::lock_guard<std::mutex> lk(arena.mu);
  auto it = arena.peers.find(pl.pid);
  if (it == arena.peers.end()) {
    const int fd = open(arena_path(arena.dir, pl.pid).c_str(), O_RDWR);
    if (fd < 0) return cudaErrorInvalidValue;
    void *m = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
    it = arena.peers.emplace(pl.pid, (unsigned char *)m).first;
  }
  *p = it->second + pl.offset;
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) {
  if (!arena_enabled()) return cudaErrorNotSupported;
  std::lock_guard<std::mutex> lk(arena.mu);
  for (const auto &kv : arena.peers)
    if ((unsigned char *)p >= kv.second && (unsigned char *)p < kv.second + ARENA_BYTES) return cudaSuccess;  // mappings stay until exit
  return cudaErrorInvalidValue;
}
cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *status) {
  if (!strcmp(symbol, "cuTensorMapEncodeTiled")) {
    *fn = (void *)&emu::encode_tiled;
    if (status) *status = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  if (status) *status = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorInvalidValue;
}
cudaError_t emuFuncSetAttribute(int attr, int value) {
  if (attr != (int)cudaFuncAttributeMaxDynamicSharedMemorySize || value < 0 || (size_t)value > emu::SMEM_MAX)
    return cudaErrorInvalidValue;
  // the limit is per kernel on the GPU; the emulator keeps the largest request (every launch passes its own size)
  emu::smem_limit = std::max(emu::smem_limit, (size_t)value);
  return cudaSuccess;
}

// ---- what the tests ask the emulator itself -----------------------------------------------------------
extern "C" {
// marks the library as the emulated build (spparks_b200/_lib.py refuses it unless a test opts in)
int spk_cuda_emulated(void) { return 1; }
// "kernel<args> count" lines of every kernel launched since the last reset; returns the length needed
size_t spk_emu_launch_log(char *buf, size_t cap) {
  std::string out;
  {
    std::lock_guard<std::mutex> lk(emu::pool.mu);
    for (const auto &kv : emu::launch_log) out += kv.first + "\t" + std::to_string(kv.second) + "\n";
  }
  if (buf && cap) {
    const size_t n = std::min(cap - 1, out.size());
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}
// checks of the emulator itself (0 = all passed, otherwise the number of the failed check)
int spk_emu_selftest(void) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, nullptr) != cudaSuccess) return 1;
  encode_fn enc = (encode_fn)fn;
  int32_t *buf = nullptr;
  if (cudaMalloc(&buf, 64 * 32 * 4) != cudaSuccess) return 2;
  for (int i = 0; i < 64 * 32; i++) buf[i] = i + 1;
  CUtensorMap m;
  cuuint64_t dims[2] = {64, 32}, strides[1] = {64 * 4};
  cuuint32_t box[2] = {16, 4}, estr[2] = {1, 1};
  auto encode = [&](const cuuint32_t *b, const cuuint64_t *st, void *base) {
    return enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, base, dims, st, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  const cuuint32_t wide[2] = {272, 4}, ragged[2] = {6, 4};
  const cuuint64_t odd_stride[1] = {64 * 4 + 4};
  if (encode(wide, strides, buf) == CUDA_SUCCESS) return 3;        // box extent above 256
  if (encode(ragged, strides, buf) == CUDA_SUCCESS) return 4;      // inner box extent not a multiple of 16 bytes
  if (encode(box, odd_stride, buf) == CUDA_SUCCESS) return 5;      // stride not a multiple of 16 bytes
  if (encode(box, strides, buf + 1) == CUDA_SUCCESS) return 6;     // base not 16-byte aligned
  if (encode(box, strides, buf) != CUDA_SUCCESS) return 7;
  // a box that hangs over the low corner: zeros outside, the tensor's elements inside; one thread
  // issues, everybody waits on the mbarrier, the barrier is reused with the other parity
  int bad = 0;
  emu::launch("selftest", nullptr, dim3(3), dim3(64), 16 * 4 * 4 + 64, [&]() {
    unsigned char *smem = emu::dynamic_smem();
    int32_t *tile = reinterpret_cast<int32_t *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16 * 4 * 4);
    if (threadIdx.x == 0) emu::mbar_init(bar, 1);
    __syncthreads();
    for (int round = 0; round < 2; round++) {
      const int cx = round ? 56 : -4, cy = round ? 30 : -1;
      if (threadIdx.x == 0) {
        emu::mbar_expect_tx(bar, 16 * 4 * 4);
        const int c[2] = {cx, cy};
        emu::tma_load(tile, &m, bar, c, 2);
      }
      emu::mbar_wait(bar, (uint32_t)round);
      const int t = (int)threadIdx.x, x = cx + t % 16, y = cy + t / 16;
      const int want = (x < 0 || x >= 64 || y < 0 || y >= 32) ? 0 : y * 64 + x + 1;
      if (tile[t] != want) __atomic_fetch_add(&bad, 1, __ATOMIC_SEQ_CST);
      __syncthreads();
    }
    // warp collectives
    const int lane = (int)threadIdx.x & 31;
    if (__reduce_add_sync(0xffffffffu, lane) != 496) __atomic_fetch_add(&bad, 1, __ATOMIC_SEQ_CST);
    if (__shfl_down_sync(0xffffffffu, lane, 1) != (lane < 31 ? lane + 1 : lane)) __atomic_fetch_add(&bad, 1, __ATOMIC_SEQ_CST);
    if (__shfl_xor_sync(0xffffffffu, lane, 5) != (lane ^ 5)) __atomic_fetch_add(&bad, 1, __ATOMIC_SEQ_CST);
    if (__match_any_sync(0xffffffffu, lane & 3) != (0x11111111u << (lane & 3))) __atomic_fetch_add(&bad, 1, __ATOMIC_SEQ_CST);
    if (__ballot_sync(0xffffffffu, lane & 1) != 0xAAAAAAAAu) __atomic_fetch_add(&bad, 1, __ATOMIC_SEQ_CST);
  });
  cudaFree(buf);
  if (bad) return 8;
  if (__viaddmin_u16x2(0x0005FFFFu, 0xFFFB0001u, 0x00010001u) != 0x00000000u) return 9;   // 5 + (-5) = 0, 0xFFFF + 1 = 0 (mod 2^16)
  if (__viaddmin_u16x2(0x00070003u, 0xFFFBFFFEu, 0x00010001u) != 0x00010001u) return 10;  // 2 and 1, capped at 1
  if (__funnelshift_l(0x80000000u, 0x00000001u, 1) != 0x00000003u || __funnelshift_r(0x00000003u, 0x00000001u, 1) != 0x80000001u)
    return 11;
  return 0;
}
// which CUDA thread of a 64-thread CTA runs first (SPK_EMU_ORDER), and whether a bulk tensor copy is visible before
// its barrier has been waited for (SPK_EMU_TMA=lazy: no): first * 2 + visible
int spk_emu_probe_modes(void) {
  int first = -1, visible = -1;
  int32_t *buf = nullptr;
  cudaMalloc(&buf, 64 * 4);
  for (int i = 0; i < 64; i++) buf[i] = 1000 + i;
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, nullptr);
  CUtensorMap m;
  cuuint64_t dims[1] = {64};
  cuuint32_t box[1] = {64}, estr[1] = {1};
  ((encode_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 1, buf, dims, nullptr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  emu::launch("probe", nullptr, dim3(1), dim3(64), 64 * 4 + 64, [&]() {
    if (first < 0) first = (int)threadIdx.x;
    unsigned char *smem = emu::dynamic_smem();
    int32_t *tile = reinterpret_cast<int32_t *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 64 * 4);
    if (threadIdx.x == 0) {
      emu::mbar_init(bar, 1);
      emu::mbar_expect_tx(bar, 64 * 4);
      const int c[1] = {0};
      emu::tma_load(tile, &m, bar, c, 1);
      visible = tile[5] == 1005;
    }
    __syncthreads();
    emu::mbar_wait(bar, 0);
    if (tile[threadIdx.x] != 1000 + (int)threadIdx.x) visible = -100;
  });
  cudaFree(buf);
  return first * 2 + visible;
}
// reads one byte at `offset` of a fresh 100-byte allocation (the test of SPK_EMU_GUARD: outside -> fault)
int spk_emu_guard_probe(long offset) {
  unsigned char *p = nullptr;
  if (cudaMalloc(&p, 100) != cudaSuccess) return -1;
  const int v = *(volatile unsigned char *)(p + offset);
  cudaFree(p);
  return v;
}
void spk_emu_launch_log_reset(void) {
  std::lock_guard<std::mutex> lk(emu::pool.mu);
  emu::launch_log.clear();
}
}
