// cusim_rt.cpp -- runtime of the CPU SIMT emulator (TEST INFRASTRUCTURE ONLY, see cusim_device.h).
//
// Execution model: one thread block at a time per OS thread; every CUDA thread of the block is a ucontext fibre that runs
// until it reaches a scheduling point (__syncthreads, a warp collective, or its end).  When no fibre is runnable the
// scheduler completes the collectives whose participants have all arrived: a warp collective needs every lane named in its
// mask at the same collective, a block barrier needs every thread that has not exited.  Anything else (a lane named in a
// mask has exited or sits at a different barrier, a barrier that part of the block never reaches) is what hangs or
// corrupts a real GPU -- here it fails the launch with a message.  Large grids are spread over a few OS threads.
#include "cuda_runtime.h"
#include "nvrtc.h"

#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <ucontext.h>
#include <unistd.h>

#include <atomic>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct cusimGraph {
  struct Op { dim3 grid, block; std::function<void()> body; };
  std::vector<Op> ops;
};
// Streams are DEFERRED: work submitted to a non-null stream is queued and drained as late as legality allows -- when the
// stream (or an event recorded behind the work) is synchronised or waited on.  A missing event dependency between two
// streams therefore shows up as a wrong result (the consumer runs before the producer's queue was drained) instead of
// being hidden by lucky timing.  The null stream executes immediately.
struct cusimStream {
  cusimGraph* cap = nullptr;
  std::deque<std::function<void()>> q;
  unsigned long long enq = 0, done = 0;   // operations submitted / executed
};
struct cusimEvent { long long ns = 0; cusimStream* st = nullptr; unsigned long long seq = 0; };
struct cusimLibrary { void* dl = nullptr; };

namespace cusim {

thread_local Ctx ctx;

namespace {
enum { ST_READY = 0, ST_WAIT_BLOCK, ST_WAIT_WARP, ST_DONE };
constexpr size_t STACK_BYTES = 128 * 1024;

struct Fiber {
  ucontext_t uc;
  int state = ST_DONE;
  int op = 0, arg = 0;
  unsigned mask = 0;
  unsigned long long in = 0, out = 0;
};
struct Pool {
  std::vector<Fiber> f;
  char* stacks = nullptr;
  size_t nstacks = 0;
  ucontext_t main;
  int cur = -1;
  const std::function<void()>* body = nullptr;
  bool in_kernel = false;
};
thread_local Pool pool;

std::mutex g_err_mutex;
std::string g_err;                 // first launch failure since the last cudaGetLastError
std::atomic<int> g_failed{0};

void set_error(const std::string& m) {
  std::lock_guard<std::mutex> lk(g_err_mutex);
  if (g_err.empty()) g_err = m;
  g_failed.store(1);
  fprintf(stderr, "[cusim] %s\n", m.c_str());
}

void fiber_entry() {
  Pool& P = pool;
  (*P.body)();
  P.f[(size_t)P.cur].state = ST_DONE;   // returning resumes uc_link = the scheduler
}

void yield_to_scheduler() {
  Pool& P = pool;
  swapcontext(&P.f[(size_t)P.cur].uc, &P.main);
}

// one thread block of n threads; false = the block could not complete
bool run_block(int n) {
  Pool& P = pool;
  if (P.nstacks < (size_t)n) {
    if (P.stacks) munmap(P.stacks, P.nstacks * STACK_BYTES);
    P.stacks = (char*)mmap(nullptr, (size_t)n * STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (P.stacks == (char*)MAP_FAILED) { P.stacks = nullptr; P.nstacks = 0; set_error("cannot allocate fibre stacks"); return false; }
    P.nstacks = (size_t)n;
    P.f.resize((size_t)n);
  }
  for (int t = 0; t < n; ++t) {
    Fiber& F = P.f[(size_t)t];
    getcontext(&F.uc);
    F.uc.uc_stack.ss_sp = P.stacks + (size_t)t * STACK_BYTES;
    F.uc.uc_stack.ss_size = STACK_BYTES;
    F.uc.uc_link = &P.main;
    makecontext(&F.uc, fiber_entry, 0);
    F.state = ST_READY;
  }
  P.in_kernel = true;
  char msg[256];
  for (;;) {
    // Between two scheduling points the threads of a block run one after the other; the ORDER is a free choice on real
    // hardware, so results must not depend on it.  CUSIM_ORDER=reverse|shuffle runs them in another order: a missing
    // __syncthreads / __syncwarp between a shared-memory (or global) write and a read by another thread then changes
    // the result (the emulator's racecheck).
    static const int order_mode = [] { const char* s = getenv("CUSIM_ORDER"); return !s ? 0 : (!strcmp(s, "reverse") ? 1 : (!strcmp(s, "shuffle") ? 2 : 0)); }();
    static thread_local unsigned long long lcg = 0x9E3779B97F4A7C15ull;
    const int rot = order_mode == 2 ? (int)((lcg = lcg * 6364136223846793005ull + 1442695040888963407ull) >> 33) % n : 0;
    for (int i = 0; i < n; ++i) {
      const int t = order_mode == 1 ? n - 1 - i : (order_mode == 2 ? (int)(((long long)i * 61 + rot) % n) : i);   // 61 is coprime to every block size used
      if (P.f[(size_t)t].state != ST_READY) continue;
      P.cur = t;
      ctx.tid = dim3((unsigned)t);
      swapcontext(&P.main, &P.f[(size_t)t].uc);
    }
    int done = 0, wb = 0;
    for (int t = 0; t < n; ++t) { done += P.f[(size_t)t].state == ST_DONE; wb += P.f[(size_t)t].state == ST_WAIT_BLOCK; }
    if (done == n) break;
    bool progressed = false;
    for (int w0 = 0; w0 < n; w0 += 32) {
      const int nl = std::min(32, n - w0);
      int first = -1;
      for (int l = 0; l < nl; ++l)
        if (P.f[(size_t)(w0 + l)].state == ST_WAIT_WARP) { first = l; break; }
      if (first < 0) continue;
      const Fiber& F0 = P.f[(size_t)(w0 + first)];
      bool ok = true;
      for (int l = 0; l < 32 && ok; ++l) {
        if (!((F0.mask >> l) & 1u)) continue;
        if (l >= nl) { snprintf(msg, sizeof msg, "warp collective names lane %d of a %d-lane warp (block %u)", l, nl, ctx.bid.x); ok = false; break; }
        const Fiber& F = P.f[(size_t)(w0 + l)];
        if (F.state == ST_DONE) { snprintf(msg, sizeof msg, "warp collective (op %d): lane %d of warp %d is named in the mask but has exited (block %u)", F0.op, l, w0 / 32, ctx.bid.x); ok = false; }
        else if (F.state == ST_WAIT_BLOCK) { snprintf(msg, sizeof msg, "warp collective (op %d): lane %d of warp %d waits at __syncthreads instead (block %u)", F0.op, l, w0 / 32, ctx.bid.x); ok = false; }
        else if (F.op != F0.op || F.mask != F0.mask) { snprintf(msg, sizeof msg, "divergent warp collective in warp %d: lane %d op %d mask %08x vs lane %d op %d mask %08x (block %u)", w0 / 32, first, F0.op, F0.mask, l, F.op, F.mask, ctx.bid.x); ok = false; }
      }
      if (!ok) { set_error(msg); P.in_kernel = false; return false; }
      for (int l = 0; l < nl; ++l) {
        Fiber& F = P.f[(size_t)(w0 + l)];
        if (F.state == ST_WAIT_WARP && !((F0.mask >> l) & 1u)) { snprintf(msg, sizeof msg, "lane %d of warp %d executes a warp collective whose mask %08x does not name it (block %u)", l, w0 / 32, F0.mask, ctx.bid.x); set_error(msg); P.in_kernel = false; return false; }
      }
      // every named lane has arrived: complete the collective
      unsigned ballot = 0;
      for (int l = 0; l < nl; ++l)
        if (((F0.mask >> l) & 1u) && P.f[(size_t)(w0 + l)].in) ballot |= 1u << l;
      for (int l = 0; l < nl; ++l) {
        if (!((F0.mask >> l) & 1u)) continue;
        Fiber& F = P.f[(size_t)(w0 + l)];
        switch (F.op) {
          case OP_BALLOT: F.out = ballot; break;
          case OP_SHFL_DOWN: {
            const int srcl = l + F.arg;
            F.out = (F.arg >= 0 && srcl < 32 && srcl < nl && ((F0.mask >> srcl) & 1u)) ? P.f[(size_t)(w0 + srcl)].in : F.in;
          } break;
          default: F.out = 0;
        }
      }
      for (int l = 0; l < nl; ++l)
        if ((F0.mask >> l) & 1u) P.f[(size_t)(w0 + l)].state = ST_READY;
      progressed = true;
    }
    if (progressed) continue;
    if (wb > 0 && wb + done == n) {
      for (int t = 0; t < n; ++t)
        if (P.f[(size_t)t].state == ST_WAIT_BLOCK) P.f[(size_t)t].state = ST_READY;
      continue;
    }
    snprintf(msg, sizeof msg, "deadlock in block %u: %d threads done, %d at __syncthreads, %d elsewhere", ctx.bid.x, done, wb, n - done - wb);
    set_error(msg);
    P.in_kernel = false;
    return false;
  }
  P.in_kernel = false;
  return true;
}

void run_grid(dim3 grid, dim3 block, const std::function<void()>& body) {
  if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) { set_error("cusim supports 1-D grids and blocks only"); return; }
  if (block.x == 0 || block.x > 1024 || grid.x == 0) { set_error("invalid launch configuration"); return; }
  static const int nworkers = [] { const char* s = getenv("CUSIM_THREADS"); int v = s ? atoi(s) : 4; return std::max(1, std::min(v, 64)); }();
  auto work = [&](std::atomic<unsigned>* next) {
    ctx.gdim = grid; ctx.bdim = block;
    pool.body = &body;
    for (;;) {
      const unsigned b = next->fetch_add(1);
      if (b >= grid.x) break;
      ctx.bid = dim3(b);
      if (!run_block((int)block.x)) break;
    }
  };
  std::atomic<unsigned> next{0};
  if (nworkers == 1 || grid.x < 64) { work(&next); return; }
  std::vector<std::thread> th;
  for (int k = 1; k < nworkers; ++k) th.emplace_back(work, &next);
  work(&next);
  for (auto& t : th) t.join();
}

}  // namespace

void barrier_block() {
  Pool& P = pool;
  if (!P.in_kernel) { set_error("__syncthreads outside a kernel"); return; }
  P.f[(size_t)P.cur].state = ST_WAIT_BLOCK;
  yield_to_scheduler();
}

unsigned long long warp_collective(int op, unsigned mask, unsigned long long in, int arg) {
  Pool& P = pool;
  if (!P.in_kernel) { set_error("warp collective outside a kernel"); return 0; }
  Fiber& F = P.f[(size_t)P.cur];
  F.op = op; F.mask = mask; F.in = in; F.arg = arg; F.state = ST_WAIT_WARP;
  yield_to_scheduler();
  return P.f[(size_t)P.cur].out;
}

long long clock_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
}

namespace {
thread_local std::vector<cusimStream*> g_streams;   // streams created by this host thread

void drain_until(cusimStream* st, unsigned long long seq) {
  while (st && st->done < seq && !st->q.empty()) {
    std::function<void()> op = std::move(st->q.front());
    st->q.pop_front();
    st->done++;
    op();
  }
}
void drain(cusimStream* st) { if (st) drain_until(st, st->enq); }
void drain_all() {
  for (size_t i = 0; i < g_streams.size(); ++i) drain(g_streams[i]);
}
void submit(cusimStream* st, std::function<void()> op) {
  if (!st) { op(); return; }
  st->q.push_back(std::move(op));
  st->enq++;
}
}  // namespace

void launch(dim3 grid, dim3 block, cudaStream_t st, std::function<void()> body) {
  if (st && st->cap) { st->cap->ops.push_back(cusimGraph::Op{grid, block, std::move(body)}); return; }
  submit(st, [grid, block, body]() { run_grid(grid, block, body); });
}

}  // namespace cusim

// ---------------------------------------------------------------------------------------------------------------------
// runtime API
// ---------------------------------------------------------------------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e) {
  static thread_local std::string s;
  if (e == cudaSuccess) return "no error";
  std::lock_guard<std::mutex> lk(cusim::g_err_mutex);
  s = "cusim error " + std::to_string(e) + (cusim::g_err.empty() ? "" : ": " + cusim::g_err);
  return s.c_str();
}
cudaError_t cudaGetLastError() {
  if (cusim::g_failed.exchange(0)) return cudaErrorLaunchFailure;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { cusim::drain_all(); return cusim::g_failed.load() ? cudaErrorLaunchFailure : cudaSuccess; }
// Device allocations end at an inaccessible guard page (and start right after one), so a kernel that reads or writes past
// the end of a table -- or before its start -- faults at the offending access instead of silently touching a neighbour
// allocation (the emulator's memcheck).  The start keeps the 16-byte alignment vector loads need.
namespace {
struct GuardedBlock { void* base; size_t total; };
std::mutex g_alloc_mutex;
std::vector<std::pair<void*, GuardedBlock>> g_allocs;
const size_t kPage = (size_t)sysconf(_SC_PAGESIZE);
}  // namespace
cudaError_t cudaMalloc(void** p, size_t bytes) {
  if (bytes == 0) bytes = 8;
  const size_t payload = (bytes + 15) / 16 * 16;
  const size_t body = (payload + kPage - 1) / kPage * kPage;
  const size_t total = body + 2 * kPage;
  char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == (char*)MAP_FAILED) return cudaErrorMemoryAllocation;
  mprotect(base, kPage, PROT_NONE);
  mprotect(base + kPage + body, kPage, PROT_NONE);
  char* q = base + kPage + body - payload;     // the payload ends exactly at the trailing guard page
  memset(q, 0xCD, payload);                    // device memory is not zero-initialised: make reads of unwritten memory visible
  {
    std::lock_guard<std::mutex> lk(g_alloc_mutex);
    g_allocs.push_back({q, GuardedBlock{base, total}});
  }
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {      // cudaFree synchronises the device
  if (!p) return cudaSuccess;
  cusim::drain_all();
  std::lock_guard<std::mutex> lk(g_alloc_mutex);
  for (size_t i = 0; i < g_allocs.size(); ++i)
    if (g_allocs[i].first == p) {
      munmap(g_allocs[i].second.base, g_allocs[i].second.total);
      g_allocs[i] = g_allocs.back();
      g_allocs.pop_back();
      return cudaSuccess;
    }
  cusim::set_error("cudaFree of a pointer that cudaMalloc did not return");
  return cudaErrorInvalidValue;
}
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
// synchronous copies run on the legacy default stream: they do not wait for non-blocking streams
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t st) {
  if (st && st->cap) return cudaErrorInvalidValue;
  cusim::submit(st, [dst, src, n]() { memmove(dst, src, n); });
  return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new cusimStream(); cusim::g_streams.push_back(*s); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  if (!s) return cudaSuccess;
  cusim::drain(s);
  auto& v = cusim::g_streams;
  v.erase(std::remove(v.begin(), v.end(), s), v.end());
  delete s;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t st) { cusim::drain(st); return cusim::g_failed.load() ? cudaErrorLaunchFailure : cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned) {
  cusimStream* src = ev->st;            // the event's state at the time of THIS call is what is waited for
  const unsigned long long seq = ev->seq;
  if (!src || src == st) return cudaSuccess;
  cusim::submit(st, [src, seq]() { cusim::drain_until(src, seq); });
  return cudaSuccess;
}
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode) {
  if (!s || s->cap) return cudaErrorInvalidValue;
  s->cap = new cusimGraph();
  return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g) {
  if (!s || !s->cap) return cudaErrorInvalidValue;
  *g = s->cap;
  s->cap = nullptr;
  return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* x, cudaGraph_t g, unsigned long long) { *x = new cusimGraph(*g); return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t x, cudaStream_t st) {
  for (const auto& op : x->ops) cusim::launch(op.grid, op.block, st, op.body);   // bodies are held by value
  return cusim::g_failed.load() ? cudaErrorLaunchFailure : cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cusimEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new cusimEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st) {
  if (st && st->cap) return cudaErrorInvalidValue;
  cusim::submit(st, [e]() { e->ns = cusim::clock_ns(); });
  e->st = st;
  e->seq = st ? st->enq : 0;
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  cusim::drain_until(a->st, a->seq);
  cusim::drain_until(b->st, b->seq);
  *ms = (float)((b->ns - a->ns) * 1e-6);
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

cudaError_t cudaLibraryLoadData(cudaLibrary_t* lib, const void* code, void*, void*, unsigned, void*, void*, unsigned) {
  void* dl = dlopen((const char*)code, RTLD_NOW | RTLD_LOCAL);
  if (!dl) { cusim::set_error(std::string("dlopen: ") + dlerror()); return cudaErrorInvalidValue; }
  *lib = new cusimLibrary{dl};
  return cudaSuccess;
}
cudaError_t cudaLibraryUnload(cudaLibrary_t lib) {
  if (lib) { dlclose(lib->dl); delete lib; }
  return cudaSuccess;
}
cudaError_t cudaLibraryGetKernel(cudaKernel_t* k, cudaLibrary_t lib, const char* name) {
  void* f = dlsym(lib->dl, name);
  if (!f) { cusim::set_error(std::string("dlsym: ") + name); return cudaErrorInvalidValue; }
  *k = f;
  return cudaSuccess;
}
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t, cudaStream_t st) {
  if (st && st->cap) return cudaErrorInvalidValue;   // argument pointers do not outlive the call
  auto f = (void (*)(void**))func;
  cusim::drain(st);                                   // ... so this launch cannot be deferred: keep stream order, run now
  cusim::launch(grid, block, nullptr, [=]() { f(args); });
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------------
// "NVRTC": g++ of the generated translation unit into a shared object
// ---------------------------------------------------------------------------------------------------------------------
struct cusimProgram {
  std::string src, log, so_path;
  std::vector<std::string> names;
};

namespace {
std::string self_path() {
  Dl_info info;
  if (dladdr((void*)&cudaLaunchKernel, &info) && info.dli_fname) return info.dli_fname;
  return "";
}
unsigned long long fnv1a(const std::string& s) {
  unsigned long long h = 1469598103934665603ull;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}
}  // namespace

nvrtcResult nvrtcCreateProgram(nvrtcProgram* p, const char* src, const char*, int, const char* const*, const char* const*) {
  *p = new cusimProgram();
  (*p)->src = src;
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcDestroyProgram(nvrtcProgram* p) { delete *p; *p = nullptr; return NVRTC_SUCCESS; }
nvrtcResult nvrtcAddNameExpression(nvrtcProgram p, const char* e) { p->names.push_back(e); return NVRTC_SUCCESS; }
nvrtcResult nvrtcCompileProgram(nvrtcProgram p, int, const char* const*) {
  std::string body = p->src;
  // the generated source declares the fixed-width integer typedefs itself (NVRTC has no <stdint.h>); here they come from
  // the system header
  for (const char* td : {"typedef unsigned char uint8_t;", "typedef unsigned short uint16_t;", "typedef int int32_t;", "typedef long long int64_t;"}) {
    const size_t at = body.find(td);
    if (at != std::string::npos) body.erase(at, strlen(td));
  }
  std::string tu = "#define __CUDACC_RTC__ 1\n#define ND_CUSIM 1\n#include \"cusim_device.h\"\n" + body + "\n";
  for (size_t k = 0; k < p->names.size(); ++k)
    tu += "extern \"C\" void cusim_entry_" + std::to_string(k) + "(void** a) { cusim::invoke(&" + p->names[k] + ", a); }\n";
  const std::string lib = self_path();
  const std::string inc = CUSIM_INCLUDE_DIR;
  const char* tmp = getenv("CUSIM_CACHE");
  const std::string dir = tmp ? tmp : "/tmp/cusim_cache";
  mkdir(dir.c_str(), 0777);
  char name[64];
  snprintf(name, sizeof name, "%016llx", fnv1a(tu + lib));
  const std::string base = dir + "/" + name;
  p->so_path = base + ".so";
  struct stat sb;
  if (stat(p->so_path.c_str(), &sb) == 0) return NVRTC_SUCCESS;
  const std::string cpp = base + "." + std::to_string((long)getpid()) + ".cpp", tmpso = base + "." + std::to_string((long)getpid()) + ".so.tmp", logf = base + "." + std::to_string((long)getpid()) + ".log";
  FILE* f = fopen(cpp.c_str(), "w");
  if (!f) { p->log = "cannot write " + cpp; return NVRTC_ERROR_COMPILATION; }
  fwrite(tu.data(), 1, tu.size(), f);
  fclose(f);
  const std::string cmd = "g++ -O1 -std=c++17 -fPIC -shared -ffp-contract=off -w -I'" + inc + "' '" + cpp + "' -o '" + tmpso + "' '" + lib + "' > '" + logf + "' 2>&1";
  const int rc = system(cmd.c_str());
  if (rc != 0) {
    if (FILE* lf = fopen(logf.c_str(), "r")) {
      char buf[4096];
      size_t n;
      while ((n = fread(buf, 1, sizeof buf, lf)) > 0) p->log.append(buf, n);
      fclose(lf);
    }
    if (p->log.empty()) p->log = "g++ failed";
    unlink(logf.c_str());
    return NVRTC_ERROR_COMPILATION;
  }
  unlink(logf.c_str());
  unlink(cpp.c_str());
  rename(tmpso.c_str(), p->so_path.c_str());
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram p, size_t* n) { *n = p->log.size() + 1; return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetProgramLog(nvrtcProgram p, char* out) { memcpy(out, p->log.c_str(), p->log.size() + 1); return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram p, size_t* n) { *n = p->so_path.size() + 1; return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetCUBIN(nvrtcProgram p, char* out) { memcpy(out, p->so_path.c_str(), p->so_path.size() + 1); return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetLoweredName(nvrtcProgram p, const char* expr, const char** lowered) {
  static thread_local std::string s;
  for (size_t k = 0; k < p->names.size(); ++k)
    if (p->names[k] == expr) { s = "cusim_entry_" + std::to_string(k); *lowered = s.c_str(); return NVRTC_SUCCESS; }
  return NVRTC_ERROR_INVALID_INPUT;
}
