// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

#include <atomic>
#include <thread>

namespace gcm_emu {

thread_local Block* tl_block = nullptr;
thread_local uint3 tl_threadIdx, tl_blockIdx;
thread_local dim3 tl_blockDim, tl_gridDim;

static const size_t kStack = 64 * 1024;

static void fiber_entry() {
  Block* b = tl_block;
  (*b->body)();
  b->fibers[b->cur].state = 3;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

void yield_state(int state) {
  Block* b = tl_block;
  if (b->nthreads == 1 && state == 1) return;
  b->fibers[b->cur].state = state;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

static void set_tid(int t, dim3 bd) {
  tl_threadIdx.x = t % bd.x;
  tl_threadIdx.y = (t / bd.x) % bd.y;
  tl_threadIdx.z = t / (bd.x * bd.y);
}

static void run_block(Block& b, dim3 bd) {
  int n = b.nthreads;
  for (int t = 0; t < n; ++t) {
    Fiber& f = b.fibers[t];
    f.state = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &b.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  int done = 0;
  while (done < n) {
    bool progressed = false;
    for (int t = 0; t < n; ++t) {
      if (b.fibers[t].state != 0) continue;
      b.cur = t;
      set_tid(t, bd);
      swapcontext(&b.sched, &b.fibers[t].ctx);
      progressed = true;
      if (b.fibers[t].state == 3) ++done;
    }
    // release barriers whose participants have all arrived (exited threads count as arrived)
    int waiting = 0;
    for (int t = 0; t < n; ++t) waiting += b.fibers[t].state == 1;
    bool released = false;
    if (waiting > 0 && waiting + done == n) {
      for (int t = 0; t < n; ++t)
        if (b.fibers[t].state == 1) b.fibers[t].state = 0;
      released = true;
    }
    for (int w = 0; w * 32 < n; ++w) {
      int lo = w * 32, hi = std::min(n, lo + 32), ww = 0, wd = 0;
      for (int t = lo; t < hi; ++t) { ww += b.fibers[t].state == 2; wd += b.fibers[t].state == 3; }
      if (ww > 0 && ww + wd == hi - lo) {
        for (int t = lo; t < hi; ++t)
          if (b.fibers[t].state == 2) b.fibers[t].state = 0;
        released = true;
      }
    }
    if (!progressed && !released && done < n) {
      fprintf(stderr, "gcm_emu: deadlock (divergent barrier) in block (%u,%u,%u)\n", tl_blockIdx.x, tl_blockIdx.y, tl_blockIdx.z);
      abort();
    }
  }
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nblocks == 0 || nthreads == 0) return;
  unsigned hw = std::thread::hardware_concurrency();
  const char* env = getenv("GCM_EMU_THREADS");
  unsigned nworkers = env ? (unsigned)atoi(env) : (hw ? hw : 1);
  nworkers = (unsigned)std::max<size_t>(1, std::min<size_t>(nworkers, nblocks));
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    Block b;
    b.nthreads = nthreads;
    b.body = &body;
    b.fibers.resize(nthreads);
    std::vector<char> stacks((size_t)nthreads * kStack);
    for (int t = 0; t < nthreads; ++t) b.fibers[t].stack = stacks.data() + (size_t)t * kStack;
    std::vector<char> smem(smem_bytes + 64);
    b.smem = (char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    tl_block = &b;
    tl_blockDim = block;
    tl_gridDim = grid;
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      tl_blockIdx.x = (unsigned)(i % grid.x);
      tl_blockIdx.y = (unsigned)((i / grid.x) % grid.y);
      tl_blockIdx.z = (unsigned)(i / ((size_t)grid.x * grid.y));
      run_block(b, block);
    }
    tl_block = nullptr;
  };
  if (nworkers == 1) { worker(); return; }
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < nworkers; ++w) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
}

}  // namespace gcm_emu
