// emu_runtime.cc -- definitions behind tests/emu/cuda_shim_full/cuda_runtime.h (test infrastructure)
#include <cuda_runtime.h>

#include <ucontext.h>

#include <condition_variable>
#include <cstdio>
#include <mutex>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace pf {
// `extern __shared__ unsigned char smem_raw[]` of the tiled kernels: one buffer per OS thread = per running block
alignas (16) thread_local unsigned char smem_raw[256 * 1024];
} // namespace pf

namespace pf_emu {
thread_local double shfl_slots[1024];
long long launches_sequential = 0, launches_cooperative = 0;

void
fail (const char *what)
{
  std::fprintf (stderr, "pf_emu: %s\n", what);
  std::abort ();
}

// ---- the CUDA threads of one cooperative block as fibers of the OS thread that runs the block ----------
namespace {
constexpr size_t STACK = 256 * 1024;
struct Fibers
{
  ucontext_t main;
  std::vector<ucontext_t *> ctx; // by pointer: a ucontext_t points into itself (fpregs) and must not move
  std::vector<char *> stack;
  std::vector<char> done;
  unsigned current = 0;
  bool active = false;
  void (*body) (void *, unsigned) = nullptr;
  void *job = nullptr;
};
thread_local Fibers *fibers = nullptr;

void
fiber_entry ()
{
  Fibers &f = *fibers;
  f.body (f.job, f.current);
  f.done[f.current] = 1;
  // returning resumes uc_link = main
}
} // namespace

void
sync ()
{
  Fibers *f = fibers;
  if (!f || !f->active)
    fail ("a kernel launched in sequential mode reached a barrier / shuffle: add it to COOPERATIVE in "
          "tests/emu/build_emulated_library.py");
  swapcontext (f->ctx[f->current], &f->main);
}

void
run_block (bool cooperative, unsigned block, void (*body) (void *, unsigned), void *job)
{
  if (!cooperative)
    {
      for (unsigned t = 0; t < block; ++t)
        body (job, t);
      return;
    }
  if (!fibers)
    fibers = new Fibers; // leaked with its OS thread
  Fibers &f = *fibers;
  while (f.ctx.size () < block)
    {
      f.ctx.push_back (new ucontext_t);
      f.stack.push_back (static_cast<char *> (std::malloc (STACK)));
      f.done.push_back (0);
      getcontext (f.ctx.back ());
    }
  f.body = body;
  f.job = job;
  f.active = true;
  for (unsigned t = 0; t < block; ++t)
    {
      f.ctx[t]->uc_stack.ss_sp = f.stack[t];
      f.ctx[t]->uc_stack.ss_size = STACK;
      f.ctx[t]->uc_link = &f.main;
      makecontext (f.ctx[t], fiber_entry, 0);
      f.done[t] = 0;
    }
  // rounds: every live fiber runs to its next barrier (or to the end); barriers sit in uniform control flow
  for (unsigned live = block; live > 0;)
    {
      live = 0;
      for (unsigned t = 0; t < block; ++t)
        if (!f.done[t])
          {
            f.current = t;
            threadIdx.x = t; // the body sets it on entry only; restore it on every resume
            swapcontext (&f.main, f.ctx[t]);
            live += !f.done[t];
          }
    }
  f.active = false;
}

// ---- the blocks of a launch on a pool of OS threads -----------------------------------------------------
namespace {
struct Workers
{
  std::mutex m;
  std::condition_variable cv_start, cv_done;
  std::vector<std::thread> threads;
  unsigned long long generation = 0;
  unsigned running = 0;
  std::atomic<unsigned> next{0};
  unsigned grid = 0;
  void (*per_block) (void *, unsigned) = nullptr;
  void *job = nullptr;

  void work ()
  {
    for (unsigned b; (b = next.fetch_add (1)) < grid;)
      per_block (job, b);
  }
  explicit Workers (unsigned n)
  {
    for (unsigned i = 0; i < n; ++i)
      threads.emplace_back ([this]() {
        unsigned long long seen = 0;
        for (;;)
          {
            {
              std::unique_lock<std::mutex> lk (m);
              cv_start.wait (lk, [&] { return generation != seen; });
              seen = generation;
            }
            work ();
            {
              std::lock_guard<std::mutex> lk (m);
              if (--running == 0)
                cv_done.notify_one ();
            }
          }
      });
    for (auto &t : threads)
      t.detach (); // they idle on cv_start for the life of the process
  }
};

Workers *
workers ()
{
  static Workers *w = [] {
    const char *e = std::getenv ("PF_EMU_THREADS");
    unsigned n = e ? (unsigned) std::atoi (e) : std::thread::hardware_concurrency ();
    if (n > 16)
      n = 16;
    return n > 1 ? new Workers (n - 1) : nullptr; // the launching thread works too
  }();
  return w;
}
} // namespace

void
run_grid (unsigned grid, unsigned block, void (*per_block) (void *, unsigned), void *job)
{
  Workers *w = workers ();
  // small launches are not worth waking the pool
  if (!w || grid < 2 || (unsigned long long) grid * block < 2048)
    {
      for (unsigned b = 0; b < grid; ++b)
        per_block (job, b);
      return;
    }
  {
    std::lock_guard<std::mutex> lk (w->m);
    w->grid = grid;
    w->per_block = per_block;
    w->job = job;
    w->next = 0;
    w->running = (unsigned) w->threads.size ();
    ++w->generation;
  }
  w->cv_start.notify_all ();
  w->work ();
  std::unique_lock<std::mutex> lk (w->m);
  w->cv_done.wait (lk, [&] { return w->running == 0; });
}
} // namespace pf_emu

extern "C" void
pf_emu_launch_counts (long long *sequential, long long *cooperative)
{
  *sequential = pf_emu::launches_sequential;
  *cooperative = pf_emu::launches_cooperative;
}

// PF_EMU_BACKTRACE=1: print a native backtrace on SIGSEGV (there is no debugger in the image)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
namespace {
void
segv_handler (int, siginfo_t *si, void *uc)
{
  const ucontext_t *u = static_cast<const ucontext_t *> (uc);
  char buf[200];
  const int len = std::snprintf (buf, sizeof buf, "pf_emu: SIGSEGV at address %p, rip %p, rsp %p\n", si->si_addr,
                                 (void *) u->uc_mcontext.gregs[REG_RIP], (void *) u->uc_mcontext.gregs[REG_RSP]);
  if (write (2, buf, (size_t) len) < 0)
    _exit (139);
  void **sp = reinterpret_cast<void **> (u->uc_mcontext.gregs[REG_RSP]);
  backtrace_symbols_fd (sp, 1, 2); // the return address of a call through a null pointer
  void *frames[64];
  const int n = backtrace (frames, 64);
  backtrace_symbols_fd (frames, n, 2);
  _exit (139);
}
struct InstallSegv
{
  InstallSegv ()
  {
    if (!std::getenv ("PF_EMU_BACKTRACE"))
      return;
    static char alt[1 << 16];
    stack_t ss{};
    ss.ss_sp = alt;
    ss.ss_size = sizeof alt;
    sigaltstack (&ss, nullptr);
    struct sigaction sa{};
    sa.sa_sigaction = segv_handler;
    sa.sa_flags = SA_ONSTACK | SA_SIGINFO;
    sigaction (SIGSEGV, &sa, nullptr);
  }
} install_segv;
} // namespace
