// fake_nccl.cc -- TEST INFRASTRUCTURE: a stand-in for libnccl.so.2 that lets several PROCESSES, each running
// the CPU emulation of the library (tests/emu/libcracks_b200_emu.so) as one rank, exchange data the way the
// library asks NCCL to: grouped ncclSend / ncclRecv between neighbours and ncclAllReduce (sum / max on
// float64, uint64, uint8).  Transport: one Unix-domain socket pair per pair of ranks, set up in
// ncclCommInitRank through a directory named after the unique id.  Blocking, in-order, no streams (the
// emulated CUDA runtime is synchronous).  The library dlopen()s "libnccl.so.2" by name, so pointing
// LD_LIBRARY_PATH at this directory is all it takes (tests/test_multirank_emulation_cpu.py).
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

namespace {

struct Comm
{
  int rank = 0, nranks = 1;
  std::vector<int> fd; // fd[peer]
};

struct Op
{
  bool send;
  void *buf;
  size_t bytes;
  int peer;
  Comm *comm;
};

thread_local int group_depth = 0;
thread_local std::vector<Op> pending;

size_t
type_size (int t)
{
  switch (t)
    {
    case 0: case 1: return 1; // int8 / uint8
    case 2: case 3: case 7: return 4;
    case 6: return 2;
    default: return 8;        // int64, uint64, float64
    }
}

void
write_all (int fd, const void *p, size_t n)
{
  const char *c = static_cast<const char *> (p);
  while (n)
    {
      const ssize_t k = ::write (fd, c, n);
      if (k < 0)
        {
          if (errno == EINTR)
            continue;
          std::perror ("fake_nccl write");
          std::abort ();
        }
      c += k;
      n -= (size_t) k;
    }
}

void
read_all (int fd, void *p, size_t n)
{
  char *c = static_cast<char *> (p);
  while (n)
    {
      const ssize_t k = ::read (fd, c, n);
      if (k <= 0)
        {
          if (k < 0 && errno == EINTR)
            continue;
          std::fprintf (stderr, "fake_nccl read: peer closed\n");
          std::abort ();
        }
      c += k;
      n -= (size_t) k;
    }
}

void
run_ops (std::vector<Op> &ops)
{
  // sends on helper threads so that two neighbours sending to each other first cannot dead-lock
  std::vector<std::thread> senders;
  for (Op &o : ops)
    if (o.send)
      senders.emplace_back ([o]() { write_all (o.comm->fd[(size_t) o.peer], o.buf, o.bytes); });
  for (Op &o : ops)
    if (!o.send)
      read_all (o.comm->fd[(size_t) o.peer], o.buf, o.bytes);
  for (auto &t : senders)
    t.join ();
  ops.clear ();
}

template <class T>
void
reduce_into (T *acc, const T *in, size_t n, int op)
{
  for (size_t i = 0; i < n; ++i)
    acc[i] = op == 2 ? (in[i] > acc[i] ? in[i] : acc[i]) : (T) (acc[i] + in[i]);
}

} // namespace

extern "C" {

typedef struct
{
  char internal[128];
} ncclUniqueId;

int
ncclGetUniqueId (ncclUniqueId *id)
{
  std::memset (id, 0, sizeof *id);
  std::snprintf (id->internal, sizeof id->internal, "/tmp/pf_fake_nccl_%d_%ld", (int) getpid (), (long) random ());
  return 0;
}

// every pair (a < b): b listens on <dir>/<a>_<b>, a connects
int
ncclCommInitRank (void **out, int nranks, ncclUniqueId id, int rank)
{
  Comm *c = new Comm;
  c->rank = rank;
  c->nranks = nranks;
  c->fd.assign ((size_t) nranks, -1);
  const std::string dir = id.internal;
  ::mkdir (dir.c_str (), 0700);
  std::vector<int> listeners ((size_t) nranks, -1);
  for (int a = 0; a < rank; ++a)
    {
      const std::string path = dir + "/" + std::to_string (a) + "_" + std::to_string (rank);
      const int s = ::socket (AF_UNIX, SOCK_STREAM, 0);
      sockaddr_un addr{};
      addr.sun_family = AF_UNIX;
      std::strncpy (addr.sun_path, path.c_str (), sizeof addr.sun_path - 1);
      ::unlink (path.c_str ());
      if (::bind (s, (sockaddr *) &addr, sizeof addr) || ::listen (s, 1))
        {
          std::perror ("fake_nccl bind/listen");
          return 1;
        }
      listeners[(size_t) a] = s;
    }
  for (int b = rank + 1; b < nranks; ++b)
    {
      const std::string path = dir + "/" + std::to_string (rank) + "_" + std::to_string (b);
      const int s = ::socket (AF_UNIX, SOCK_STREAM, 0);
      sockaddr_un addr{};
      addr.sun_family = AF_UNIX;
      std::strncpy (addr.sun_path, path.c_str (), sizeof addr.sun_path - 1);
      for (int tries = 0;; ++tries)
        {
          if (::connect (s, (sockaddr *) &addr, sizeof addr) == 0)
            break;
          if (tries > 3000)
            {
              std::perror ("fake_nccl connect");
              return 1;
            }
          ::usleep (10000);
        }
      c->fd[(size_t) b] = s;
    }
  for (int a = 0; a < rank; ++a)
    {
      c->fd[(size_t) a] = ::accept (listeners[(size_t) a], nullptr, nullptr);
      ::close (listeners[(size_t) a]);
      if (c->fd[(size_t) a] < 0)
        {
          std::perror ("fake_nccl accept");
          return 1;
        }
    }
  *out = c;
  return 0;
}

int
ncclCommDestroy (void *comm)
{
  Comm *c = static_cast<Comm *> (comm);
  for (int f : c->fd)
    if (f >= 0)
      ::close (f);
  delete c;
  return 0;
}

int
ncclGroupStart ()
{
  ++group_depth;
  return 0;
}

int
ncclGroupEnd ()
{
  if (--group_depth == 0)
    run_ops (pending);
  return 0;
}

int
ncclSend (const void *buf, size_t count, int type, int peer, void *comm, void *)
{
  pending.push_back ({true, const_cast<void *> (buf), count * type_size (type), peer, static_cast<Comm *> (comm)});
  if (group_depth == 0)
    run_ops (pending);
  return 0;
}

int
ncclRecv (void *buf, size_t count, int type, int peer, void *comm, void *)
{
  pending.push_back ({false, buf, count * type_size (type), peer, static_cast<Comm *> (comm)});
  if (group_depth == 0)
    run_ops (pending);
  return 0;
}

// gather on rank 0, reduce in rank order (deterministic), broadcast
int
ncclAllReduce (const void *send, void *recv, size_t count, int type, int op, void *comm, void *)
{
  Comm *c = static_cast<Comm *> (comm);
  const size_t bytes = count * type_size (type);
  if (recv != send)
    std::memmove (recv, send, bytes);
  if (c->rank == 0)
    {
      std::vector<char> tmp (bytes);
      for (int r = 1; r < c->nranks; ++r)
        {
          read_all (c->fd[(size_t) r], tmp.data (), bytes);
          if (type == 8)
            reduce_into (static_cast<double *> (recv), reinterpret_cast<const double *> (tmp.data ()), count, op);
          else if (type == 5)
            reduce_into (static_cast<uint64_t *> (recv), reinterpret_cast<const uint64_t *> (tmp.data ()), count, op);
          else if (type == 7)
            reduce_into (static_cast<float *> (recv), reinterpret_cast<const float *> (tmp.data ()), count, op);
          else if (type == 1)
            reduce_into (static_cast<uint8_t *> (recv), reinterpret_cast<const uint8_t *> (tmp.data ()), count, op);
          else
            {
              std::fprintf (stderr, "fake_nccl: all-reduce type %d not implemented\n", type);
              return 1;
            }
        }
      for (int r = 1; r < c->nranks; ++r)
        write_all (c->fd[(size_t) r], recv, bytes);
    }
  else
    {
      write_all (c->fd[0], recv, bytes);
      read_all (c->fd[0], recv, bytes);
    }
  return 0;
}

const char *
ncclGetErrorString (int)
{
  return "fake_nccl error";
}
}
