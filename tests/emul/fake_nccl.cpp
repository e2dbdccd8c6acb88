// In-process stand-in for the few NCCL entry points libcfdl dlopens (comm.cu), for the cuemu multi-rank
// tests: the "ranks" are threads of one process, a send is a copy into a mailbox, an all-reduce a
// rendezvous that combines in rank order.  TEST INFRASTRUCTURE ONLY (loaded through CFDL_NCCL_PATH by
// tests/emul/multirank_check.py); it lets the NCCL exchange mode of the library — pack / send / recv
// of ghost values per colour, all-reduced residual norms and dot products, the broadcast of pc(1) —
// run on a box without GPUs or NCCL.  Streams are ignored (the emulated runtime is synchronous).
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

namespace {

struct World {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<std::vector<char>>> box;  // (src, dst) -> messages in order
  // rendezvous for collectives
  int arrived = 0;
  unsigned long long gen = 0;
  std::vector<std::vector<char>> slot;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long g = gen;
    if (++arrived == nranks) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

std::mutex g_mu;
std::map<unsigned long long, World*> g_worlds;
unsigned long long g_next_id = 1;

struct Comm { World* w; int rank; };
struct Op { int kind; const void* send; void* recv; size_t bytes; int peer, op; Comm* c; };  // kind 0 send, 1 recv, 2 allreduce, 3 bcast
thread_local int t_depth = 0;
thread_local std::vector<Op> t_ops;

void do_send(const Op& o) {
  World* w = o.c->w;
  std::vector<char> m((const char*)o.send, (const char*)o.send + o.bytes);
  { std::lock_guard<std::mutex> lk(w->mu); w->box[{o.c->rank, o.peer}].push_back(std::move(m)); }
  w->cv.notify_all();
}
void do_recv(const Op& o) {
  World* w = o.c->w;
  std::unique_lock<std::mutex> lk(w->mu);
  auto& q = w->box[{o.peer, o.c->rank}];
  w->cv.wait(lk, [&] { return !q.empty(); });
  std::memcpy(o.recv, q.front().data(), o.bytes < q.front().size() ? o.bytes : q.front().size());
  q.pop_front();
}
void do_allreduce(const Op& o) {  // doubles; op 0 = sum, 2 = max; combined in rank order on every rank
  World* w = o.c->w;
  { std::lock_guard<std::mutex> lk(w->mu); w->slot[o.c->rank].assign((const char*)o.send, (const char*)o.send + o.bytes); }
  w->barrier();
  const size_t n = o.bytes / 8;
  std::vector<double> acc(n);
  for (int r = 0; r < w->nranks; ++r) {
    const double* v = (const double*)w->slot[r].data();
    for (size_t i = 0; i < n; ++i) acc[i] = r == 0 ? v[i] : (o.op == 2 ? (v[i] > acc[i] ? v[i] : acc[i]) : acc[i] + v[i]);
  }
  w->barrier();  // everyone has read the slots before they are reused
  std::memcpy(o.recv, acc.data(), o.bytes);
}
void do_bcast(const Op& o) {
  World* w = o.c->w;
  if (o.c->rank == o.peer) { std::lock_guard<std::mutex> lk(w->mu); w->slot[o.peer].assign((const char*)o.send, (const char*)o.send + o.bytes); }
  w->barrier();
  std::memcpy(o.recv, w->slot[o.peer].data(), o.bytes);
  w->barrier();
}
void run(const std::vector<Op>& ops) {
  for (const Op& o : ops) if (o.kind == 0) do_send(o);  // all sends are posted before any receive blocks
  for (const Op& o : ops) {
    if (o.kind == 1) do_recv(o);
    else if (o.kind == 2) do_allreduce(o);
    else if (o.kind == 3) do_bcast(o);
  }
}
int submit(const Op& o) {
  if (t_depth > 0) { t_ops.push_back(o); return 0; }
  run(std::vector<Op>{o});
  return 0;
}

}  // namespace

extern "C" {

int ncclGetUniqueId(char* id128) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::memset(id128, 0, 128);
  const unsigned long long id = g_next_id++;
  std::memcpy(id128, &id, sizeof id);
  g_worlds[id] = new World;
  return 0;
}
struct Id128 { char b[128]; };
int ncclCommInitRank(void** comm, int nranks, Id128 id, int rank) {
  unsigned long long key;
  std::memcpy(&key, id.b, sizeof key);
  World* w;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_worlds.find(key);
    if (it == g_worlds.end()) return 1;
    w = it->second;
    std::lock_guard<std::mutex> lk2(w->mu);
    if (w->nranks == 0) { w->nranks = nranks; w->slot.resize(nranks); }
  }
  *comm = new Comm{w, rank};
  w->barrier();
  return 0;
}
int ncclCommDestroy(void* comm) { delete (Comm*)comm; return 0; }
int ncclGroupStart() { ++t_depth; return 0; }
int ncclGroupEnd() {
  if (--t_depth == 0) { std::vector<Op> ops; ops.swap(t_ops); run(ops); }
  return 0;
}
int ncclSend(const void* buf, size_t count, int, int peer, void* comm, void*) { return submit(Op{0, buf, nullptr, count * 8, peer, 0, (Comm*)comm}); }
int ncclRecv(void* buf, size_t count, int, int peer, void* comm, void*) { return submit(Op{1, nullptr, buf, count * 8, peer, 0, (Comm*)comm}); }
int ncclAllReduce(const void* s, void* r, size_t count, int, int op, void* comm, void*) { return submit(Op{2, s, r, count * 8, 0, op, (Comm*)comm}); }
int ncclBroadcast(const void* s, void* r, size_t count, int, int root, void* comm, void*) { return submit(Op{3, s, r, count * 8, root, 0, (Comm*)comm}); }
const char* ncclGetErrorString(int) { return "fake nccl error"; }

}  // extern "C"
