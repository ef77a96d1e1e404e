// Transports of the k-shard exchange (see kshard.h).
#include "kshard.h"
#include "rtrg_device.h"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is resolved at run time

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace rtrg {

// ------------------------------------------------------------------------ pack / unpack
// one block per segment (grid-stride), threads along the rank's part of the segment
__global__ void k_xch_pack(GatherPlan plan, const double *__restrict__ base, int rank, double *__restrict__ send) {
  for (int s = blockIdx.x; s < plan.nseg; s += gridDim.x) {
    const int len = plan.len[s];
    const double *src = base + plan.off[s] + (long long)rank * len;
    double *dst = send + plan.prefix[s];
    RT_ASSERT(len >= 0 && plan.prefix[s] >= 0 && plan.prefix[s] + len <= plan.total);
    for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
  }
}
// recv: [nranks][total] packed blocks in rank order; the rank's own part is already in place
__global__ void k_xch_unpack(GatherPlan plan, double *__restrict__ base, int rank, int nranks,
                             const double *__restrict__ recv) {
  for (int s = blockIdx.x; s < plan.nseg; s += gridDim.x) {
    const int len = plan.len[s];
    for (int r = 0; r < nranks; r++) {
      if (r == rank) continue;
      const double *src = recv + (long long)r * plan.total + plan.prefix[s];
      double *dst = base + plan.off[s] + (long long)r * len;
      for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
    }
  }
}
static inline int xch_blocks(const GatherPlan &p) { return p.nseg < 1 ? 1 : (p.nseg > 1024 ? 1024 : p.nseg); }

// ------------------------------------------------------------------------ P2P mailboxes
// Mailbox of one rank (in ITS device memory, mapped into every peer by cudaIpc):
//   data [2][G][cap]  8-byte words written by source rank g
// Every 8-byte word on the wire is self-validating (the "LL" scheme of NCCL's low-latency protocol):
// 32 bits of payload + a 32-bit tag derived from the exchange number, written by ONE store, so the
// receiver simply polls the word until the tag matches -- no fence, no separate flag, no second trip
// over NVLink: the latency of an exchange is one remote store plus the poll.  A double travels as two
// such words (one 16-byte vector store; each half is validated on its own).
// Two slots alternate with the parity of the exchange number: a rank can run at most one exchange
// ahead of the slowest peer (it needs everybody's words of exchange n to start n + 1), so the slot
// it overwrites in exchange n + 1 holds words of exchange n - 1, which every peer has consumed, and
// their tags differ from the ones polled for.
struct P2pDev {
  int rank, nranks;
  long long cap;                       // words per (slot, source rank); payload capacity cap / 2
  char *const *peer;                   // [G] device table: base of every rank's mailbox (own included)
  unsigned long long *seq;             // exchanges completed by this rank
  int *err;                            // set when a wait gave up
  __device__ ulonglong2 *data(int owner, int slot, int src) const {
    return reinterpret_cast<ulonglong2 *>(reinterpret_cast<unsigned long long *>(peer[owner]) + 2 * nranks +
                                          ((long long)slot * nranks + src) * cap);
  }
};
__device__ __forceinline__ void st_ll(ulonglong2 *p, unsigned long long v, unsigned long long tag) {
  const unsigned long long w0 = (v & 0xffffffffULL) | (tag << 32), w1 = (v >> 32) | (tag << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
// poll until both halves carry the tag (bounded: ~10 s of SM clock -- ranks may enter rtrg_run seconds
// apart, e.g. after building weight tables; a peer that never comes sets err)
__device__ __forceinline__ unsigned long long ld_ll(const ulonglong2 *p, unsigned long long tag, int *err) {
  unsigned long long w0, w1;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
    if (clock64() - t0 > 20000000000LL) {
      atomicExch(err, 1);
      return 0ULL;
    }
  }
  return (w0 & 0xffffffffULL) | (w1 << 32);
}

// ONE kernel per exchange: store this rank's words into every peer's mailbox over NVLink, then
// collect the peers' words from the own mailbox: MODE 0 scatters the ln P rows into `base`, MODE 1
// reduces vals[i] = max over ranks.  One CTA: the payload is 3 nk / G doubles per cosmology.
template <int MODE>
__global__ void __launch_bounds__(512) k_xch_p2p(P2pDev d, GatherPlan plan, double *__restrict__ base,
                                                 unsigned long long *__restrict__ vals, int n) {
  const int tid = threadIdx.x, G = d.nranks, me = d.rank;
  if (*d.err) return;  // a peer is gone: do not wait for it again
  const unsigned long long s = *d.seq, want = s + 1;
  const unsigned long long tag = (want & 0x7fffffffULL) | 0x80000000ULL;  // never 0 (the mailbox starts zeroed)
  const int slot = (int)(s & 1ULL);
  // 1. remote stores
  if (MODE == 0) {
    for (int sg = 0; sg < plan.nseg; sg++) {
      const int len = plan.len[sg];
      const long long pre = plan.prefix[sg];
      const unsigned long long *src =
          reinterpret_cast<const unsigned long long *>(base + plan.off[sg] + (long long)me * len);
      RT_ASSERT(2 * (pre + len) <= d.cap);
      for (int i = tid; i < len; i += blockDim.x) {
        const unsigned long long v = src[i];
        for (int r = 0; r < G; r++)
          if (r != me) st_ll(d.data(r, slot, me) + pre + i, v, tag);
      }
    }
  } else {
    RT_ASSERT(2LL * n <= d.cap);
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long v = vals[i];
      for (int r = 0; r < G; r++)
        if (r != me) st_ll(d.data(r, slot, me) + i, v, tag);
    }
  }
  // 2. collect from my mailbox (the words are written by the peers: volatile loads, straight from L2)
  if (MODE == 0) {
    for (int sg = 0; sg < plan.nseg; sg++) {
      const int len = plan.len[sg];
      const long long pre = plan.prefix[sg];
      for (int r = 0; r < G; r++) {
        if (r == me) continue;
        const ulonglong2 *src = d.data(me, slot, r) + pre;
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(base + plan.off[sg] + (long long)r * len);
        for (int i = tid; i < len; i += blockDim.x) dst[i] = ld_ll(src + i, tag, d.err);
      }
    }
  } else {
    for (int i = tid; i < n; i += blockDim.x) {
      unsigned long long m = vals[i];
      for (int r = 0; r < G; r++) {
        if (r == me) continue;
        const unsigned long long v = ld_ll(d.data(me, slot, r) + i, tag, d.err);
        m = v > m ? v : m;
      }
      vals[i] = m;
    }
  }
  __syncthreads();
  if (tid == 0) *d.seq = want;
}

// ------------------------------------------------------------------------------------ NCCL
namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi *nccl_api(std::string *err) {
  static NcclApi api;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!tried) {
    tried = true;
    // if the host program (e.g. torch) already loaded an NCCL under this soname, share it
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      api.lib = lib;
#define RT_SYM(field, name) *(void **)(&api.field) = dlsym(lib, name)
      RT_SYM(GetUniqueId, "ncclGetUniqueId");
      RT_SYM(CommInitRank, "ncclCommInitRank");
      RT_SYM(CommDestroy, "ncclCommDestroy");
      RT_SYM(AllGather, "ncclAllGather");
      RT_SYM(AllReduce, "ncclAllReduce");
      RT_SYM(GetErrorString, "ncclGetErrorString");
#undef RT_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce ||
          !api.GetErrorString)
        api.lib = nullptr;
    }
  }
  if (!api.lib) {
    if (err) *err = "libnccl.so.2 could not be loaded";
    return nullptr;
  }
  return &api;
}

// NCCL communicator; exchanges go through peer mailboxes when the mappings could be set up on
// EVERY rank (agreed by an all-reduce), through pack -> ncclAllGather -> unpack otherwise and for
// bulk transfers.  A rank that dies inside a collective leaves the others blocked in NCCL: the
// application has to ncclCommAbort / kill the job (as with any NCCL program); the P2P waits give
// up after ~10 s and report through check().
class NcclExchange : public Exchange {
 public:
  NcclExchange(const NcclApi *api, ncclComm_t comm, int nranks, int rank, int device)
      : api_(api), comm_(comm), nranks_(nranks), rank_(rank), device_(device) {}
  ~NcclExchange() override {
    cudaSetDevice(device_);
    for (int r = 0; r < (int)peer_.size(); r++)
      if (r != rank_ && peer_[r]) cudaIpcCloseMemHandle(peer_[r]);
    if (mailbox_) cudaFree(mailbox_);
    if (d_peer_) cudaFree(d_peer_);
    if (d_state_) cudaFree(d_state_);
    if (send_) cudaFree(send_);
    if (recv_) cudaFree(recv_);
    if (comm_) api_->CommDestroy(comm_);
  }
  int nranks() const override { return nranks_; }
  int rank() const override { return rank_; }
  bool capturable() const override { return true; }
  bool device_side() const override { return p2p_; }
  const char *name() const override {
    return p2p_ ? "P2P mailboxes over NVLink (one kernel per exchange: self-validating 8-byte words stored into peer memory, polled by the receiver; "
                  "NCCL for bootstrap and the final bulk gather)"
                : "NCCL: pack -> one ncclAllGather -> unpack per exchange, ncclAllReduce(max) per attempt";
  }

  // peer mailboxes: allocate, exchange the IPC handles through the communicator, map the peers
  void setup_p2p(long long cap_words) {
    const char *force = std::getenv("RTRG_KSHARD_TRANSPORT");
    int ok = !(force && std::strcmp(force, "nccl") == 0);
    const size_t bytes = (size_t)(2 * nranks_) * sizeof(unsigned long long) * (size_t)(1 + cap_words);
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof mine);
    if (ok && (cudaMalloc(&mailbox_, bytes) != cudaSuccess || cudaMemset(mailbox_, 0, bytes) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine, mailbox_) != cudaSuccess))
      ok = 0;
    // all-gather of the handles (device staging buffer; NCCL is the only channel the ranks share)
    char *d_h = nullptr;
    std::vector<cudaIpcMemHandle_t> all(nranks_);
    cudaStream_t st = nullptr;
    bool comm_ok = cudaMalloc((void **)&d_h, sizeof(mine) * nranks_) == cudaSuccess &&
                   cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    if (comm_ok) {
      cudaMemcpyAsync(d_h + sizeof(mine) * rank_, &mine, sizeof mine, cudaMemcpyHostToDevice, st);
      comm_ok = api_->AllGather(d_h + sizeof(mine) * rank_, d_h, sizeof mine, ncclChar, comm_, st) == ncclSuccess &&
                cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * nranks_, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                cudaStreamSynchronize(st) == cudaSuccess;
    }
    peer_.assign(nranks_, nullptr);
    if (ok && comm_ok) {
      for (int r = 0; r < nranks_ && ok; r++) {
        if (r == rank_) {
          peer_[r] = mailbox_;
        } else if (cudaIpcOpenMemHandle(&peer_[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          peer_[r] = nullptr;
          ok = 0;
        }
      }
    } else {
      ok = 0;
    }
    if (ok) {
      ok = cudaMalloc((void **)&d_peer_, sizeof(char *) * nranks_) == cudaSuccess &&
           cudaMalloc((void **)&d_state_, 64) == cudaSuccess && cudaMemset(d_state_, 0, 64) == cudaSuccess &&
           cudaMemcpy(d_peer_, peer_.data(), sizeof(char *) * nranks_, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    // every rank must take the same transport: min over ranks of "it worked here"
    int all_ok = 0;
    if (comm_ok) {
      int *d_ok = reinterpret_cast<int *>(d_h);
      cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, st);
      if (api_->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm_, st) == ncclSuccess &&
          cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
          cudaStreamSynchronize(st) == cudaSuccess) {
      } else {
        all_ok = 0;
      }
    }
    cudaGetLastError();
    if (d_h) cudaFree(d_h);
    if (st) cudaStreamDestroy(st);
    p2p_ = all_ok != 0;
    cap_ = cap_words;
  }

  int reserve(long long max_total, size_t n_u64, std::string *err) override {
    (void)n_u64;
    if (max_total <= send_cap_) return 0;
    if (send_) cudaFree(send_);
    if (recv_) cudaFree(recv_);
    send_ = recv_ = nullptr;
    send_cap_ = 0;
    if (cudaMalloc((void **)&send_, sizeof(double) * (size_t)max_total) != cudaSuccess ||
        cudaMalloc((void **)&recv_, sizeof(double) * (size_t)max_total * nranks_) != cudaSuccess) {
      if (err) *err = "k-shard exchange: staging buffers";
      cudaGetLastError();
      return -1;
    }
    send_cap_ = max_total;
    return 0;
  }
  int gather(double *base, const GatherPlan &plan, bool small, cudaStream_t st, std::string *err) override {
    if (p2p_ && small && 2 * plan.total <= cap_) {
      k_xch_p2p<0><<<1, 512, 0, st>>>(dev(), plan, base, nullptr, 0);
      return 0;
    }
    if (plan.total > send_cap_) {
      if (err) *err = "k-shard exchange: reserve() was not called for this plan";
      return -1;
    }
    k_xch_pack<<<xch_blocks(plan), 128, 0, st>>>(plan, base, rank_, send_);
    const ncclResult_t r = api_->AllGather(send_, recv_, (size_t)plan.total, ncclDouble, comm_, st);
    if (check_nccl(r, err)) return -1;
    k_xch_unpack<<<xch_blocks(plan), 128, 0, st>>>(plan, base, rank_, nranks_, recv_);
    return 0;
  }
  int allreduce_max_u64(unsigned long long *dev_vals, size_t n, cudaStream_t st, std::string *err) override {
    if (p2p_ && 2 * (long long)n <= cap_) {
      k_xch_p2p<1><<<1, 512, 0, st>>>(dev(), GatherPlan(), nullptr, dev_vals, (int)n);
      return 0;
    }
    return check_nccl(api_->AllReduce(dev_vals, dev_vals, n, ncclUint64, ncclMax, comm_, st), err);
  }
  int check(std::string *err) override {
    if (!p2p_) return 0;
    int e = 0;
    if (cudaMemcpy(&e, reinterpret_cast<int *>(d_state_ + 8), sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) e = 1;
    if (e && err) *err = "a peer rank did not arrive at an exchange within the time limit";
    return e ? -1 : 0;
  }

 private:
  P2pDev dev() const {
    P2pDev d;
    d.rank = rank_, d.nranks = nranks_, d.cap = cap_;
    d.peer = reinterpret_cast<char *const *>(d_peer_);
    d.seq = reinterpret_cast<unsigned long long *>(d_state_);
    d.err = reinterpret_cast<int *>(d_state_ + 8);
    return d;
  }
  int check_nccl(ncclResult_t r, std::string *err) {
    if (r == ncclSuccess) return 0;
    if (err) *err = std::string("NCCL: ") + api_->GetErrorString(r);
    return -1;
  }
  const NcclApi *api_;
  ncclComm_t comm_;
  int nranks_, rank_, device_;
  double *send_ = nullptr, *recv_ = nullptr;
  long long send_cap_ = 0;
  bool p2p_ = false;
  long long cap_ = 0;
  void *mailbox_ = nullptr;
  std::vector<void *> peer_;
  char *d_peer_ = nullptr, *d_state_ = nullptr;
};
}  // namespace

int nccl_unique_id(char id[128], std::string *err) {
  const NcclApi *api = nccl_api(err);
  if (!api) return -1;
  ncclUniqueId u;
  const ncclResult_t r = api->GetUniqueId(&u);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r);
    return -1;
  }
  std::memcpy(id, u.internal, 128);
  return 0;
}

std::unique_ptr<Exchange> make_nccl_exchange(const char id[128], int nranks, int rank, int device, std::string *err) {
  const NcclApi *api = nccl_api(err);
  if (!api) return nullptr;
  ncclUniqueId u;
  std::memcpy(u.internal, id, 128);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api->CommInitRank(&comm, nranks, u, rank);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
    return nullptr;
  }
  NcclExchange *x = new NcclExchange(api, comm, nranks, rank, device);
  x->setup_p2p(/*cap_words=*/1 << 16);  // 512 KB per (slot, source): 3 nk/G doubles x up to ~680 cosmologies at nk=256, G=8
  return std::unique_ptr<Exchange>(x);
}

// -------------------------------------------------------------------------------- loopback
struct LoopbackGroup {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long generation = 0;
  bool aborted = false;
  std::vector<const double *> send;   // packed block published per rank
  std::vector<int> device;
  std::vector<std::vector<unsigned long long>> vals;
  explicit LoopbackGroup(int n_) : n(n_), send(n_, nullptr), device(n_, 0), vals(n_) {}
  // false: a rank left with an error, nobody waits any more
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return false;
    const long long gen = generation;
    if (++arrived == n) {
      arrived = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen || aborted; });
    }
    return !aborted;
  }
  void abort() {
    {
      std::lock_guard<std::mutex> lk(mu);
      aborted = true;
    }
    cv.notify_all();
  }
};

std::shared_ptr<LoopbackGroup> make_loopback_group(int nranks) { return std::make_shared<LoopbackGroup>(nranks); }

namespace {
class LoopbackExchange : public Exchange {
 public:
  LoopbackExchange(std::shared_ptr<LoopbackGroup> g, int rank, int device) : g_(g), rank_(rank), device_(device) {
    g_->device[rank] = device;
  }
  ~LoopbackExchange() override {
    cudaSetDevice(device_);
    if (send_) cudaFree(send_);
    if (recv_) cudaFree(recv_);
  }
  int nranks() const override { return g_->n; }
  int rank() const override { return rank_; }
  bool capturable() const override { return false; }  // host-side rendezvous between the ranks
  const char *name() const override { return "in-process loopback (one peer copy per rank, host rendezvous)"; }
  void abort() override { g_->abort(); }
  int reserve(long long max_total, size_t n_u64, std::string *err) override {
    (void)n_u64;
    if (max_total <= cap_) return 0;
    if (send_) cudaFree(send_);
    if (recv_) cudaFree(recv_);
    send_ = recv_ = nullptr;
    cap_ = 0;
    if (cudaMalloc((void **)&send_, sizeof(double) * (size_t)max_total) != cudaSuccess ||
        cudaMalloc((void **)&recv_, sizeof(double) * (size_t)max_total * g_->n) != cudaSuccess)
      return fail(err, "staging buffers");
    cap_ = max_total;
    return 0;
  }
  int gather(double *base, const GatherPlan &plan, bool small, cudaStream_t st, std::string *err) override {
    (void)small;
    if (plan.total > cap_) return fail(err, "reserve() was not called for this plan");
    k_xch_pack<<<xch_blocks(plan), 128, 0, st>>>(plan, base, rank_, send_);
    // my block must be complete before anybody reads it
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(err, nullptr);
    g_->send[rank_] = send_;
    if (!g_->barrier()) return fail(err, "another rank aborted");
    for (int r = 0; r < g_->n; r++) {  // ONE peer copy per rank
      if (r == rank_) continue;
      if (cudaMemcpyPeerAsync(recv_ + (size_t)r * plan.total, g_->device[rank_], g_->send[r], g_->device[r],
                              sizeof(double) * (size_t)plan.total, st) != cudaSuccess)
        return fail(err, nullptr);
    }
    k_xch_unpack<<<xch_blocks(plan), 128, 0, st>>>(plan, base, rank_, g_->n, recv_);
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(err, nullptr);
    if (!g_->barrier()) return fail(err, "another rank aborted");  // nobody repacks while others still read
    return 0;
  }
  int allreduce_max_u64(unsigned long long *dev, size_t n, cudaStream_t st, std::string *err) override {
    std::vector<unsigned long long> mine(n);
    if (cudaMemcpyAsync(mine.data(), dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail(err, nullptr);
    g_->vals[rank_] = mine;
    if (!g_->barrier()) return fail(err, "another rank aborted");
    for (int r = 0; r < g_->n; r++)
      for (size_t i = 0; i < n; i++)
        if (g_->vals[r][i] > mine[i]) mine[i] = g_->vals[r][i];
    if (!g_->barrier()) return fail(err, "another rank aborted");
    if (cudaMemcpyAsync(dev, mine.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail(err, nullptr);
    return 0;
  }

 private:
  int fail(std::string *err, const char *what) {
    if (err) *err = std::string("loopback exchange: ") + (what ? what : cudaGetErrorString(cudaGetLastError()));
    g_->abort();  // whoever waits for this rank must not wait for ever
    return -1;
  }
  std::shared_ptr<LoopbackGroup> g_;
  int rank_, device_;
  double *send_ = nullptr, *recv_ = nullptr;
  long long cap_ = 0;
};
}  // namespace

std::unique_ptr<Exchange> make_loopback_exchange(std::shared_ptr<LoopbackGroup> g, int rank, int device) {
  return std::unique_ptr<Exchange>(new LoopbackExchange(g, rank, device));
}

}  // namespace rtrg
