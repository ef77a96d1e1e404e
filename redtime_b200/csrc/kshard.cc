// Transports of the k-shard exchange (see kshard.h).
#include "kshard.h"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is resolved at run time

#include <condition_variable>
#include <cstring>
#include <mutex>

namespace rtrg {

// ------------------------------------------------------------------------------------ NCCL
namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
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
      RT_SYM(GroupStart, "ncclGroupStart");
      RT_SYM(GroupEnd, "ncclGroupEnd");
      RT_SYM(GetErrorString, "ncclGetErrorString");
#undef RT_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce ||
          !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
        api.lib = nullptr;
    }
  }
  if (!api.lib) {
    if (err) *err = "libnccl.so.2 could not be loaded";
    return nullptr;
  }
  return &api;
}

class NcclExchange : public Exchange {
 public:
  NcclExchange(const NcclApi *api, ncclComm_t comm, int nranks, int rank)
      : api_(api), comm_(comm), nranks_(nranks), rank_(rank) {}
  ~NcclExchange() override {
    if (comm_) api_->CommDestroy(comm_);
  }
  int nranks() const override { return nranks_; }
  int rank() const override { return rank_; }
  bool capturable() const override { return true; }
  const char *name() const override { return "NCCL all-gather / all-reduce per exchange"; }
  int allgather(const std::vector<Segment> &segs, cudaStream_t st, std::string *err) override {
    ncclResult_t r = api_->GroupStart();
    for (const Segment &s : segs) {
      if (r != ncclSuccess) break;
      // in place: sendbuff = recvbuff + rank * count
      r = api_->AllGather((const char *)s.base + (size_t)rank_ * s.per_rank, s.base, s.per_rank, ncclChar, comm_, st);
    }
    const ncclResult_t r2 = api_->GroupEnd();
    if (r == ncclSuccess) r = r2;
    return check(r, err);
  }
  int allreduce_max_u64(unsigned long long *dev, size_t n, cudaStream_t st, std::string *err) override {
    return check(api_->AllReduce(dev, dev, n, ncclUint64, ncclMax, comm_, st), err);
  }

 private:
  int check(ncclResult_t r, std::string *err) {
    if (r == ncclSuccess) return 0;
    if (err) *err = std::string("NCCL: ") + api_->GetErrorString(r);
    return -1;
  }
  const NcclApi *api_;
  ncclComm_t comm_;
  int nranks_, rank_;
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

std::unique_ptr<Exchange> make_nccl_exchange(const char id[128], int nranks, int rank, std::string *err) {
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
  return std::unique_ptr<Exchange>(new NcclExchange(api, comm, nranks, rank));
}

// -------------------------------------------------------------------------------- loopback
struct LoopbackGroup {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long generation = 0;
  std::vector<std::vector<Segment>> segs;      // published per rank
  std::vector<int> device;
  std::vector<std::vector<unsigned long long>> vals;
  explicit LoopbackGroup(int n_) : n(n_), segs(n_), device(n_, 0), vals(n_) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const long long gen = generation;
    if (++arrived == n) {
      arrived = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

std::shared_ptr<LoopbackGroup> make_loopback_group(int nranks) { return std::make_shared<LoopbackGroup>(nranks); }

namespace {
class LoopbackExchange : public Exchange {
 public:
  LoopbackExchange(std::shared_ptr<LoopbackGroup> g, int rank, int device) : g_(g), rank_(rank) {
    g_->device[rank] = device;
  }
  int nranks() const override { return g_->n; }
  int rank() const override { return rank_; }
  bool capturable() const override { return false; }  // host-side rendezvous between the ranks
  const char *name() const override { return "in-process loopback (peer copies, host rendezvous)"; }
  int allgather(const std::vector<Segment> &segs, cudaStream_t st, std::string *err) override {
    // my block must be complete before anybody reads it
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(err);
    g_->segs[rank_] = segs;
    g_->barrier();
    // pull every other rank's block into my arrays
    for (int r = 0; r < g_->n; r++) {
      if (r == rank_) continue;
      for (size_t i = 0; i < segs.size(); i++) {
        const size_t off = (size_t)r * segs[i].per_rank;
        if (cudaMemcpyPeerAsync((char *)segs[i].base + off, g_->device[rank_], (const char *)g_->segs[r][i].base + off,
                                g_->device[r], segs[i].per_rank, st) != cudaSuccess)
          return fail(err);
      }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(err);
    g_->barrier();  // nobody overwrites its block while others still read it
    return 0;
  }
  int allreduce_max_u64(unsigned long long *dev, size_t n, cudaStream_t st, std::string *err) override {
    std::vector<unsigned long long> mine(n);
    if (cudaMemcpyAsync(mine.data(), dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail(err);
    g_->vals[rank_] = mine;
    g_->barrier();
    for (int r = 0; r < g_->n; r++)
      for (size_t i = 0; i < n; i++)
        if (g_->vals[r][i] > mine[i]) mine[i] = g_->vals[r][i];
    g_->barrier();
    if (cudaMemcpyAsync(dev, mine.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail(err);
    return 0;
  }

 private:
  int fail(std::string *err) {
    if (err) *err = std::string("loopback exchange: ") + cudaGetErrorString(cudaGetLastError());
    return -1;
  }
  std::shared_ptr<LoopbackGroup> g_;
  int rank_;
};
}  // namespace

std::unique_ptr<Exchange> make_loopback_exchange(std::shared_ptr<LoopbackGroup> g, int rank, int device) {
  return std::unique_ptr<Exchange>(new LoopbackExchange(g, rank, device));
}

}  // namespace rtrg
