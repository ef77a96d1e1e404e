// Exchange step of the k-sharded single-cosmology mode.  Each rank owns a
// contiguous block of k-rows; every integral evaluation needs the full ln P_ab, so the
// rank's block of the three ln P components is all-gathered in place before it, and the
// RKF45 error norm is max-reduced so that all ranks take the identical accept/reject
// decision.  Two transports:
//   NcclExchange      one process per GPU, NCCL over NVLink/NVSwitch (libnccl is dlopen'ed, so
//                     the batch-only use of the library has no NCCL dependency)
//   LoopbackExchange  all ranks are handles of ONE process (each driven by its own host thread);
//                     blocks move with cudaMemcpyPeerAsync.  Used to test the k-sharded path on a
//                     single GPU and as a no-NCCL peer-copy path inside one process.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <string>
#include <vector>

namespace rtrg {

struct Segment {
  void *base;        // start of the full array on this rank
  size_t per_rank;   // bytes per rank; rank r owns [base + r*per_rank, base + (r+1)*per_rank)
};

class Exchange {
 public:
  virtual ~Exchange() {}
  virtual int nranks() const = 0;
  virtual int rank() const = 0;
  // true when allgather / allreduce only enqueue work on the stream (CUDA-graph capturable)
  virtual bool capturable() const = 0;
  // true when the exchange consists of plain kernel / memcpy nodes only, so that it may sit in the
  // body of a conditional (while) graph node
  virtual bool device_side() const { return false; }
  virtual const char *name() const = 0;
  // in-place all-gather of every segment, ordered on `st`; returns 0 or sets err
  virtual int allgather(const std::vector<Segment> &segs, cudaStream_t st, std::string *err) = 0;
  // in-place max over ranks of n unsigned 64-bit values (bit patterns of non-negative doubles)
  virtual int allreduce_max_u64(unsigned long long *dev, size_t n, cudaStream_t st, std::string *err) = 0;
};

// NCCL
int nccl_unique_id(char id[128], std::string *err);
std::unique_ptr<Exchange> make_nccl_exchange(const char id[128], int nranks, int rank, std::string *err);

// in-process loopback: create the shared group once, then one Exchange per rank
struct LoopbackGroup;
std::shared_ptr<LoopbackGroup> make_loopback_group(int nranks);
std::unique_ptr<Exchange> make_loopback_exchange(std::shared_ptr<LoopbackGroup> g, int rank, int device);

}  // namespace rtrg
