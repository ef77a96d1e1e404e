// Exchange step of the k-sharded single-cosmology mode.  Each rank owns a
// contiguous block of k-rows; every integral evaluation needs the full ln P_ab, so the
// rank's block of the three ln P components is gathered in place before it, and the
// RKF45 error norm is max-reduced so that all ranks take the identical accept/reject
// decision.  What moves is described once per batch by a GatherPlan: a list of segments, each
// the same array slice on every rank, of which rank r owns the r-th part.  Every transport packs
// its own parts into ONE contiguous block (k_xch_pack), moves one block per peer and scatters
// what arrived (k_xch_unpack) -- never one transfer per segment.  Three transports:
//   P2pExchange       one process per GPU; the packed block is STORED straight into a mailbox in
//                     every peer's memory over NVLink (cudaIpc mappings, bootstrapped through the
//                     NCCL communicator), followed by a sequence flag; the same kernel then waits
//                     for the peers' flags and unpacks.  One kernel per exchange, no host or NCCL
//                     call on the critical path, and -- being plain kernel nodes -- it sits inside
//                     the conditional WHILE graph of rtrg_run.
//   NcclExchange      pack -> ONE ncclAllGather -> unpack (fallback when peer mappings are not
//                     available; libnccl is dlopen'ed, so the batch-only use of the library has no
//                     NCCL dependency)
//   LoopbackExchange  all ranks are handles of ONE process (each driven by its own host thread);
//                     blocks move with cudaMemcpyPeerAsync.  Used to test the k-sharded path on a
//                     single GPU (kernels of different ranks must never wait on one another there).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <string>
#include <vector>

namespace rtrg {

// Device-resident description of one gather: segment s is base[off[s] .. off[s] + nranks * len[s]),
// rank r owning the r-th part of len[s] doubles; prefix[s] = sum of len[0..s) locates the segment
// in a rank's packed block of `total` doubles.
struct GatherPlan {
  int nseg = 0;
  long long total = 0;                // doubles per rank
  const long long *off = nullptr;     // [nseg] device
  const int *len = nullptr;           // [nseg] device
  const long long *prefix = nullptr;  // [nseg] device
};

class Exchange {
 public:
  virtual ~Exchange() {}
  virtual int nranks() const = 0;
  virtual int rank() const = 0;
  virtual const char *name() const = 0;
  // true when gather / allreduce only enqueue work on the stream (CUDA-graph capturable)
  virtual bool capturable() const = 0;
  // true when an exchange consists of kernel nodes only, so that it may sit in the body of a
  // conditional (while) graph node
  virtual bool device_side() const { return false; }
  // called by rtrg_prepare (on every rank) with the largest packed block of the batch
  virtual int reserve(long long max_total_doubles, size_t n_u64, std::string *err) = 0;
  // in-place gather of every segment of the plan, ordered on `st`; returns 0 or sets err.
  // small = false: a bulk transfer outside the evolution loop (the output tables)
  virtual int gather(double *base, const GatherPlan &plan, bool small, cudaStream_t st, std::string *err) = 0;
  // in-place max over ranks of n unsigned 64-bit values (bit patterns of non-negative doubles)
  virtual int allreduce_max_u64(unsigned long long *dev, size_t n, cudaStream_t st, std::string *err) = 0;
  // after the stream has been synchronised: did a device-side wait give up (a peer never arrived)?
  virtual int check(std::string *err) {
    (void)err;
    return 0;
  }
  // a rank is leaving with an error: wake whoever waits for it (loopback; NCCL needs ncclCommAbort)
  virtual void abort() {}
};

// NCCL communicator (+ P2P mailboxes over it unless RTRG_KSHARD_TRANSPORT=nccl or peer mappings fail)
int nccl_unique_id(char id[128], std::string *err);
std::unique_ptr<Exchange> make_nccl_exchange(const char id[128], int nranks, int rank, int device, std::string *err);

// in-process loopback: create the shared group once, then one Exchange per rank
struct LoopbackGroup;
std::shared_ptr<LoopbackGroup> make_loopback_group(int nranks);
std::unique_ptr<Exchange> make_loopback_exchange(std::shared_ptr<LoopbackGroup> g, int rank, int device);

}  // namespace rtrg
