// Double-buffered batch pipeline on top of the C-ABI (include/redtime_b200.h, rtrg_pipeline_*).
//
// What it replaces: the sequential loop of scripts/runRedTimeBatch:91-99, where model i+1 is not
// even parsed before model i has been written.  Here a job (one batch of cosmologies) passes
// through two stages that overlap across jobs:
//
//   stage thread : rtrg_add_cosmologies  (host staging + H2D on the handle's copy stream)
//                  rtrg_prepare          (device-side linear theory, sigma_8, 1-loop cache)
//   run thread   : rtrg_run              (Time-RG evolution, output integrals, tables)
//   fetch thread : rtrg_fetch_outputs    (D2H into page-locked memory of the handle; the next job's
//                                         evolution is already running on the other handle)
//
// `depth` handles (2 = double buffering) live on the same device, each with its own streams and
// work arena, so batch i+1 is staged, uploaded and initialised while batch i evolves: the host
// work and the PCIe transfers disappear behind the GPU time of the previous batch.  Every result
// is bit-identical to the serial add/prepare/run sequence -- the handles are independent.
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/redtime_b200.h"

namespace {
struct Job {
  long long ticket = 0;
  std::vector<const rtrg_cosmology *> list;
  int slot = 0;
  int rc = RTRG_OK;
  std::string err;
  bool staged = false, done = false, released = false;
  const double *out = nullptr, *hdr = nullptr, *hdr0 = nullptr;
  size_t out_len = 0;
  std::vector<int> status;
  double t[6] = {0, 0, 0, 0, 0, 0};  // seconds since pipeline creation: stage, run, fetch begin / end
};
}  // namespace

struct rtrg_pipeline {
  std::vector<rtrg_handle *> handles;
  std::vector<bool> slot_busy;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job *> to_stage, to_run, to_fetch;
  std::vector<Job *> jobs;  // by ticket (never shrinks during the life of the pipeline; small)
  long long next_ticket = 0;
  bool stop = false;
  std::thread stage_thread, run_thread, fetch_thread;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double now() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

static void stage_loop(rtrg_pipeline *p) {
  for (;;) {
    Job *j = nullptr;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv.wait(lk, [&] { return p->stop || (!p->to_stage.empty() && !p->slot_busy[p->to_stage.front()->slot]); });
      if (p->stop) return;
      j = p->to_stage.front();
      p->to_stage.pop_front();
      p->slot_busy[j->slot] = true;
    }
    rtrg_handle *h = p->handles[j->slot];
    j->t[0] = p->now();
    int rc = rtrg_clear_cosmologies(h);
    if (rc == RTRG_OK) rc = rtrg_add_cosmologies(h, (int)j->list.size(), j->list.data());
    if (rc == RTRG_OK) rc = rtrg_prepare(h);
    j->t[1] = p->now();
    {
      std::lock_guard<std::mutex> lk(p->mu);
      j->rc = rc;
      if (rc != RTRG_OK) j->err = rtrg_last_error();
      j->staged = true;
      p->to_run.push_back(j);
    }
    p->cv.notify_all();
  }
}

static void run_loop(rtrg_pipeline *p) {
  for (;;) {
    Job *j = nullptr;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv.wait(lk, [&] { return p->stop || !p->to_run.empty(); });
      if (p->stop) return;
      j = p->to_run.front();
      p->to_run.pop_front();
    }
    int rc = j->rc;
    std::string err = j->err;
    j->t[2] = p->now();
    if (rc == RTRG_OK) {
      rtrg_handle *h = p->handles[j->slot];
      j->status.assign(j->list.size(), 0);
      rc = rtrg_run(h, nullptr, 0, nullptr, nullptr, j->status.data());
      if (rc == RTRG_EODE) rc = RTRG_OK;  // per-cosmology failures are reported through status[]
      if (rc != RTRG_OK) err = rtrg_last_error();
    }
    j->t[3] = p->now();
    {
      std::lock_guard<std::mutex> lk(p->mu);
      j->rc = rc;
      j->err = err;
      p->to_fetch.push_back(j);
    }
    p->cv.notify_all();
  }
}

static void fetch_loop(rtrg_pipeline *p) {
  for (;;) {
    Job *j = nullptr;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv.wait(lk, [&] { return p->stop || !p->to_fetch.empty(); });
      if (p->stop) return;
      j = p->to_fetch.front();
      p->to_fetch.pop_front();
    }
    int rc = j->rc;
    std::string err = j->err;
    j->t[4] = p->now();
    if (rc == RTRG_OK) {
      rc = rtrg_fetch_outputs(p->handles[j->slot], &j->out, &j->out_len, &j->hdr, &j->hdr0);
      if (rc != RTRG_OK) err = rtrg_last_error();
    }
    j->t[5] = p->now();
    {
      std::lock_guard<std::mutex> lk(p->mu);
      j->rc = rc;
      j->err = err;
      j->done = true;
    }
    p->cv.notify_all();
  }
}

extern "C" {

int rtrg_pipeline_create(const rtrg_config *cfg, int depth, rtrg_pipeline **out) {
  if (!cfg || !out || depth < 1 || depth > 8) return RTRG_EINVAL;
  *out = nullptr;
  rtrg_pipeline *p = new rtrg_pipeline();
  for (int i = 0; i < depth; i++) {
    rtrg_handle *h = nullptr;
    const int rc = rtrg_create(cfg, &h);
    if (rc != RTRG_OK) {
      for (rtrg_handle *x : p->handles) rtrg_destroy(x);
      delete p;
      return rc;
    }
    p->handles.push_back(h);
  }
  p->slot_busy.assign(depth, false);
  p->stage_thread = std::thread(stage_loop, p);
  p->run_thread = std::thread(run_loop, p);
  p->fetch_thread = std::thread(fetch_loop, p);
  *out = p;
  return RTRG_OK;
}

int rtrg_pipeline_submit(rtrg_pipeline *p, int n, const rtrg_cosmology *const *list, long long *ticket) {
  if (!p || n <= 0 || !list) return RTRG_EINVAL;
  Job *j = new Job();
  j->list.assign(list, list + n);
  {
    std::lock_guard<std::mutex> lk(p->mu);
    j->ticket = p->next_ticket++;
    j->slot = (int)(j->ticket % (long long)p->handles.size());
    p->jobs.push_back(j);
    p->to_stage.push_back(j);
    if (ticket) *ticket = j->ticket;
  }
  p->cv.notify_all();
  return RTRG_OK;
}

int rtrg_pipeline_wait(rtrg_pipeline *p, long long ticket, const double **out, size_t *out_len, const double **hdr,
                       const double **hdr0, const int **status) {
  if (!p) return RTRG_EINVAL;
  std::unique_lock<std::mutex> lk(p->mu);
  if (ticket < 0 || ticket >= (long long)p->jobs.size() || !p->jobs[ticket] || p->jobs[ticket]->released)
    return RTRG_EINVAL;
  Job *j = p->jobs[ticket];
  p->cv.wait(lk, [&] { return j->done; });
  if (out) *out = j->out;
  if (out_len) *out_len = j->out_len;
  if (hdr) *hdr = j->hdr;
  if (hdr0) *hdr0 = j->hdr0;
  if (status) *status = j->status.data();
  return j->rc;
}

int rtrg_pipeline_columns(rtrg_pipeline *p, long long ticket, int icosmo) {
  if (!p) return RTRG_EINVAL;
  std::lock_guard<std::mutex> lk(p->mu);
  if (ticket < 0 || ticket >= (long long)p->jobs.size() || !p->jobs[ticket] || !p->jobs[ticket]->done ||
      p->jobs[ticket]->released)
    return RTRG_EINVAL;
  return rtrg_num_columns(p->handles[p->jobs[ticket]->slot], icosmo);
}

int rtrg_pipeline_times(rtrg_pipeline *p, long long ticket, double t[6]) {
  if (!p || !t) return RTRG_EINVAL;
  std::lock_guard<std::mutex> lk(p->mu);
  if (ticket < 0 || ticket >= (long long)p->jobs.size() || !p->jobs[ticket] || !p->jobs[ticket]->done) return RTRG_EINVAL;
  for (int i = 0; i < 6; i++) t[i] = p->jobs[ticket]->t[i];
  return RTRG_OK;
}

int rtrg_pipeline_release(rtrg_pipeline *p, long long ticket) {
  if (!p) return RTRG_EINVAL;
  {
    std::unique_lock<std::mutex> lk(p->mu);
    if (ticket < 0 || ticket >= (long long)p->jobs.size() || !p->jobs[ticket] || p->jobs[ticket]->released)
      return RTRG_EINVAL;
    Job *j = p->jobs[ticket];
    p->cv.wait(lk, [&] { return j->done; });
    j->released = true;
    p->slot_busy[j->slot] = false;
    j->list.clear();
    j->status.clear();
    j->status.shrink_to_fit();
  }
  p->cv.notify_all();
  return RTRG_OK;
}

int rtrg_pipeline_destroy(rtrg_pipeline *p) {
  if (!p) return RTRG_OK;
  {
    std::unique_lock<std::mutex> lk(p->mu);
    // let the jobs in flight finish: their handles are about to be destroyed
    p->cv.wait(lk, [&] {
      for (Job *j : p->jobs)
        if (j && !j->done && (j->staged || p->slot_busy[j->slot])) return false;
      return true;
    });
    p->stop = true;
  }
  p->cv.notify_all();
  if (p->stage_thread.joinable()) p->stage_thread.join();
  if (p->run_thread.joinable()) p->run_thread.join();
  if (p->fetch_thread.joinable()) p->fetch_thread.join();
  for (Job *j : p->jobs) delete j;
  for (rtrg_handle *h : p->handles) rtrg_destroy(h);
  delete p;
  return RTRG_OK;
}

}  // extern "C"
