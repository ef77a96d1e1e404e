// Batch front-end, C++ only: what replaces the sequential loop of scripts/runRedTimeBatch:91-99
// (one `cd $OUTPUT_DIR; redTime > redTime_$MODEL.dat` process per model, scripts/runRedTime:196-229)
// with ONE GPU pass over all models, file to file.
//
//   redTimeBatch_b200 <manifest> [first [stride]]
//
// <manifest>: text file, one run directory per line ('#' comments), each holding params_redTime.dat
// and the CAMB files it names.  For every directory D the table is written to
// D/redTime_<basename(D)>.dat, byte-compatible with the reference's stdout.  `first`/`stride`
// select lines first, first+stride, ... so that N processes (one per GPU, RTRG_DEVICE=g) share
// one manifest without any collective.  Environment knobs as for redTime_b200, plus
// RTRG_BATCH_CHUNK (models per GPU batch, default 256).
//
// Three stages overlap, chunk by chunk:
//   reader thread   rtrg_read_run_dirs of chunk c+1 (mapped files, host threads)
//   GPU             rtrg_pipeline_*: staging / H2D / initialisation of chunk c overlap the evolution of
//                   chunk c-1 (two handles)
//   main thread     waits for chunk c-1 and writes its redTime_<MODEL>.dat files on host threads
// so the wall time is that of the slowest stage, not the sum.  The last line on stdout reports the
// stage times: "redTimeBatch_b200: N models, F failed, T s wall (start-up S s, parse P s, gpu wait G s,
// write W s)".
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/redtime_b200.h"

static int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Chunk {
  int begin = 0, n = 0;
  std::vector<rtrg_run_inputs *> in;
  std::vector<const rtrg_cosmology *> cos;
  long long ticket = -1;
  int rc = RTRG_OK;
};

// one model's table -> D/redTime_<basename(D)>.dat; returns true when the model counts as failed
static bool write_model(const std::string &dir, int nk, int ncols, int n_out, int status, const double *out,
                        const double *hd, const double *hd0) {
  std::string d = dir;
  while (d.size() > 1 && d.back() == '/') d.pop_back();
  const std::string model = d.substr(d.rfind('/') == std::string::npos ? 0 : d.rfind('/') + 1);
  const std::string path = d + "/redTime_" + model + ".dat";
  if (status == 101 || status == 103) {
    // the reference abort()s on these (hdr:528-531, 646-649): no table, and no stale one either
    std::remove(path.c_str());
    std::fprintf(stderr, "redTimeBatch_b200: %s: status %d, no table written\n", d.c_str(), status);
    return true;
  }
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) {
    std::fprintf(stderr, "redTimeBatch_b200: cannot write %s\n", path.c_str());
    return true;
  }
  std::setvbuf(f, nullptr, _IOFBF, 1 << 20);
  int n_done = n_out;
  if (status) {  // integrator failure: the outputs reached, then the warning (rt:1631-1632)
    n_done = 0;
    while (n_done < n_out && hd[(size_t)n_done * 5 + 1] != 0.0) n_done++;
  }
  if (n_done > 0) rtrg_print_result(f, "params_redTime.dat", nk, ncols, n_done, out, hd, hd0);
  if (status) std::fprintf(f, "#WARNING: integrator failed, status = %d\n", status);
  std::fclose(f);
  return status != 0;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <manifest> [first [stride]]\n", argv[0]);
    return 2;
  }
  const int first = argc > 2 ? std::atoi(argv[2]) : 0, stride = argc > 3 ? std::atoi(argv[3]) : 1;
  std::ifstream mf(argv[1]);
  if (!mf.is_open() || first < 0 || stride < 1) {
    std::fprintf(stderr, "redTimeBatch_b200: cannot open %s\n", argv[1]);
    return 2;
  }
  std::string base(argv[1]);
  base = base.find('/') == std::string::npos ? std::string(".") : base.substr(0, base.rfind('/'));
  std::vector<std::string> dirs;
  int line_no = 0;
  for (std::string line; std::getline(mf, line);) {
    line = line.substr(0, line.find('#'));
    const size_t a = line.find_first_not_of(" \t\r"), b = line.find_last_not_of(" \t\r");
    if (a == std::string::npos) continue;
    line = line.substr(a, b - a + 1);
    if (line_no >= first && (line_no - first) % stride == 0) dirs.push_back(line[0] == '/' ? line : base + "/" + line);
    line_no++;
  }
  if (dirs.empty()) return 0;

  rtrg_config cfg;
  rtrg_default_config(&cfg);
  cfg.nk = env_int("RTRG_NK", cfg.nk);
  cfg.device = env_int("RTRG_DEVICE", 0);
  cfg.print_A = env_int("RTRG_PRINTA", 0);
  cfg.print_I = env_int("RTRG_PRINTI", 0);
  cfg.print_Q = env_int("RTRG_PRINTQ", 0);
  cfg.print_bias = env_int("RTRG_PRINTBIAS", 0);
  cfg.reduce_beta = 1;  // only rtrg_run is used: send what it consumes
  if (env_int("RTRG_HIACC", 0)) cfg.beta_kmin = 1e-5, cfg.beta_kmax = 20.0, cfg.n_lnk = 1000, cfg.a_early = 1e-50;
  if (env_int("RTRG_HIGH_ACCURACY", 0)) cfg.nk = 512, cfg.eps_abs = 1e-15, cfg.eps_rel = 1e-6;
  const int camb_modern = env_int("RTRG_CAMB_MODERN", 0);
  const int n = (int)dirs.size();
  const int chunk_size = std::max(1, std::min(env_int("RTRG_BATCH_CHUNK", 256), n));
  const int n_chunks = (n + chunk_size - 1) / chunk_size;
  const double t_start = now_s();

  // ---- reader thread: parses one chunk ahead of the GPU (at most two parsed chunks wait)
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Chunk *> parsed;
  double t_parse = 0;
  bool reader_failed = false, stop = false;
  std::thread reader([&]() {
    for (int c = 0; c < n_chunks; c++) {
      {
        std::lock_guard<std::mutex> lk(mu);
        if (stop) return;
      }
      Chunk *ch = new Chunk();
      ch->begin = c * chunk_size;
      ch->n = std::min(chunk_size, n - ch->begin);
      ch->in.assign(ch->n, nullptr);
      std::vector<const char *> cdirs(ch->n);
      for (int i = 0; i < ch->n; i++) cdirs[i] = dirs[ch->begin + i].c_str();
      const double t0 = now_s();
      ch->rc = rtrg_read_run_dirs(ch->n, cdirs.data(), camb_modern, ch->in.data());
      t_parse += now_s() - t0;
      if (ch->rc == RTRG_OK) {
        ch->cos.resize(ch->n);
        for (int i = 0; i < ch->n; i++) ch->cos[i] = rtrg_inputs_cosmology(ch->in[i]);
      }
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return parsed.size() < 2 || stop; });
      parsed.push_back(ch);  // (drained by main when stopping)
      if (ch->rc != RTRG_OK) reader_failed = true;
      lk.unlock();
      cv.notify_all();
      if (ch->rc != RTRG_OK) return;
    }
  });

  rtrg_pipeline *pipe = nullptr;
  int rc = rtrg_pipeline_create(&cfg, 2, &pipe);
  const double t_startup = now_s() - t_start;  // CUDA context, two handles, weight tables
  if (rc != RTRG_OK) {
    std::fprintf(stderr, "redTimeBatch_b200: %s\n", rtrg_last_error());
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    reader.join();
    for (Chunk *ch : parsed) {
      for (rtrg_run_inputs *p : ch->in) rtrg_free_run_inputs(p);
      delete ch;
    }
    return 3;
  }

  int failed = 0, exit_code = 0;
  double t_gpu_wait = 0, t_write = 0;
  const int n_writers = std::max(1, std::min(32, (int)std::thread::hardware_concurrency()));
  std::deque<Chunk *> in_flight;
  auto finish_oldest = [&]() {
    Chunk *ch = in_flight.front();
    in_flight.pop_front();
    const double *out = nullptr, *hdr = nullptr, *hdr0 = nullptr;
    const int *status = nullptr;
    size_t len = 0;
    const double t0 = now_s();
    const int wrc = rtrg_pipeline_wait(pipe, ch->ticket, &out, &len, &hdr, &hdr0, &status);
    t_gpu_wait += now_s() - t0;
    if (wrc != RTRG_OK) {
      std::fprintf(stderr, "redTimeBatch_b200: models %d-%d: %s\n", ch->begin, ch->begin + ch->n - 1, rtrg_last_error());
      failed += ch->n;
      exit_code = 3;
    } else {
      const double t1 = now_s();
      std::vector<size_t> off(ch->n + 1, 0);
      std::vector<int> ncols(ch->n);
      for (int i = 0; i < ch->n; i++) {
        ncols[i] = rtrg_pipeline_columns(pipe, ch->ticket, i);
        off[i + 1] = off[i] + (size_t)ch->cos[i]->n_out * cfg.nk * ncols[i];
      }
      std::vector<int> bad(n_writers, 0);
      std::vector<std::thread> th;
      for (int t = 0; t < n_writers; t++)
        th.emplace_back([&, t]() {
          for (int i = t; i < ch->n; i += n_writers)
            bad[t] += write_model(dirs[ch->begin + i], cfg.nk, ncols[i], ch->cos[i]->n_out, status[i], out + off[i],
                                  hdr + (size_t)i * RTRG_MAX_OUT * 5, hdr0 + 2 * (size_t)i);
        });
      for (auto &t : th) t.join();
      for (int b : bad) failed += b;
      t_write += now_s() - t1;
    }
    rtrg_pipeline_release(pipe, ch->ticket);
    for (rtrg_run_inputs *p : ch->in) rtrg_free_run_inputs(p);
    delete ch;
  };

  for (int c = 0; c < n_chunks; c++) {
    Chunk *ch = nullptr;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return !parsed.empty(); });
      ch = parsed.front();
      parsed.pop_front();
    }
    cv.notify_all();
    if (ch->rc != RTRG_OK) {
      std::fprintf(stderr, "redTimeBatch_b200: cannot read one of the run directories %d-%d\n", ch->begin,
                   ch->begin + ch->n - 1);
      delete ch;
      exit_code = 2;
      break;
    }
    if (rtrg_pipeline_submit(pipe, ch->n, ch->cos.data(), &ch->ticket) != RTRG_OK) {
      std::fprintf(stderr, "redTimeBatch_b200: submit failed\n");
      for (rtrg_run_inputs *p : ch->in) rtrg_free_run_inputs(p);
      delete ch;
      exit_code = 3;
      break;
    }
    in_flight.push_back(ch);
    if (in_flight.size() >= 2) finish_oldest();  // chunk c-1 is written while chunk c runs
  }
  while (!in_flight.empty()) finish_oldest();
  {
    std::lock_guard<std::mutex> lk(mu);
    stop = true;
  }
  cv.notify_all();
  reader.join();
  for (Chunk *ch : parsed) {
    for (rtrg_run_inputs *p : ch->in) rtrg_free_run_inputs(p);
    delete ch;
  }
  rtrg_pipeline_destroy(pipe);
  std::printf("redTimeBatch_b200: %d models, %d failed, %.3f s wall (start-up %.3f s, parse %.3f s, gpu wait %.3f s, "
              "write %.3f s)\n", n, failed, now_s() - t_start, t_startup, t_parse, t_gpu_wait, t_write);
  if (exit_code) return exit_code;
  return failed ? 1 : 0;
}
