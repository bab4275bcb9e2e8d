// main.cpp — `sketchy` command line host over libsketchy_b200.so: the same sub-commands, flags, defaults, stdout rows
// and error texts as the reference CLI (src/cli.rs:23-133, src/main.rs:17-68, src/sketchy.rs), with the hot path on the
// GPU through the C ABI. `info` and the hidden `msh-*` helpers need no GPU.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unistd.h>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sketchy_b200.h"
#include "fastx.hpp"
#include "ingest.hpp"
#include "msh.hpp"

namespace {

struct Args {
  std::string cmd;
  std::map<std::string, std::vector<std::string>> opt;
  bool has(const std::string& k) const { return opt.count(k) != 0; }
  std::string one(const std::string& k, const std::string& dflt = "") const {
    auto it = opt.find(k);
    return it == opt.end() || it->second.empty() ? dflt : it->second[0];
  }
};

const std::map<std::string, std::string> kShort = {
    {"-i", "input"}, {"-o", "output"}, {"-s", "sketch-size|stream"}, {"-k", "kmer-size"}, {"-c", "scale|consensus"},
    {"-e", "seed"}, {"-p", "params"}, {"-r", "reference"}, {"-g", "genotypes"}, {"-q", "query"}, {"-t", "top"},
    {"-l", "limit"}, {"-H", "header"}};

Args parse(int argc, char** argv) {
  Args a;
  if (argc < 2) throw std::runtime_error("usage: sketchy <sketch|info|check|shared|predict> [options]");
  a.cmd = argv[1];
  std::string cur;
  for (int i = 2; i < argc; ++i) {
    std::string t = argv[i];
    if (t.size() >= 2 && t[0] == '-' && !(t.size() > 1 && (isdigit((unsigned char)t[1]) || t[1] == '.'))) {
      std::string name;
      std::string inline_value;
      bool has_inline = false;
      if (t.rfind("--", 0) == 0) {
        name = t.substr(2);
        const size_t eq = name.find('=');  // clap also takes --name=value
        if (eq != std::string::npos) { inline_value = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
      } else {
        auto it = kShort.find(t);
        if (it == kShort.end()) throw std::runtime_error("unknown option " + t);
        name = it->second;
        const size_t bar = name.find('|');
        if (bar != std::string::npos) {  // -s / -c mean different things per sub-command (src/cli.rs:33, 124; :39, 127)
          name = a.cmd == "predict" ? name.substr(bar + 1) : name.substr(0, bar);
        }
      }
      cur = name;
      a.opt[cur];
      if (has_inline) a.opt[cur].push_back(inline_value);
    } else {
      if (cur.empty()) throw std::runtime_error("unexpected argument " + t);
      a.opt[cur].push_back(t);
    }
  }
  return a;
}

std::string basename_of(const std::string& p) {
  const size_t s = p.find_last_of('/');
  return s == std::string::npos ? p : p.substr(s + 1);
}
std::string ext_of(const std::string& p) {
  const std::string b = basename_of(p);
  const size_t d = b.find_last_of('.');
  return d == std::string::npos ? "" : b.substr(d + 1);
}

void require_msh(const std::string& path) {
  const std::string e = ext_of(path);
  if (e == "fsh") throw std::runtime_error("Finch (.fsh) scaled sketches are not supported by the B200 build (DESIGN.md §7)");
  if (e != "msh") throw std::runtime_error("reference sketch file must have Mash (.msh) or Finch (.fsh) extension");
}

// One process per GPU. A launcher (torchrun, mpirun, a shell loop) starts `world` copies with the usual environment:
// rank / world size / local rank from SKETCHY_B200_RANK / _WORLD / _DEVICE, else RANK / WORLD_SIZE / LOCAL_RANK (torchrun),
// else OMPI_COMM_WORLD_*; the NCCL id travels through the file SKETCHY_B200_COMM_FILE (rank 0 writes it, the others wait
// for it). Every rank reads the same inputs; rank 0 prints / writes the output.
int env_int(std::initializer_list<const char*> names, int dflt) {
  for (const char* n : names)
    if (const char* v = getenv(n)) return atoi(v);
  return dflt;
}

struct Ctx {
  skb_ctx* c = nullptr;
  int rank = 0, world = 1;
  // with_comm = false: a sub-command whose ranks exchange nothing on the device (`sketch`) skips the NCCL set-up
  static void rank_from_env(int& rank, int& world) {
    rank = env_int({"SKETCHY_B200_RANK", "RANK", "OMPI_COMM_WORLD_RANK"}, 0);
    world = env_int({"SKETCHY_B200_WORLD", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE"}, 1);
    if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("bad rank / world size in the environment");
  }
  explicit Ctx(bool with_comm = true) {
    rank_from_env(rank, world);
    int dev = env_int({"SKETCHY_B200_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"}, 0);
    // One process per GPU: a rank that can see every GPU of the box initialises all of them, and the ranks' start-ups
    // queue behind each other in the driver (two ranks: 1.1 s until the context is ready, against 0.5 s when each sees
    // only its own GPU — profiles/r02_sketch_cli_trace.txt). So unless the caller has chosen the visible devices, the
    // rank binds itself to its GPU before the first CUDA call. SKETCHY_B200_NO_DEVICE_BINDING=1 leaves the process alone.
    if (world > 1 && !getenv("CUDA_VISIBLE_DEVICES") && !getenv("SKETCHY_B200_NO_DEVICE_BINDING")) {
      setenv("CUDA_VISIBLE_DEVICES", std::to_string(dev).c_str(), 1);
      dev = 0;
    }
    const int rc = skb_create(dev, &c);
    if (rc != SKB_OK) throw std::runtime_error("no B200 (sm_100) device: the B200 build has no CPU fallback");
    if (world > 1 && with_comm) join();
  }
  ~Ctx() { if (c) { if (skb_comm_world(c) > 1) skb_comm_destroy(c); skb_destroy(c); } }
  void check(int rc) const { if (rc != SKB_OK) throw std::runtime_error(skb_last_error(c)); }
  void join() {
    // stdout carries the result rows: whatever NCCL prints while the communicator is set up (its version banner goes to
    // stdout under NCCL_DEBUG=VERSION) is sent to stderr instead
    fflush(stdout);
    const int saved_stdout = dup(1);
    dup2(2, 1);
    struct Restore { int fd; ~Restore() { fflush(stdout); dup2(fd, 1); close(fd); } } restore{saved_stdout};
    const char* path = getenv("SKETCHY_B200_COMM_FILE");
    if (!path) throw std::runtime_error("multi-GPU run: set SKETCHY_B200_COMM_FILE to a path every rank can reach");
    uint8_t id[SKB_COMM_ID_BYTES];
    if (rank == 0) {
      if (skb_comm_unique_id(id) != SKB_OK) throw std::runtime_error("NCCL is not available (libnccl.so.2)");
      const std::string tmp = std::string(path) + ".tmp";
      FILE* fp = fopen(tmp.c_str(), "wb");
      if (!fp || fwrite(id, 1, sizeof id, fp) != sizeof id) throw std::runtime_error("cannot write the communicator file");
      fclose(fp);
      if (rename(tmp.c_str(), path) != 0) throw std::runtime_error("cannot publish the communicator file");
    } else {
      for (int waited = 0;; ++waited) {
        FILE* fp = fopen(path, "rb");
        if (fp) {
          const size_t n = fread(id, 1, sizeof id, fp);
          fclose(fp);
          if (n == sizeof id) break;
        }
        if (waited > 6000) throw std::runtime_error("timed out waiting for the communicator file");
        usleep(10000);
      }
    }
    check(skb_comm_init(c, id, rank, world));
  }
  // Host-side gather without a communicator: every rank but 0 publishes its bytes as <SKETCHY_B200_COMM_FILE>.<tag>.<rank>
  // (written under a temporary name, then renamed), rank 0 collects them in rank order (and removes the files).
  void gather_files(const char* tag, const std::vector<uint8_t>& mine, std::vector<uint8_t>& all) const {
    const char* path = getenv("SKETCHY_B200_COMM_FILE");
    if (!path) throw std::runtime_error("multi-GPU run: set SKETCHY_B200_COMM_FILE to a path every rank can reach");
    auto name = [&](int r) { return std::string(path) + "." + tag + "." + std::to_string(r); };
    if (rank != 0) {
      const std::string tmp = name(rank) + ".tmp";
      FILE* fp = fopen(tmp.c_str(), "wb");
      if (!fp || (mine.size() && fwrite(mine.data(), 1, mine.size(), fp) != mine.size())) throw std::runtime_error("cannot write the rank's result file");
      fclose(fp);
      if (rename(tmp.c_str(), name(rank).c_str()) != 0) throw std::runtime_error("cannot publish the rank's result file");
      return;
    }
    all.assign(mine.size() * (size_t)world, 0);
    if (!mine.empty()) memcpy(all.data(), mine.data(), mine.size());
    for (int r = 1; r < world; ++r) {
      for (int waited = 0;; ++waited) {
        FILE* fp = fopen(name(r).c_str(), "rb");
        if (fp) {
          const size_t n = mine.empty() ? 0 : fread(all.data() + (size_t)r * mine.size(), 1, mine.size(), fp);
          fclose(fp);
          if (n == mine.size()) { remove(name(r).c_str()); break; }
        }
        if (waited > 360000) throw std::runtime_error("timed out waiting for a rank's result file");
        usleep(10000);
      }
    }
  }
  // contiguous share of n items of this rank
  void range(uint64_t n, uint64_t& begin, uint64_t& count) const { skb_dist_range(n, rank, world, &begin, &count); }
};

// ---- genotype table (src/sketchy.rs:538-571): TSV with header; header minus the first column; name -> columns
struct Genotypes {
  std::string header;
  std::vector<std::vector<std::string>> rows;
  std::unordered_map<std::string, std::vector<std::string>> map;
};
std::vector<std::string> split_tab(const std::string& l) {
  std::vector<std::string> f;
  size_t s = 0;
  for (;;) {
    const size_t e = l.find('\t', s);
    f.push_back(l.substr(s, e == std::string::npos ? std::string::npos : e - s));
    if (e == std::string::npos) break;
    s = e + 1;
  }
  return f;
}
Genotypes read_genotypes(const std::string& path) {
  FILE* fp = fopen(path.c_str(), "r");
  if (!fp) throw std::runtime_error("failed to open genotype file or record with CSV");
  Genotypes g;
  std::string line;
  char buf[1 << 16];
  bool first = true;
  std::string acc;
  while (fgets(buf, sizeof buf, fp)) {
    acc += buf;
    if (acc.empty() || acc.back() != '\n') continue;
    while (!acc.empty() && (acc.back() == '\n' || acc.back() == '\r')) acc.pop_back();
    if (!acc.empty()) {
      std::vector<std::string> f = split_tab(acc);
      if (first) {
        for (size_t i = 1; i < f.size(); ++i) g.header += (i > 1 ? "\t" : "") + f[i];
        first = false;
      } else {
        g.map[f[0]] = std::vector<std::string>(f.begin() + 1, f.end());  // duplicate names: last wins
        g.rows.push_back(std::move(f));
      }
    }
    acc.clear();
  }
  if (!acc.empty()) {
    std::vector<std::string> f = split_tab(acc);
    if (!first) { g.map[f[0]] = std::vector<std::string>(f.begin() + 1, f.end()); g.rows.push_back(std::move(f)); }
  }
  fclose(fp);
  return g;
}

std::string join_tab(const std::vector<std::string>& v) {
  std::string s;
  for (size_t i = 0; i < v.size(); ++i) s += (i ? "\t" : "") + v[i];
  return s;
}

// src/sketchy.rs:358-413; consensus ties go to the value met first in rank order (the reference's HashMap order is
// nondeterministic there — DESIGN.md §2). The parts of a row that depend on the reference sketch alone (its name, its
// genotype columns joined, the columns as small integers for the consensus vote) are made once; a row is then an
// append of the read number, two ready strings and the count (streaming predict prints reads x top rows).
struct RowTable {
  std::vector<std::string> name_part, geno_part;        // "\t" + name + "\t",  "\t" + genotype columns + "\n"
  size_t n_features = 0;
  std::vector<uint32_t> value_id;                       // [sketch][feature]: index into values[feature]
  std::vector<std::vector<std::string>> values;         // distinct values of a feature
  RowTable(const msh::File& ref, const Genotypes& g) {
    const size_t N = ref.sketches.size();
    name_part.reserve(N); geno_part.reserve(N);
    for (size_t i = 0; i < N; ++i) n_features = std::max(n_features, g.map.at(ref.sketches[i].name).size());
    values.resize(n_features);
    value_id.assign(N * n_features, kNoValue);  // a row with fewer columns casts no vote for the missing ones
    n_cols.reserve(N);
    std::vector<std::unordered_map<std::string, uint32_t>> seen(n_features);
    for (size_t i = 0; i < N; ++i) {
      const std::string& name = ref.sketches[i].name;
      const std::vector<std::string>& cols = g.map.at(name);
      name_part.push_back("\t" + name + "\t");
      geno_part.push_back("\t" + join_tab(cols) + "\n");
      n_cols.push_back((uint32_t)cols.size());
      for (size_t j = 0; j < cols.size(); ++j) {
        auto it = seen[j].find(cols[j]);
        if (it == seen[j].end()) { it = seen[j].emplace(cols[j], (uint32_t)values[j].size()).first; values[j].push_back(cols[j]); }
        value_id[i * n_features + j] = it->second;
      }
    }
  }
  static constexpr uint32_t kNoValue = 0xFFFFFFFFu;
  std::vector<uint32_t> n_cols;                         // genotype columns of a sketch's row
  static void put(std::string& out, uint64_t v) {
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) out.push_back(buf[--n]);
  }
  // the rows of one read, appended to `out`
  void rows(std::string& out, uint64_t read, const uint32_t* idx, const uint64_t* sum, uint32_t top, bool consensus) const {
    if (consensus) {
      put(out, read);
      out += "\t-\t-\t";
      const size_t nf = n_cols[idx[0]];  // the best row's column count decides the width (src/sketchy.rs:368)
      for (size_t j = 0; j < nf; ++j) {
        uint32_t best = 0;
        size_t best_n = 0;
        for (uint32_t t = 0; t < top; ++t) {
          const uint32_t v = value_id[(size_t)idx[t] * n_features + j];
          if (v == kNoValue) continue;
          size_t c = 0;
          for (uint32_t u = 0; u < top; ++u) c += value_id[(size_t)idx[u] * n_features + j] == v;
          if (c > best_n) { best_n = c; best = v; }
        }
        if (j) out.push_back('\t');
        if (best_n) out += values[j][best];
      }
      out.push_back('\n');
    } else {
      for (uint32_t t = 0; t < top; ++t) {
        put(out, read);
        out += name_part[idx[t]];
        put(out, sum[t]);
        out += geno_part[idx[t]];
      }
    }
  }
};

// this rank's contiguous range of the reference sketches (all of them on one GPU); indices reported by the library are global
void upload_reference(const Ctx& c, const msh::File& ref) {
  uint64_t lo = 0, cnt = 0;
  c.range(ref.sketches.size(), lo, cnt);
  std::vector<uint64_t> flat, off{0};
  for (uint64_t i = lo; i < lo + cnt; ++i) {
    const auto& s = ref.sketches[i];
    flat.insert(flat.end(), s.hashes.begin(), s.hashes.end());
    off.push_back(flat.size());
  }
  c.check(skb_ref_upload(c.c, flat.data(), off.data(), (uint32_t)cnt, (uint32_t)lo));
}

// A finished single-GPU command leaves without tearing the CUDA context down piece by piece (freeing the device and
// page-locked buffers one by one and the runtime's exit handlers cost a short run more than its kernels): the
// results are flushed, the process ends, the driver reclaims everything.
[[noreturn]] void leave(int rc) {
  fflush(nullptr);
  _exit(rc);
}

// ---- sub-commands -----------------------------------------------------------------------------------------------
constexpr uint64_t kWindowBytes = 128ull << 20;  // files of one skb_sketch call (on disk); the next window is read meanwhile

int cmd_sketch(const Args& a) {
  if (!a.has("output")) throw std::runtime_error("error: The following required arguments were not provided: --output <output>");
  const std::string out = a.one("output");
  require_msh(out);
  const uint32_t s = (uint32_t)std::stoul(a.one("sketch-size", "1000"));
  const uint32_t k = (uint32_t)std::stoul(a.one("kmer-size", "16"));
  const uint64_t seed = std::stoull(a.one("seed", "0"));
  const double scale = std::stod(a.one("scale", "0.001"));
  if (!(scale >= 0.0 && scale <= 1.0)) throw std::runtime_error("Scale parameter must be between 0 and 1");
  std::vector<std::string> files;
  if (a.has("input")) files = a.opt.at("input");
  else for (std::string l; std::getline(std::cin, l);) if (!l.empty()) files.push_back(l);  // src/sketchy.rs:137-146
  const auto t_main = std::chrono::steady_clock::now();
  int rank = 0, world = 1;
  Ctx::rank_from_env(rank, world);
  if (rank == 0) {
    FILE* fp = fopen(out.c_str(), "wb");  // created before sketching, like the reference (:153)
    if (!fp) throw std::runtime_error("failed to open file");
    fclose(fp);
  }
  // The reference runs its files on a rayon pool, one sketcher per file (src/sketchy.rs:470-473). Here: the files are
  // partitioned over the GPUs by contiguous range (no collective in the hashing); a rank takes its files in windows
  // (<= 128 MB on disk each): window i+1 is read into memory (plain files as they are, compressed ones decoded) and
  // split into record slices on the host threads while window i is packed from those slices, copied and sketched, one
  // skb_sketch call per window; results are kept in file order.
  // the ranks of one box share its host cores
  const unsigned host_threads = std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, world));
  const uint32_t G = (uint32_t)files.size();
  uint64_t f_lo = 0, f_cnt = 0;
  skb_dist_range(G, rank, world, &f_lo, &f_cnt);
  const size_t per_file = 24 + (size_t)s * 12;  // n, bases, kmers, hashes[s], counts[s]
  uint64_t share = 0, dummy = 0;
  skb_dist_range(G, 0, world, &dummy, &share);  // the largest share: the fixed record size of the exchange
  std::vector<uint8_t> mine((size_t)share * per_file, 0);
  auto rec = [&](std::vector<uint8_t>& buf, size_t i) { return buf.data() + i * per_file; };
  // Three stages, each on its own thread, joined by bounded queues: (1) the loader reads window i+2 into memory (its
  // own pool of threads); (2) the packer normalises and 2-bit packs window i+1 from the record slices into one of two
  // batches and starts its copy to the device; (3) this thread calls skb_sketch on window i and files the results.
  // Three windows of files and two batches circulate; nothing is allocated once they have been round once.
  struct Loaded { ingest::Files* f; size_t g0, g1; };
  struct Packed { skb_batch* b; size_t g0, g1; double pack_ms; };
  std::vector<std::pair<size_t, size_t>> wins;
  const size_t g_end = f_lo + f_cnt;
  uint64_t window_bytes = kWindowBytes;
  if (const char* e = getenv("SKETCHY_B200_WINDOW_BYTES")) window_bytes = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
  for (size_t g0 = f_lo; g0 < g_end;) {
    const size_t g1 = std::min(g_end, ingest::window_end(files, g0, window_bytes));
    wins.emplace_back(g0, g1);
    g0 = g1;
  }
  ingest::Files pool[3];
  ingest::Channel<ingest::Files*> free_files(3);
  for (auto& f : pool) free_files.push(&f);
  ingest::Channel<Loaded> loaded(2);
  ingest::Channel<Packed> packed(1);
  ingest::Channel<skb_batch*> free_batches(2);
  std::mutex err_m;
  std::exception_ptr err;
  auto failed = [&]() {  // a stage failed: the first error is kept, every queue lets go of its waiters
    { std::lock_guard<std::mutex> l(err_m); if (!err) err = std::current_exception(); }
    free_files.abort(); loaded.abort(); packed.abort(); free_batches.abort();
  };
  const bool trace = getenv("SKB_TRACE_SKETCH") != nullptr;
  auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  // the first windows are read while the CUDA context comes up (the larger part of a short run's wall time)
  std::thread loader([&]() {
    try {
      for (const auto& w : wins) {
        ingest::Files* f = nullptr;
        if (!free_files.pop(f)) return;
        ingest::load_files(files, w.first, w.second, host_threads, *f);
        if (!loaded.push({f, w.first, w.second})) return;
      }
      loaded.close();
    } catch (...) { failed(); }
  });
  struct Join { std::thread& t; std::function<void()> before; ~Join() { if (t.joinable()) { if (before) before(); t.join(); } } };
  Join join_loader{loader, [&]() { free_files.abort(); loaded.abort(); }};  // also when something below throws
  Ctx c(false);  // the ranks sketch their files independently: no communicator
  if (trace) fprintf(stderr, "[sketch rank %d] context ready %.1f ms after main\n", rank, ms_since(t_main));
  skb_batch* batches[2] = {nullptr, nullptr};
  for (auto& b : batches) { c.check(skb_batch_create(c.c, &b)); free_batches.push(b); }
  std::thread packer([&]() {
    try {
      Loaded L;
      while (loaded.pop(L)) {
        skb_batch* b = nullptr;
        if (!free_batches.pop(b)) return;
        const auto t0 = std::chrono::steady_clock::now();
        c.check(skb_batch_clear(b));
        if (L.f->n()) c.check(skb_batch_add_records(b, L.f->rec.data(), L.f->len.data(), L.f->grp.data(), L.f->n(), host_threads));
        if (skb_batch_num_groups(b)) c.check(skb_batch_stage(b));
        const double pack_ms = ms_since(t0);
        if (!free_files.push(L.f)) return;  // the slices have been read: the window's memory goes back to the loader
        if (!packed.push({b, L.g0, L.g1, pack_ms})) return;
      }
      packed.close();
    } catch (...) { failed(); }
  });
  Join join_packer{packer, [&]() { loaded.abort(); packed.abort(); free_batches.abort(); free_files.abort(); }};
  try {
    std::vector<uint64_t> hs, bases, kmers;
    std::vector<uint32_t> cnt, n;
    Packed P;
    auto t_wait = std::chrono::steady_clock::now();
    while (packed.pop(P)) {
      const double wait_ms = ms_since(t_wait);
      const auto t0 = std::chrono::steady_clock::now();
      const uint32_t W = (uint32_t)(P.g1 - P.g0);
      const uint32_t have = skb_batch_num_groups(P.b);  // trailing empty files have no group
      hs.resize((size_t)W * s); cnt.resize((size_t)W * s);
      bases.assign(W, 0); kmers.assign(W, 0); n.assign(W, 0);
      if (have) c.check(skb_sketch(c.c, P.b, k, s, seed, hs.data(), cnt.data(), n.data(), bases.data(), kmers.data()));
      for (uint32_t g = 0; g < W; ++g) {
        uint8_t* r = rec(mine, P.g0 - f_lo + g);
        const uint64_t nn = n[g];
        memcpy(r, &nn, 8); memcpy(r + 8, &bases[g], 8); memcpy(r + 16, &kmers[g], 8);
        memcpy(r + 24, &hs[(size_t)g * s], (size_t)nn * 8);
        memcpy(r + 24 + (size_t)s * 8, &cnt[(size_t)g * s], (size_t)nn * 4);
      }
      if (trace)  // SKB_TRACE_SKETCH=1: where a window's time goes (stderr; stdout carries nothing for `sketch`)
        fprintf(stderr, "[sketch rank %d] files %zu-%zu: waited %.2f ms for the packed window (its pack + copy start took %.2f ms), sketch call %.2f ms\n",
                rank, P.g0, P.g1, wait_ms, P.pack_ms, ms_since(t0));
      free_batches.push(P.b);
      t_wait = std::chrono::steady_clock::now();
    }
  } catch (...) { failed(); }
  loader.join();
  packer.join();
  if (err) std::rethrow_exception(err);
  // (the batches are not destroyed: a dozen frees of page-locked and device buffers each, a second on a slow box, and the
  // process leaves as soon as its results are out)
  if (getenv("SKB_TRACE_SKETCH"))
    fprintf(stderr, "[sketch rank %d] all windows done %.1f ms after main\n", rank,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_main).count());
  std::vector<uint8_t> all;
  if (c.world > 1) c.gather_files("sketch", mine, all);
  if (c.rank != 0) leave(0);
  msh::File f;
  f.kmer_size = k; f.sketch_size = 0; f.hash_seed = seed;  // minHashesPerWindow = the largest sketch, as finch writes it [RECALLED]
  for (uint32_t g = 0; g < G; ++g) {
    const int owner = share ? (int)(g / share) : 0;
    const uint8_t* r = c.world > 1 ? all.data() + (size_t)owner * mine.size() + (size_t)(g - owner * share) * per_file
                                   : rec(mine, g);
    uint64_t nn;
    msh::Sketch sk;
    sk.name = basename_of(files[g]);
    memcpy(&nn, r, 8); memcpy(&sk.seq_length, r + 8, 8); memcpy(&sk.num_valid_kmers, r + 16, 8);
    sk.hashes.resize(nn); sk.counts.resize(nn);
    memcpy(sk.hashes.data(), r + 24, nn * 8);
    memcpy(sk.counts.data(), r + 24 + (size_t)s * 8, nn * 4);
    f.sketch_size = std::max<uint32_t>(f.sketch_size, (uint32_t)sk.hashes.size());
    f.sketches.push_back(std::move(sk));
  }
  msh::write_file(out, f);
  if (getenv("SKB_TRACE_SKETCH"))
    fprintf(stderr, "[sketch rank %d] .msh written %.1f ms after main\n", rank,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_main).count());
  leave(0);
}

int cmd_info(const Args& a) {
  const std::string in = a.one("input");
  require_msh(in);
  const msh::File f = msh::read_file(in);
  if (f.sketches.empty()) throw std::runtime_error("sketch file holds no sketches");
  if (a.has("params")) {
    // the reference derives sketch_size from the first sketch's length (src/sketchy.rs:520-527, :195)
    printf("type=mash sketch_size=%zu kmer_size=%u seed=%llu\n", f.sketches[0].hashes.size(), f.kmer_size,
           (unsigned long long)f.hash_seed);
  } else {
    for (const auto& s : f.sketches) {
      // finch::statistics::cardinality [RECALLED]: (len - 1) / (max_hash / usize::MAX) in f32
      unsigned long long card = 0;
      if (!s.hashes.empty()) {
        const float frac = (float)s.hashes.back() / (float)UINT64_MAX;
        card = (unsigned long long)((float)(s.hashes.size() - 1) / frac);
      }
      printf("%s %llu %llu\n", s.name.c_str(), (unsigned long long)s.seq_length, card);
    }
  }
  return 0;
}

int cmd_check(const Args& a) {
  const msh::File f = msh::read_file(a.one("reference"));
  const Genotypes g = read_genotypes(a.one("genotypes"));
  // the reference builds an InvalidIdentifier error for mismatching names but discards it (src/sketchy.rs:219-228):
  // only the sizes are enforced
  if (f.sketches.size() != g.rows.size()) throw std::runtime_error("reference sketch and genotype table must have the same length");
  puts("ok");
  return 0;
}

int cmd_shared(const Args& a) {
  require_msh(a.one("reference"));
  require_msh(a.one("query"));
  const msh::File ref = msh::read_file(a.one("reference")), qry = msh::read_file(a.one("query"));
  if (ref.kmer_size != qry.kmer_size)
    throw std::runtime_error("reference (" + ref.sketches[0].name + ") k-mer size (" + std::to_string(ref.kmer_size) +
                             ") does not match query (" + qry.sketches[0].name + ") k-mer size (" + std::to_string(qry.kmer_size) + ")");
  Ctx c;
  upload_reference(c, ref);
  std::vector<uint64_t> flat, off{0};
  for (const auto& s : qry.sketches) { flat.insert(flat.end(), s.hashes.begin(), s.hashes.end()); off.push_back(flat.size()); }
  const uint32_t N = (uint32_t)ref.sketches.size(), Q = (uint32_t)qry.sketches.size();
  std::vector<uint64_t> out((size_t)N * Q);
  c.check(skb_shared_counts(c.c, flat.data(), off.data(), Q, out.data()));
  for (uint32_t i = 0; i < N; ++i)
    for (uint32_t j = 0; j < Q; ++j)
      printf("%s %s %llu\n", ref.sketches[i].name.c_str(), qry.sketches[j].name.c_str(), (unsigned long long)out[(size_t)i * Q + j]);
  return 0;
}

int cmd_predict(const Args& a) {
  const uint32_t top = (uint32_t)std::stoul(a.one("top", "1"));
  const uint64_t limit = std::stoull(a.one("limit", "0"));
  const bool stream = a.has("stream"), consensus = a.has("consensus"), header = a.has("header");
  if (consensus && top % 2 != 1) throw std::runtime_error("--top must be an odd number when using --consensus");
  // checked before anything is printed (the reference has no such limit). The limit is the streaming kernels'; the
  // read-set mode ranks once, and beyond the limit it ranks its count vector on the host.
  if (top < 1 || (stream && top > SKB_MAX_TOP))
    throw std::runtime_error("--top must be between 1 and " + std::to_string(SKB_MAX_TOP) + " for streaming predict in the B200 build");
  require_msh(a.one("reference"));
  const msh::File ref = msh::read_file(a.one("reference"));
  if (ref.sketches.empty()) throw std::runtime_error("reference sketch file holds no sketches");
  const uint32_t k = ref.kmer_size;
  const uint32_t s_query = (uint32_t)std::max<size_t>(1, ref.sketches[0].hashes.size());  // src/sketchy.rs:82, 522
  const uint64_t seed = ref.hash_seed;
  Ctx c;
  if (c.world > 1 && !a.has("input")) throw std::runtime_error("a multi-GPU run reads its reads from a file (-i): stdin cannot be shared by the ranks");
  const bool speaker = c.rank == 0;  // rank 0 prints; every rank computes
  ingest::ChunkReader reads(a.has("input") ? a.one("input") : std::string("-"));
  const Genotypes g = read_genotypes(a.one("genotypes"));
  for (const auto& s : ref.sketches)
    if (!g.map.count(s.name)) throw std::runtime_error("reference sketch identifier " + s.name + " has no genotype row");
  if (header && speaker) printf("reads\tsketch_id\tshared_hashes\t%s\n", g.header.c_str());
  if (top > ref.sketches.size()) throw std::runtime_error("--top exceeds the number of reference sketches");
  upload_reference(c, ref);
  const RowTable table(ref, g);
  const unsigned host_threads = std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, c.world));
  size_t kChunkReads = 65536;
  if (const char* e = getenv("SKETCHY_B200_CHUNK_READS")) kChunkReads = std::max<size_t>(1, strtoull(e, nullptr, 10));  // (tests: many small chunks)
  constexpr uint64_t kChunkBytes = 64ull << 20;  // sequence bytes of a chunk (its buffer: about twice that for FASTQ)
  if (stream) {  // src/sketchy.rs:317-356
    // Four stages on their own threads, joined by bounded queues: (1) the reader decodes the input into a chunk's buffer
    // and splits it into record slices; (2) the packer normalises and 2-bit packs this rank's share of the chunk into
    // one of two batches and starts its copy; (3) this thread runs the (collective) predict call of the chunk;
    // (4) the printer formats and writes the rows of the chunk before. Three chunks and two batches circulate.
    struct Packed { skb_batch* b; uint64_t n; };
    struct Result { std::vector<uint32_t> idx; std::vector<uint64_t> sum; uint64_t first, n; };
    ingest::Chunk chunk_pool[3];
    ingest::Channel<ingest::Chunk*> free_chunks(3), loaded(2);
    for (auto& ch : chunk_pool) free_chunks.push(&ch);
    ingest::Channel<Packed> packed(1);
    ingest::Channel<skb_batch*> free_batches(2);
    ingest::Channel<Result> results(2);
    std::mutex err_m;
    std::exception_ptr err;
    auto failed = [&]() {
      { std::lock_guard<std::mutex> l(err_m); if (!err) err = std::current_exception(); }
      free_chunks.abort(); loaded.abort(); packed.abort(); free_batches.abort(); results.abort();
    };
    skb_batch* batches[2] = {nullptr, nullptr};
    for (auto& b : batches) { c.check(skb_batch_create(c.c, &b)); free_batches.push(b); }
    std::thread reader([&]() {
      try {
        uint64_t fed = 0;
        for (;;) {
          ingest::Chunk* ch = nullptr;
          if (!free_chunks.pop(ch)) return;
          const size_t want = limit ? (size_t)std::min<uint64_t>(kChunkReads, limit - fed) : kChunkReads;  // -l: stop feeding reads (:350-353)
          // a live stream that is pausing: predict what has arrived instead of waiting for a full chunk (one rank only:
          // the ranks of a multi-GPU run must cut the same chunks)
          if (want == 0 || !reads.next(*ch, want, kChunkBytes, c.world == 1)) break;
          fed += ch->n();
          if (!loaded.push(ch)) return;
        }
        loaded.close();
      } catch (...) {
        // a bad record: the chunks in front of it run through the stages and are printed, as the reference prints every
        // read it has handled before it stops; the error is reported when the pipeline has drained
        { std::lock_guard<std::mutex> l(err_m); if (!err) err = std::current_exception(); }
        loaded.close();
      }
    });
    std::thread packer([&]() {
      try {
        ingest::Chunk* ch = nullptr;
        while (loaded.pop(ch)) {
          skb_batch* b = nullptr;
          if (!free_batches.pop(b)) return;
          // every rank has parsed the chunk; it packs, copies and hashes only its share of it
          uint64_t lo = 0, cnt = 0;
          c.range(ch->n(), lo, cnt);
          c.check(skb_batch_clear(b));
          if (cnt) {
            c.check(skb_batch_add_records(b, ch->rec.data() + lo, ch->len.data() + lo, nullptr, cnt, host_threads));
            c.check(skb_batch_stage(b));
          }
          const uint64_t n = ch->n();
          if (!free_chunks.push(ch)) return;
          if (!packed.push({b, n})) return;
        }
        packed.close();
      } catch (...) { failed(); }
    });
    std::thread printer([&]() {
      try {
        Result r;
        std::string out;
        while (results.pop(r)) {
          out.clear();
          for (uint64_t i = 0; i < r.n; ++i) table.rows(out, r.first + i, &r.idx[i * top], &r.sum[i * top], top, consensus);
          if (fwrite(out.data(), 1, out.size(), stdout) != out.size()) throw std::runtime_error("failed to write to stdout");
          fflush(stdout);  // the reference's println! reaches a pipe line by line: a consumer of the live stream sees each chunk at once
        }
      } catch (...) { failed(); }
    });
    try {
      uint64_t read = 1;
      Packed P;
      while (packed.pop(P)) {
        Result r;
        r.first = read; r.n = P.n;
        if (speaker) { r.idx.resize(P.n * top); r.sum.resize(P.n * top); }
        c.check(skb_predict_stream_dist(c.c, P.b, P.n, k, s_query, seed, top, speaker ? r.idx.data() : nullptr, speaker ? r.sum.data() : nullptr));
        free_batches.push(P.b);
        read += P.n;
        if (speaker && !results.push(std::move(r))) break;
      }
      results.close();
    } catch (...) { failed(); }
    reader.join(); packer.join(); printer.join();
    if (err) std::rethrow_exception(err);
    if (c.world == 1) leave(0);  // (without freeing the batches one buffer at a time)
    for (auto& b : batches) skb_batch_destroy(b);
    return 0;
  } else {  // src/sketchy.rs:281-315: one sketcher for all reads
    skb_batch* b = nullptr;
    c.check(skb_batch_create(c.c, &b));
    uint64_t read = 0;
    bool any = false;
    ingest::Chunk ch;
    std::vector<uint32_t> zeros;
    for (;;) {
      const size_t want = limit ? (size_t)std::min<uint64_t>(1u << 20, limit - read) : (size_t)1 << 20;
      if (want == 0 || !reads.next(ch, want, 256ull << 20, false)) break;
      zeros.assign(ch.n(), 0);  // every read joins group 0
      c.check(skb_batch_add_records(b, ch.rec.data(), ch.len.data(), zeros.data(), ch.n(), host_threads));
      any = true;
      read += ch.n();
    }
    std::vector<uint64_t> q(s_query);
    uint32_t qn = 0;
    uint64_t bases = 0, kmers = 0;
    if (any) c.check(skb_sketch(c.c, b, k, s_query, seed, q.data(), nullptr, &qn, &bases, &kmers));
    const uint64_t qoff[2] = {0, qn};
    // shared hashes against this rank's rows, its best `top` of them, then the ranks' lists merged by (count desc, index asc)
    uint64_t lo = 0, cnt = 0;
    c.range(ref.sketches.size(), lo, cnt);
    std::vector<uint64_t> counts(std::max<uint64_t>(cnt, 1));
    if (cnt) c.check(skb_shared_counts(c.c, q.data(), qoff, 1, counts.data()));
    const uint32_t ltop = (uint32_t)std::min<uint64_t>(top, cnt);
    std::vector<uint32_t> idx(top, 0xFFFFFFFFu);
    std::vector<uint64_t> sum(top, 0);
    if (ltop && ltop <= SKB_MAX_TOP) {
      c.check(skb_rank_counts(c.c, counts.data(), (uint32_t)cnt, ltop, idx.data(), sum.data()));
    } else if (ltop) {  // a longer list than the device ranks: the same order (count desc, index asc) on the host
      std::vector<uint32_t> order(cnt);
      for (uint32_t i = 0; i < cnt; ++i) order[i] = i;
      std::partial_sort(order.begin(), order.begin() + ltop, order.end(),
                        [&](uint32_t x, uint32_t y) { return counts[x] != counts[y] ? counts[x] > counts[y] : x < y; });
      for (uint32_t t = 0; t < ltop; ++t) { idx[t] = order[t]; sum[t] = counts[order[t]]; }
    }
    for (uint32_t t = 0; t < ltop; ++t) idx[t] += (uint32_t)lo;
    if (c.world > 1) {
      std::vector<uint8_t> mine((size_t)top * 12), all((size_t)top * 12 * c.world);
      memcpy(mine.data(), sum.data(), (size_t)top * 8);
      memcpy(mine.data() + (size_t)top * 8, idx.data(), (size_t)top * 4);
      c.check(skb_comm_allgather_host(c.c, mine.data(), all.data(), mine.size()));
      std::vector<std::pair<uint64_t, uint32_t>> cand;
      for (int w = 0; w < c.world; ++w)
        for (uint32_t t = 0; t < top; ++t) {
          uint64_t sv; uint32_t iv;
          memcpy(&sv, all.data() + (size_t)w * mine.size() + (size_t)t * 8, 8);
          memcpy(&iv, all.data() + (size_t)w * mine.size() + (size_t)top * 8 + (size_t)t * 4, 4);
          if (iv != 0xFFFFFFFFu) cand.push_back({sv, iv});
        }
      std::sort(cand.begin(), cand.end(), [](const auto& x, const auto& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
      for (uint32_t t = 0; t < top; ++t) { sum[t] = cand[t].first; idx[t] = cand[t].second; }
    }
    if (speaker) {
      std::string out;
      table.rows(out, read, idx.data(), sum.data(), top, consensus);
      fwrite(out.data(), 1, out.size(), stdout);
    }
    skb_batch_destroy(b);
  }
  return 0;
}

// hidden helpers for the CPU tests of the .msh codec (no GPU needed): text <-> msh
//   line 1: k seed sketch_size n ; per sketch: "name<TAB>comment<TAB>seq_length<TAB>num_valid_kmers", a line of hashes,
//   a line of counts
int cmd_msh_from_text(const std::string& in, const std::string& out) {
  std::ifstream is(in);
  if (!is) throw std::runtime_error("failed to open file");
  msh::File f;
  size_t n = 0;
  std::string line;
  std::getline(is, line);
  { std::istringstream h(line); h >> f.kmer_size >> f.hash_seed >> f.sketch_size >> n; }
  for (size_t i = 0; i < n; ++i) {
    msh::Sketch s;
    std::getline(is, line);
    const std::vector<std::string> fl = split_tab(line);
    if (fl.size() < 4) throw std::runtime_error("bad sketch line");
    s.name = fl[0]; s.comment = fl[1]; s.seq_length = std::stoull(fl[2]); s.num_valid_kmers = std::stoull(fl[3]);
    std::getline(is, line);
    { std::istringstream hs(line); for (uint64_t v; hs >> v;) s.hashes.push_back(v); }
    std::getline(is, line);
    { std::istringstream cs(line); for (uint32_t v; cs >> v;) s.counts.push_back(v); }
    f.sketches.push_back(std::move(s));
  }
  msh::write_file(out, f);
  return 0;
}

int cmd_msh_to_text(const std::string& in) {
  const msh::File f = msh::read_file(in);
  printf("%u %llu %u %zu\n", f.kmer_size, (unsigned long long)f.hash_seed, f.sketch_size, f.sketches.size());
  for (const auto& s : f.sketches) {
    printf("%s\t%s\t%llu\t%llu\n", s.name.c_str(), s.comment.c_str(), (unsigned long long)s.seq_length,
           (unsigned long long)s.num_valid_kmers);
    for (size_t i = 0; i < s.hashes.size(); ++i) printf(i ? " %llu" : "%llu", (unsigned long long)s.hashes[i]);
    printf("\n");
    for (size_t i = 0; i < s.counts.size(); ++i) printf(i ? " %u" : "%u", s.counts[i]);
    printf("\n");
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    if (argc == 4 && std::string(argv[1]) == "msh-from-text") return cmd_msh_from_text(argv[2], argv[3]);
    if (argc == 3 && std::string(argv[1]) == "msh-to-text") return cmd_msh_to_text(argv[2]);
    const Args a = parse(argc, argv);
    if (a.cmd == "sketch") return cmd_sketch(a);
    if (a.cmd == "info") return cmd_info(a);
    if (a.cmd == "check") return cmd_check(a);
    if (a.cmd == "shared") return cmd_shared(a);
    if (a.cmd == "predict") return cmd_predict(a);
    throw std::runtime_error("unknown sub-command " + a.cmd);
  } catch (const std::exception& e) {
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
}
