// walt_main.cpp -- the `walt` command of the B200 engine.
//
// Same surface as the reference's src/walt/walt.cpp:88-283 (options -i -r -1 -2 -o -m -N -a -u
// -C -A -b -k -L -sam -v -t, single-dash long names, comma separated file lists, one shared or
// one-per-input output name, truncation of outputs at start, exit codes) plus -P/-pbat and
// -gpus.  The batch drivers (mapping.cpp:421-526, paired.cpp:572-713) become:
//   load index ONCE into HBM -> per part of a batch, three parts under way
//   at a time { load FASTQ (walt_host) | walt_group_map_se/pe on the GPU(s) | write SAM/MR (walt_host) }
//   -> mapstats.
// Mapping has no CPU implementation here: without a CUDA device the program exits 1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <chrono>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/walt_b200.h"
#include "../../include/walt_host.h"

namespace {

// ---- option grammar (smithlab_cpp/OptionParser.cpp:157-192,314-352) --------------------------
struct Opt {
  enum Kind { STR, U32, I32, BOOL } kind;
  std::string long_name;
  char short_name;
  std::string descr;
  bool required;
  void* dst;
  bool specified = false;

  bool matches(const std::string& a) const {
    return a == long_name ||
           (a.size() > 1 && a[0] == '-' && (a.substr(1) == long_name || (a[1] == short_name && a.size() == 2)));
  }
  void assign(const std::string& v) {
    switch (kind) {
      case STR: *(std::string*)dst = v; break;
      case U32: *(uint32_t*)dst = (uint32_t)strtoul(v.c_str(), nullptr, 10); break;
      case I32: *(int*)dst = (int)strtol(v.c_str(), nullptr, 10); break;
      case BOOL: {
        bool& b = *(bool*)dst;
        if (v == "true" || v == "on") b = true; else if (v == "false" || v == "off") b = false; else b = !b;
        break;
      }
    }
  }
  // every occurrence anywhere on the command line is consumed; value options also eat the next token
  bool parse(std::vector<std::string>& args) {
    for (size_t i = 0; i < args.size();) {
      if (!matches(args[i])) { ++i; continue; }
      if (specified && kind != BOOL) throw std::runtime_error("Duplicate assignment top option: " + long_name);
      assign(i + 1 < args.size() ? args[i + 1] : std::string());
      specified = true;
      args.erase(args.begin() + i);
      if (kind != BOOL && i < args.size()) args.erase(args.begin() + i);
    }
    return specified || !required;
  }
  std::string display() const {
    std::ostringstream s;
    if (short_name) s << '-' << short_name << ", -" << long_name; else s << "    -" << long_name;
    return s.str();
  }
};

std::string help_text(const std::string& prog, const std::vector<Opt>& opts) {
  size_t w = 0;
  for (auto& o : opts) w = std::max(w, o.display().size());
  std::ostringstream s;
  s << "Usage: " << prog << " [OPTIONS]\n\nOptions:\n";
  for (size_t i = 2; i < opts.size(); ++i) {
    std::string d = opts[i].display();
    s << "  " << d << std::string(w - d.size(), ' ') << "  " << opts[i].descr << " \n";
  }
  s << "\nHelp options:\n";
  for (size_t i = 0; i < 2; ++i) {
    std::string d = opts[i].display();
    s << "  " << d << std::string(w - d.size(), ' ') << "  " << opts[i].descr << " \n";
  }
  return s.str();
}

std::vector<std::string> split_csv(std::string s) {
  std::replace(s.begin(), s.end(), ',', ' ');
  std::istringstream ss(s);
  std::vector<std::string> v;
  for (std::string t; ss >> t;) v.push_back(t);
  return v;
}

bool fastq_suffix(const std::string& s) {
  auto ends = [&](const char* x) { size_t n = strlen(x); return s.size() >= n && s.compare(s.size() - n, n, x) == 0; };
  // the reference underflows (and aborts) on names shorter than the suffix; here they are just invalid
  return ends(".fastq") || ends(".fq");
}

void require_regular(const std::string& p, const char* what) {
  struct stat st;
  if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) throw std::runtime_error(std::string(what) + p);
}

struct Settings {
  std::string index, se_csv, pe1_csv, pe2_csv, out_csv, adaptor;
  bool sam = false, ambiguous = false, unmapped = false, ag = false, verbose = false, pbat = false;
  uint32_t m = 6, batch = 10000000, b = 5000, top_k = 50, gpus = 1;
  int frag = 1000, threads = 0;   // 0: -t not given, one host thread per hardware thread
};

void engine_check(int rc) {
  if (rc != WALT_OK) throw std::runtime_error(std::string("walt engine: ") + walt_last_error());
}

// the GPUs of a run: one engine per device behind one handle (walt_group, include/walt_b200.h), which cuts
// every batch into contiguous ranges and keeps the results in input order.  (Starting the engines
// beside the first reads' load was measured and dropped: context creation and the pinned / device
// allocations serialise with the loader's page faults on the process's address-space lock, and the index
// residency went from 0.64 s to 0.7 - 1.5 s to hide a first part's load of 0.04 s.)
struct Engines {
  walt_group* g = nullptr;
  ~Engines() { walt_group_destroy(g); }
};

// stage timings on stderr when WALT_TIMING is set (never part of the outputs).  The three stages run
// beside each other on consecutive parts of a batch: each figure is the time its stage was busy.
struct StageClock {
  bool on = getenv("WALT_TIMING") != nullptr;
  double load = 0, map = 0, write = 0;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  static double since(std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
  }
  void report(const char* what, uint64_t n, uint32_t part) const {
    if (on) fprintf(stderr, "[walt timing] %s: %llu reads in parts of %u, stages overlapped: load %.3f s, map %.3f s, "
                            "write %.3f s, total %.3f s, %u host threads\n", what, (unsigned long long)n, part,
                    load, map, write, since(t0), walt_host_threads());
  }
};

// A blocking queue between two stages; close() lets every waiting and later pop() return nullptr.
template <class T>
struct Chan {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<T*> q;
  bool closed = false, drain = false;
  void push(T* v) { { std::lock_guard<std::mutex> lk(mu); q.push_back(v); } cv.notify_one(); }
  T* pop() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&]() { return closed || !q.empty(); });
    if (q.empty() || (closed && !drain)) return nullptr;
    T* v = q.front(); q.pop_front();
    return v;
  }
  // drain_first: what was pushed before is still handed out
  void close(bool drain_first = false) { { std::lock_guard<std::mutex> lk(mu); closed = true; drain = drain_first; } cv.notify_all(); }
};

// One part of a batch on its way through load -> map -> write: the reads of one file (two for pairs) as
// ASCII (for the writers) and in 2-bit form (what crosses PCIe), and the results.
struct Part {
  walt_batch* b[2] = {nullptr, nullptr};
  const uint8_t* packed[2] = {nullptr, nullptr};
  int64_t n[2] = {0, 0};
  bool first = false;            // the first part of a batch
  bool last = false;             // a file ended in this part: nothing follows
  std::string err[2];
  std::vector<walt_best> best;
  std::vector<walt_pe_result> pairs;
  uint32_t n_short[2] = {0, 0};
  explicit Part(int files) { for (int i = 0; i < files; ++i) b[i] = walt_batch_create(); }
  ~Part() { walt_batch_free(b[0]); walt_batch_free(b[1]); }
};

// reads per part: a batch (-N reads: one srand(0) stream, mapping.cpp:73) is loaded, mapped and written in
// parts of this many reads, three parts under way at a time
uint32_t part_reads(const Settings& s) {
  const char* v = getenv("WALT_PART_READS");
  const unsigned long p = v ? strtoul(v, nullptr, 10) : (1ul << 20);
  return (uint32_t)std::max<unsigned long>(1ul, std::min<unsigned long>(p, std::max<uint32_t>(1u, s.batch)));
}

// The batch loop of both modes (mapping.cpp:421-526, paired.cpp:572-713) as three stages on their own threads:
//   load   : the next part of each file (both files of a pair side by side)
//   map    : the calling thread, the engines
//   write  : the part's lines behind everything written so far
// `map` returns false to end the run early (paired-end files of unequal length).
template <class MapFn, class WriteFn>
uint64_t run_parts(const Settings& s, int files, walt_fastq* const* fq, const std::string* adaptor, StageClock& clk,
                   MapFn map, WriteFn write) {
  constexpr int N_PARTS = 4;
  const uint32_t P = part_reads(s);
  std::vector<std::unique_ptr<Part>> parts;
  Chan<Part> free_q, loaded_q, mapped_q;
  for (int i = 0; i < N_PARTS; ++i) { parts.emplace_back(new Part(files)); free_q.push(parts.back().get()); }
  std::string write_err;
  std::atomic<bool> write_failed{false};

  std::thread loader([&]() {
    uint32_t left = s.batch;                       // reads of the current batch still to load
    auto load = [&](Part* p, int f, uint32_t want, bool first) {   // (each file has its own error text: two threads)
      try {
        p->n[f] = walt_fastq_next_part(fq[f], p->b[f], want, adaptor[f].c_str(), first ? 1 : 0);
        if (p->n[f] < 0) p->err[f] = walt_host_last_error();
        if (p->n[f] > 0 && !(p->packed[f] = walt_batch_packed(p->b[f]))) { p->n[f] = -1; p->err[f] = walt_host_last_error(); }
      } catch (const std::exception& e) {   // bad_alloc on the loader's own tables
        p->n[f] = -1;
        p->err[f] = std::string("ERROR: could not load reads: ") + e.what();
      }
    };
    for (;;) {
      Part* p = free_q.pop();
      if (!p) return;
      const auto t = std::chrono::steady_clock::now();
      const uint32_t want = std::min(P, left);
      const bool first = left == s.batch;
      p->err[0].clear(); p->err[1].clear();
      p->first = first;
      // (the order in which two files are read only matters for their rand() streams, and each file has its own)
      std::future<void> other;
      if (files == 2) other = std::async(std::launch::async, load, p, 1, want, first);
      load(p, 0, want, first);
      if (other.valid()) other.get();
      clk.load += StageClock::since(t);
      const bool bad = p->n[0] < 0 || (files == 2 && p->n[1] < 0);
      p->last = bad || p->n[0] < (int64_t)want || (files == 2 && p->n[1] < (int64_t)want);
      left -= (uint32_t)std::max<int64_t>(0, std::min<int64_t>(p->n[0], left));
      if (left == 0) left = s.batch;
      const bool stop = p->last;
      loaded_q.push(p);
      if (stop) return;
    }
  });
  std::thread writer([&]() {
    for (;;) {
      Part* p = mapped_q.pop();
      if (!p) return;
      if (!write_failed) {
        const auto t = std::chrono::steady_clock::now();
        try { write(*p); } catch (const std::exception& e) { write_err = e.what(); write_failed = true; }
        clk.write += StageClock::since(t);
      }
      free_q.push(p);
    }
  });
  auto shut = [&]() {
    free_q.close(); loaded_q.close();
    loader.join();
    mapped_q.close(true);   // the writer first finishes what it was handed
    writer.join();
  };

  uint64_t total = 0;
  try {
    for (;;) {
      Part* p = loaded_q.pop();
      if (!p) break;
      if (p->n[0] < 0) throw std::runtime_error(p->err[0]);
      if (files == 2 && p->n[1] < 0) throw std::runtime_error(p->err[1]);
      // paired.cpp:650-651: when the first file has no further batch the second one is not looked at -- but
      // inside a batch its extra reads are a count that differs (map reports it)
      if (p->n[0] == 0 && (files == 1 || p->first || p->n[1] == 0)) break;
      if (write_failed) break;
      const auto t = std::chrono::steady_clock::now();
      const bool go = map(*p);
      clk.map += StageClock::since(t);
      if (!go) break;
      total += (uint64_t)p->n[0];
      const bool last = p->last;
      mapped_q.push(p);
      if (last) break;
    }
  } catch (...) {
    shut();
    throw;
  }
  shut();
  if (write_failed) throw std::runtime_error(write_err);
  return total;
}

void process_single_end(Engines& eng, const walt_chroms* chroms, const Settings& s, const std::string& reads,
                        const std::string& output) {
  const bool ag = s.ag || s.pbat;   // single-end PBAT reads are A-rich: same as -A
  walt_fastq* fq = walt_fastq_open(reads.c_str());
  if (!fq) throw std::runtime_error("cannot open input file " + reads);
  walt_se_writer* w = walt_se_writer_open(output.c_str(), chroms, ag, s.ambiguous, s.unmapped, s.sam);
  if (!w) { walt_fastq_close(fq); throw std::runtime_error(walt_host_last_error()); }
  if (s.verbose) std::cerr << "input_file: " << reads << std::endl << "output_file: " << output << std::endl;
  StageClock clk;
  uint64_t total = 0;
  try {
    total = run_parts(
        s, 1, &fq, &s.adaptor, clk,
        [&](Part& p) {
          const uint32_t n = (uint32_t)p.n[0];
          p.best.resize(n);
          p.n_short[0] = 0;
          engine_check(walt_group_map_se_packed(eng.g, p.packed[0], walt_batch_offsets(p.b[0]), n, ag ? 1 : 0, s.m, s.b,
                                                p.best.data(), &p.n_short[0]));
          return true;
        },
        [&](Part& p) {
          walt_se_writer_add_short(w, p.n_short[0]);
          if (walt_se_writer_write(w, p.b[0], p.best.data(), (uint32_t)p.n[0])) throw std::runtime_error(walt_host_last_error());
        });
  } catch (...) {
    walt_fastq_close(fq); walt_se_writer_close(w);
    throw;
  }
  walt_fastq_close(fq);
  if (walt_se_writer_close(w)) throw std::runtime_error(walt_host_last_error());
  clk.report("single-end", total, part_reads(s));
}

void process_paired_end(Engines& eng, const walt_chroms* chroms, const Settings& s, const std::string& reads1,
                        const std::string& reads2, const std::string& output) {
  std::string ad[2] = {s.adaptor, s.adaptor};   // extract_adaptors, util.hpp:221-233
  const size_t sep = s.adaptor.find(':');
  if (s.adaptor.rfind(':') != sep) throw std::runtime_error("ERROR: adaptor format \"T_adaptor[:A_adaptor]\"");
  if (sep != std::string::npos) { ad[0] = s.adaptor.substr(0, sep); ad[1] = s.adaptor.substr(sep + 1); }
  if (s.pbat) std::swap(ad[0], ad[1]);   // the T-rich adaptor belongs to the C->T mate, which is mate 2 under PBAT
  walt_fastq* fq[2] = {walt_fastq_open(reads1.c_str()), nullptr};
  if (!fq[0]) throw std::runtime_error("cannot open input file " + reads1);
  fq[1] = walt_fastq_open(reads2.c_str());
  if (!fq[1]) { walt_fastq_close(fq[0]); throw std::runtime_error("cannot open input file " + reads2); }
  walt_pe_writer* w = walt_pe_writer_open(output.c_str(), chroms, s.m, s.top_k, s.frag, s.ambiguous, s.unmapped, s.sam, s.pbat);
  if (!w) { walt_fastq_close(fq[0]); walt_fastq_close(fq[1]); throw std::runtime_error(walt_host_last_error()); }
  fprintf(stderr, "[MAPPING PAIRED-END READS FROM THE FOLLOWING TWO FILES]\n   %s (AND)\n   %s\n", reads1.c_str(), reads2.c_str());
  fprintf(stderr, "[OUTPUT MAPPING RESULTS TO %s]\n", output.c_str());
  bool unequal = false;
  StageClock clk;
  uint64_t total = 0;
  try {
    total = run_parts(
        s, 2, fq, ad, clk,
        [&](Part& p) {
          // paired.cpp:673-677.  (The reference compares the two counts of a whole batch before it maps any of
          // it; here the parts of the batch in front of the one where a file ends have been written by then.)
          if (p.n[0] != p.n[1]) { unequal = true; return false; }
          const uint32_t n = (uint32_t)p.n[0];
          p.pairs.resize(n);
          p.n_short[0] = p.n_short[1] = 0;
          engine_check(walt_group_map_pe_compact_packed(eng.g, p.packed[0], walt_batch_offsets(p.b[0]), p.packed[1],
                                                        walt_batch_offsets(p.b[1]), n, s.m, s.b, s.top_k, s.frag, s.pbat ? 1 : 0,
                                                        p.pairs.data(), &p.n_short[0], &p.n_short[1]));
          return true;
        },
        [&](Part& p) {
          walt_pe_writer_add_short(w, p.n_short[0], p.n_short[1]);
          if (walt_pe_writer_write_compact(w, p.b[0], p.b[1], p.pairs.data(), (uint32_t)p.n[0]))
            throw std::runtime_error(walt_host_last_error());
        });
  } catch (...) {
    walt_fastq_close(fq[0]); walt_fastq_close(fq[1]); walt_pe_writer_close(w);
    throw;
  }
  walt_fastq_close(fq[0]); walt_fastq_close(fq[1]);
  if (unequal) {   // paired.cpp:673-677: exits without writing mapstats
    fprintf(stderr, "The number of reads in paired-end files should be the same.\n");
    exit(EXIT_FAILURE);
  }
  if (walt_pe_writer_close(w)) throw std::runtime_error(walt_host_last_error());
  clk.report("paired-end", total, part_reads(s));
}

}  // namespace

int main(int argc, const char** argv) {
  try {
    Settings s;
    bool help = false, about = false;
    std::string prog = argv[0];
    prog = prog.substr(prog.find_last_of('/') == std::string::npos ? 0 : prog.find_last_of('/') + 1);
    std::vector<Opt> opts = {
        {Opt::BOOL, "help", '?', "print this help message", false, &help},
        {Opt::BOOL, "about", '\0', "print about message", false, &about},
        {Opt::STR, "index", 'i', "index file created by makedb command (the suffix of the index file should be '.dbindex')", true, &s.index},
        {Opt::STR, "reads", 'r', "comma-sep list of read files for singled-end mapping (expect suffix .fastq or .fq)", false, &s.se_csv},
        {Opt::STR, "reads1", '1', "comma-separated list of read files for mate 1 (expect suffix .fastq or .fq)", false, &s.pe1_csv},
        {Opt::STR, "reads2", '2', "comma-separated list of read files for mate 2 (expect suffix .fastq or .fq)", false, &s.pe2_csv},
        {Opt::STR, "output", 'o', "output file names (comma sep)", true, &s.out_csv},
        {Opt::U32, "mismatch", 'm', "max allowed mismatches", false, &s.m},
        {Opt::U32, "number", 'N', "number of reads per batch", false, &s.batch},
        {Opt::BOOL, "ambiguous", 'a', "output one random location for ambiguously mapping reads in separate file", false, &s.ambiguous},
        {Opt::BOOL, "unmapped", 'u', "output unmapped reads in separate file", false, &s.unmapped},
        {Opt::STR, "clip", 'C', "clip the specified adaptor", false, &s.adaptor},
        {Opt::BOOL, "ag-wild", 'A', "map using A/G bisulfite wildcards (single-end)", false, &s.ag},
        {Opt::BOOL, "pbat", 'P', "map post-bisulfite adaptor tagging reads (mate 1 A-rich, mate 2 T-rich)", false, &s.pbat},
        {Opt::U32, "bucket", 'b', "maximum candidates for a seed", false, &s.b},
        {Opt::U32, "topk", 'k', "maximum allowed mappings for a read (paired-end)", false, &s.top_k},
        {Opt::I32, "fraglen", 'L', "max fragment length (paired-end)", false, &s.frag},
        {Opt::BOOL, "sam", '\0', "output sam format", false, &s.sam},
        {Opt::BOOL, "verbose", 'v', "print more run info", false, &s.verbose},
        {Opt::I32, "thread", 't', "number of host threads", false, &s.threads},
        {Opt::U32, "gpus", '\0', "number of GPUs to shard each batch over", false, &s.gpus},
    };
    std::vector<std::string> args(argv + 1, argv + argc);
    std::string missing;
    for (auto& o : opts)
      if (!o.parse(args) && missing.empty()) missing = o.display();
    if (argc == 1 || help) { std::cerr << help_text(prog, opts) << std::endl; return EXIT_SUCCESS; }
    if (about) { std::cerr << "PROGRAM: " << prog << "\nmap Illumina BS-seq reads " << std::endl; return EXIT_SUCCESS; }
    if (!missing.empty()) { std::cerr << "required argument missing: [" << missing << "]" << std::endl; return EXIT_SUCCESS; }
    if (!args.empty()) { std::cerr << help_text(prog, opts) << std::endl; return EXIT_SUCCESS; }

    require_regular(s.index, "bad index file: ");
    for (const char* sfx : {"_CT00", "_CT01", "_GA10", "_GA11"}) require_regular(s.index + sfx, "bad table file: ");

    const std::vector<std::string> se = split_csv(s.se_csv), pe1 = split_csv(s.pe1_csv), pe2 = split_csv(s.pe2_csv);
    for (auto& f : se) if (!fastq_suffix(f)) throw std::runtime_error("read file invalid suffix: " + f);
    if (pe1.size() != pe2.size()) throw std::runtime_error("unequal number of end1 and end2 files");
    for (auto& f : pe1) if (!fastq_suffix(f)) throw std::runtime_error("read file invalid suffix: " + f);
    for (auto& f : pe2) if (!fastq_suffix(f)) throw std::runtime_error("read file invalid suffix: " + f);
    std::vector<std::string> outs = split_csv(s.out_csv);
    if (outs.size() != 1 && outs.size() != se.size() + pe1.size())
      throw std::runtime_error("wrong number of output files: " + s.out_csv);
    if (outs.size() == 1) outs.assign(se.size() + pe1.size(), outs[0]);
    for (auto& o : outs) { std::ofstream a(o); std::ofstream b(o + ".mapstats"); }   // walt.cpp:230-233

    if (s.verbose) std::cerr << "max_mismatches: " << s.m << std::endl << "threads: " << (s.threads > 0 ? (unsigned)s.threads : walt_host_threads()) << std::endl;
    if (s.batch > 100000000u) throw std::runtime_error("batch size may not exceed" + std::to_string(100000000u));
    if (s.top_k < 2 || s.top_k > 300) throw std::runtime_error("paired-end candidates must be in [2, 300]");
    if (s.gpus < 1) s.gpus = 1;
    walt_host_set_threads(s.threads > 0 ? (unsigned)s.threads : 0u);

    walt_chroms* chroms = walt_chroms_read(s.index.c_str());
    if (!chroms) throw std::runtime_error(walt_host_last_error());
    if (s.verbose) {
      uint64_t total = 0;
      for (uint32_t i = 0; i < walt_chroms_count(chroms); ++i) total += walt_chroms_lengths(chroms)[i];
      std::cerr << "n_chroms: " << walt_chroms_count(chroms) << std::endl << "genome_length: " << total << std::endl;
    }

    // index residency: once, for everything this run needs
    uint32_t mask = 0;
    if (!se.empty()) mask |= (s.ag || s.pbat) ? (1u << WALT_GA10 | 1u << WALT_GA11) : (1u << WALT_CT00 | 1u << WALT_CT01);
    if (!pe1.empty()) mask |= 0xFu;
    Engines eng;
    if (mask) {
      const auto t_index = std::chrono::steady_clock::now();
      // one engine (one index replica) per shard.  More shards than devices would put several replicas of
      // the index on one device: the shard count is clamped (WALT_SHARE_DEVICES=1 keeps it, for tests on small indexes)
      const int n_dev = walt_device_count();
      if (n_dev > 0 && s.gpus > (uint32_t)n_dev && !getenv("WALT_SHARE_DEVICES")) {
        std::cerr << "[-gpus " << s.gpus << ": only " << n_dev << " device(s) visible, using " << n_dev << "]" << std::endl;
        s.gpus = (uint32_t)n_dev;
      }
      std::vector<int> devs(s.gpus);
      for (uint32_t i = 0; i < s.gpus; ++i) devs[i] = n_dev > 0 ? (int)(i % (uint32_t)n_dev) : (int)i;
      engine_check(walt_group_create(&eng.g, devs.data(), (int)s.gpus));
      const double t_create = StageClock::since(t_index);
      // the files are read once; the other devices get their replicas device to device (NVLink / NVSwitch)
      engine_check(walt_group_load_dbindex(eng.g, s.index.c_str(), mask));
      if (getenv("WALT_TIMING") && s.gpus > 1)
        fprintf(stderr, "[walt timing] index read once, cloned over NVLink to %u more device%s\n", s.gpus - 1, s.gpus > 2 ? "s" : "");
      if (getenv("WALT_TIMING"))
        fprintf(stderr, "[walt timing] engine start + index residency (%u GPU%s): %.3f s (contexts %.3f s, index %.3f s)\n", s.gpus,
                s.gpus > 1 ? "s" : "", StageClock::since(t_index), t_create, StageClock::since(t_index) - t_create);
    }

    size_t oi = 0;
    if (s.verbose) std::cerr << "n_se_read_files: " << se.size() << std::endl;
    for (auto& f : se) process_single_end(eng, chroms, s, f, outs[oi++]);
    if (s.verbose) std::cerr << "n_pe_read_files: " << pe1.size() << std::endl;
    for (size_t i = 0; i < pe1.size(); ++i) process_paired_end(eng, chroms, s, pe1[i], pe2[i], outs[oi++]);
    walt_chroms_free(chroms);
  } catch (const std::runtime_error& e) {
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  } catch (std::bad_alloc&) {
    std::cerr << "ERROR: could not allocate memory" << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
