// walt_main.cpp -- the `walt` command of the B200 engine.
//
// Same surface as the reference's src/walt/walt.cpp:88-283 (options -i -r -1 -2 -o -m -N -a -u
// -C -A -b -k -L -sam -v -t, single-dash long names, comma separated file lists, one shared or
// one-per-input output name, truncation of outputs at start, exit codes) plus -P/-pbat and
// -gpus.  The batch drivers (mapping.cpp:421-526, paired.cpp:572-713) become:
//   load index ONCE into HBM -> per batch { load FASTQ (walt_host) -> walt_engine_map_se/pe on
//   the GPU(s) -> write SAM/MR (walt_host) } -> mapstats.
// Mapping has no CPU implementation here: without a CUDA device the program exits 1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <chrono>
#include <future>
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
// every batch into contiguous ranges and keeps the results in input order
struct Engines {
  walt_group* g = nullptr;
  ~Engines() { walt_group_destroy(g); }
};

// stage timings on stderr when WALT_TIMING is set (never part of the outputs)
struct StageClock {
  bool on = getenv("WALT_TIMING") != nullptr;
  double load = 0, map = 0, write = 0;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  static double since(std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
  }
  void report(const char* what, uint64_t n) const {
    if (on) fprintf(stderr, "[walt timing] %s: %llu reads, load %.3f s (overlapped with map+write after the first batch), map %.3f s, "
                            "write %.3f s, total %.3f s, %u host threads\n", what, (unsigned long long)n, load, map, write,
                    since(t0), walt_host_threads());
  }
};

// one loaded batch: the ASCII reads (for the writers) and their 2-bit form (what crosses PCIe)
struct Loaded { int64_t n = 0; const uint8_t* packed = nullptr; std::string err; double seconds = 0; };
Loaded load_batch(walt_fastq* fq, walt_batch* b, uint32_t max_reads, const std::string& adaptor) {
  Loaded l;
  const auto t = std::chrono::steady_clock::now();
  l.n = walt_fastq_next_batch(fq, b, max_reads, adaptor.c_str());
  if (l.n < 0) l.err = walt_host_last_error();
  if (l.n > 0 && !(l.packed = walt_batch_packed(b))) { l.n = -1; l.err = walt_host_last_error(); }
  l.seconds = StageClock::since(t);
  return l;
}

void map_se_batch(Engines& eng, const walt_batch* b, const uint8_t* packed, const Settings& s, bool ag,
                  std::vector<walt_best>& res, uint32_t& n_short) {
  const uint32_t n = walt_batch_size(b);
  res.resize(n);
  n_short = 0;
  engine_check(walt_group_map_se_packed(eng.g, packed, walt_batch_offsets(b), n, ag ? 1 : 0, s.m, s.b, res.data(), &n_short));
}

void process_single_end(Engines& eng, const walt_chroms* chroms, const Settings& s, const std::string& reads,
                        const std::string& output) {
  const bool ag = s.ag || s.pbat;   // single-end PBAT reads are A-rich: same as -A
  walt_fastq* fq = walt_fastq_open(reads.c_str());
  if (!fq) throw std::runtime_error("cannot open input file " + reads);
  walt_se_writer* w = walt_se_writer_open(output.c_str(), chroms, ag, s.ambiguous, s.unmapped, s.sam);
  if (!w) { walt_fastq_close(fq); throw std::runtime_error(walt_host_last_error()); }
  if (s.verbose) std::cerr << "input_file: " << reads << std::endl << "output_file: " << output << std::endl;
  // two batches: the next one is loaded (and packed) while the current one is mapped and written
  walt_batch* bb[2] = {walt_batch_create(), walt_batch_create()};
  std::vector<walt_best> res;
  StageClock clk;
  uint64_t total = 0;
  std::future<Loaded> next;
  try {
    int k = 0;
    Loaded cur = load_batch(fq, bb[0], s.batch, s.adaptor);
    clk.load += cur.seconds;
    for (;;) {
      if (cur.n < 0) throw std::runtime_error(cur.err);
      if (cur.n == 0) break;
      const bool last = (uint32_t)cur.n < s.batch;
      if (!last) next = std::async(std::launch::async, load_batch, fq, bb[1 - k], s.batch, s.adaptor);
      uint32_t n_short = 0;
      auto t = std::chrono::steady_clock::now();
      map_se_batch(eng, bb[k], cur.packed, s, ag, res, n_short);
      clk.map += StageClock::since(t);
      t = std::chrono::steady_clock::now();
      walt_se_writer_add_short(w, n_short);
      if (walt_se_writer_write(w, bb[k], res.data(), (uint32_t)cur.n)) throw std::runtime_error(walt_host_last_error());
      clk.write += StageClock::since(t);
      total += (uint64_t)cur.n;
      if (last) break;
      cur = next.get();
      clk.load += cur.seconds;
      k = 1 - k;
    }
  } catch (...) {
    if (next.valid()) next.wait();
    walt_batch_free(bb[0]); walt_batch_free(bb[1]); walt_fastq_close(fq); walt_se_writer_close(w);
    throw;
  }
  walt_batch_free(bb[0]); walt_batch_free(bb[1]);
  walt_fastq_close(fq);
  if (walt_se_writer_close(w)) throw std::runtime_error(walt_host_last_error());
  clk.report("single-end", total);
}

void process_paired_end(Engines& eng, const walt_chroms* chroms, const Settings& s, const std::string& reads1,
                        const std::string& reads2, const std::string& output) {
  std::string ad1 = s.adaptor, ad2 = s.adaptor;   // extract_adaptors, util.hpp:221-233
  const size_t sep = s.adaptor.find(':');
  if (s.adaptor.rfind(':') != sep) throw std::runtime_error("ERROR: adaptor format \"T_adaptor[:A_adaptor]\"");
  if (sep != std::string::npos) { ad1 = s.adaptor.substr(0, sep); ad2 = s.adaptor.substr(sep + 1); }
  if (s.pbat) std::swap(ad1, ad2);   // the T-rich adaptor belongs to the C->T mate, which is mate 2 under PBAT
  walt_fastq* f1 = walt_fastq_open(reads1.c_str());
  if (!f1) throw std::runtime_error("cannot open input file " + reads1);
  walt_fastq* f2 = walt_fastq_open(reads2.c_str());
  if (!f2) { walt_fastq_close(f1); throw std::runtime_error("cannot open input file " + reads2); }
  walt_pe_writer* w = walt_pe_writer_open(output.c_str(), chroms, s.m, s.top_k, s.frag, s.ambiguous, s.unmapped, s.sam, s.pbat);
  if (!w) { walt_fastq_close(f1); walt_fastq_close(f2); throw std::runtime_error(walt_host_last_error()); }
  fprintf(stderr, "[MAPPING PAIRED-END READS FROM THE FOLLOWING TWO FILES]\n   %s (AND)\n   %s\n", reads1.c_str(), reads2.c_str());
  fprintf(stderr, "[OUTPUT MAPPING RESULTS TO %s]\n", output.c_str());
  // two pairs of batches: both mate files of the next batch are loaded concurrently while the
  // current batch is mapped and written
  walt_batch* b1[2] = {walt_batch_create(), walt_batch_create()};
  walt_batch* b2[2] = {walt_batch_create(), walt_batch_create()};
  std::vector<walt_pe_result> res;
  bool unequal = false;
  StageClock clk;
  uint64_t total = 0;
  std::future<Loaded> next1, next2;
  auto free_all = [&]() {
    for (int i = 0; i < 2; ++i) { walt_batch_free(b1[i]); walt_batch_free(b2[i]); }
    walt_fastq_close(f1); walt_fastq_close(f2);
  };
  try {
    int k = 0;
    // (the order in which the two files are read only matters for the rand() stream, which every
    // load restarts from srand(0), mapping.cpp:73)
    next1 = std::async(std::launch::async, load_batch, f1, b1[0], s.batch, ad1);
    Loaded c2 = load_batch(f2, b2[0], s.batch, ad2);
    Loaded c1 = next1.get();
    clk.load += std::max(c1.seconds, c2.seconds);
    for (;;) {
      if (c1.n < 0) throw std::runtime_error(c1.err);
      if (c1.n == 0) break;   // paired.cpp:650-651: the second file is not even read when the first one is exhausted
      if (c2.n < 0) throw std::runtime_error(c2.err);
      if (c1.n != c2.n) { unequal = true; break; }
      const uint32_t n = (uint32_t)c1.n;
      const bool last = n < s.batch;
      if (!last) {
        next1 = std::async(std::launch::async, load_batch, f1, b1[1 - k], s.batch, ad1);
        next2 = std::async(std::launch::async, load_batch, f2, b2[1 - k], s.batch, ad2);
      }
      auto t = std::chrono::steady_clock::now();
      res.resize(n);
      uint32_t short1 = 0, short2 = 0;
      engine_check(walt_group_map_pe_compact_packed(eng.g, c1.packed, walt_batch_offsets(b1[k]), c2.packed, walt_batch_offsets(b2[k]),
                                                    n, s.m, s.b, s.top_k, s.frag, s.pbat ? 1 : 0, res.data(), &short1, &short2));
      walt_pe_writer_add_short(w, short1, short2);
      clk.map += StageClock::since(t);
      t = std::chrono::steady_clock::now();
      if (walt_pe_writer_write_compact(w, b1[k], b2[k], res.data(), n))
        throw std::runtime_error(walt_host_last_error());
      clk.write += StageClock::since(t);
      total += n;
      if (last) break;
      c1 = next1.get(); c2 = next2.get();
      clk.load += std::max(c1.seconds, c2.seconds);
      k = 1 - k;
    }
  } catch (...) {
    if (next1.valid()) next1.wait();
    if (next2.valid()) next2.wait();
    free_all(); walt_pe_writer_close(w);
    throw;
  }
  free_all();
  if (unequal) {   // paired.cpp:673-677: exits without writing mapstats
    fprintf(stderr, "The number of reads in paired-end files should be the same.\n");
    exit(EXIT_FAILURE);
  }
  if (walt_pe_writer_close(w)) throw std::runtime_error(walt_host_last_error());
  clk.report("paired-end", total);
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
    const auto t_index = std::chrono::steady_clock::now();
    if (mask) {
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
      // the files are read once; the other devices get their replicas device to device (NVLink / NVSwitch)
      engine_check(walt_group_load_dbindex(eng.g, s.index.c_str(), mask));
      if (getenv("WALT_TIMING") && s.gpus > 1)
        fprintf(stderr, "[walt timing] index read once, cloned over NVLink to %u more device%s\n", s.gpus - 1, s.gpus > 2 ? "s" : "");
      if (getenv("WALT_TIMING"))
        fprintf(stderr, "[walt timing] engine start + index residency (%u GPU%s): %.3f s\n", s.gpus, s.gpus > 1 ? "s" : "",
                StageClock::since(t_index));
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
