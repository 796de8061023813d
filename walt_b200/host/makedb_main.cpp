// makedb_main.cpp -- the `makedb` command of the B200 engine: FASTA -> .dbindex files.
//
// Same surface as the reference's src/walt/makedb.cpp:87-168 (-c/-chrom FASTA file or directory
// of *.fa, -o/-output <name>.dbindex) and the same files out (reference.cpp:302-322,353-379),
// but counting, bucketing and bucket sorting run on the GPU (walt_engine_build_from_sequence).
// Differences that the reference leaves unspecified anyway: N/IUPAC bases are replaced by
// rand() once (the reference re-reads the FASTA and re-draws them for each of the four
// sub-indexes, with srand(time(NULL))), and ties inside a bucket are ordered by position.
#include <dirent.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/walt_b200.h"
#include "../../include/walt_host.h"

namespace {

bool ends_with(const std::string& s, const std::string& x) {
  return s.size() >= x.size() && s.compare(s.size() - x.size(), x.size(), x) == 0;
}

// chromosome files: a directory is scanned for names ending in "fa" in readdir order
// (IdentifyChromosomes, reference.cpp:62-77)
std::vector<std::string> chromosome_files(const std::string& path) {
  struct stat st;
  std::vector<std::string> files;
  if (stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) {
    DIR* d = opendir(path.c_str());
    if (!d) throw std::runtime_error("could not open directory: " + path);
    while (dirent* e = readdir(d))
      if (ends_with(e->d_name, "fa")) files.push_back(path + "/" + e->d_name);
    closedir(d);
  } else {
    files.push_back(path);
  }
  return files;
}

// FASTA -> names (cut at the first blank) and upper-cased, N-replaced sequences (ReadGenome,
// reference.cpp:79-129)
void read_fasta(const std::string& file, std::vector<std::string>& names, std::string& genome,
                std::vector<uint32_t>& lengths) {
  std::ifstream in(file);
  if (!in) throw std::runtime_error("cannot open input file " + file);
  std::string line;
  bool open = false;
  uint64_t cur = 0;
  auto flush = [&]() { if (open) lengths.push_back((uint32_t)cur); cur = 0; };
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '>') {
      flush();
      open = true;
      const size_t cutp = line.find_first_of(" \t");
      names.push_back(line.substr(1, cutp == std::string::npos ? std::string::npos : cutp - 1));
    } else if (open) {
      for (char c : line) {
        c = (char)toupper((unsigned char)c);
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) c = "ACGT"[rand() % 4];
        genome.push_back(c);
      }
      cur += line.size();
    }
  }
  flush();
}

}  // namespace

int main(int argc, const char** argv) {
  srand((unsigned)time(nullptr));
  try {
    std::string chrom, out;
    int device = 0;
    std::vector<std::string> args(argv + 1, argv + argc);
    auto take = [&](const char* l, char s, std::string& dst) {
      for (size_t i = 0; i < args.size();) {
        const std::string& a = args[i];
        if (a == l || (a.size() > 1 && a[0] == '-' && (a.substr(1) == l || (a.size() == 2 && a[1] == s)))) {
          dst = i + 1 < args.size() ? args[i + 1] : "";
          args.erase(args.begin() + i, args.begin() + std::min(args.size(), i + 2));
        } else {
          ++i;
        }
      }
    };
    std::string dev;
    take("chrom", 'c', chrom); take("output", 'o', out); take("device", 'd', dev);
    if (!dev.empty()) device = atoi(dev.c_str());
    if (argc == 1 || chrom.empty() || out.empty() || !args.empty()) {
      fprintf(stderr, "Usage: makedb [OPTIONS]\n\nOptions:\n  -c, -chrom   chromosomes in FASTA file or dir (the suffix of the "
                      "chromosome file should be '.fa') \n  -o, -output  output file name (the suffix of the file should be "
                      "'.dbindex') \n  -d, -device  CUDA device ordinal \n");
      return EXIT_SUCCESS;
    }
    if (!ends_with(out, ".dbindex")) { fprintf(stderr, "The suffix of the output file should be '.dbindex'\n"); return EXIT_FAILURE; }
    if (out.size() > 1000) { fprintf(stderr, "The output file name is too long, please select a shorter name\n"); return EXIT_FAILURE; }

    std::vector<std::string> names;
    std::vector<uint32_t> lengths;
    std::string genome;
    fprintf(stderr, "[READING CHROMOSOMES]\n");
    for (auto& f : chromosome_files(chrom)) read_fasta(f, names, genome, lengths);
    if (names.empty()) throw std::runtime_error("no chromosomes found in " + chrom);
    fprintf(stderr, "[THERE ARE %zu CHROMOSOMES IN THE GENOME]\n[THE TOTAL LENGTH OF ALL CHROMOSOMES IS %zu]\n", names.size(),
            genome.size());

    walt_engine* e = nullptr;
    auto check = [&](int rc) { if (rc) throw std::runtime_error(std::string("walt engine: ") + walt_last_error()); };
    check(walt_engine_create(&e, device));
    std::vector<const char*> np;
    for (auto& n : names) np.push_back(n.c_str());
    check(walt_engine_set_chromosomes(e, (uint32_t)names.size(), lengths.data(), np.data()));
    walt_chroms* chroms = walt_chroms_create((uint32_t)names.size(), np.data(), lengths.data());
    static const char* SFX[4] = {"_CT00", "_CT01", "_GA10", "_GA11"};
    uint32_t size_of_index = 0;
    std::vector<char> seq(genome.size());
    std::vector<uint32_t> counter((1u << 24) + 1u), index;
    for (int which = 0; which < 4; ++which) {
      fprintf(stderr, "[BUILD INDEX %s ON THE GPU]\n", SFX[which]);
      check(walt_engine_build_from_sequence(e, genome.data(), 1u << which));
      uint32_t n = 0;
      check(walt_engine_subindex_info(e, which, &n, nullptr, nullptr));
      index.resize(std::max<uint32_t>(n, 1));
      check(walt_engine_export_subindex(e, which, seq.data(), counter.data(), index.data(), &n));
      fprintf(stderr, "[WRITING INDEX TO %s%s]\n", out.c_str(), SFX[which]);
      if (walt_write_subindex((out + SFX[which]).c_str(), (which & 1) ? '-' : '+', seq.data(), genome.size(), counter.data(),
                              index.data(), n))
        throw std::runtime_error(walt_host_last_error());
      size_of_index = std::max(size_of_index, n);   // makedb.cpp:82-84
    }
    if (walt_write_dbindex_header(out.c_str(), chroms, size_of_index)) throw std::runtime_error(walt_host_last_error());
    walt_chroms_free(chroms);
    walt_engine_destroy(e);
  } catch (const std::runtime_error& e) {
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  } catch (std::bad_alloc&) {
    fprintf(stderr, "ERROR: could not allocate memory\n");
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
