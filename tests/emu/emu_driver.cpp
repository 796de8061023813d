// tests/emu/emu_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Steps the lane-level kernel source (walt_b200/csrc/walt_core.cuh) on the CPU: every
// emulated warp is 32 cooperative fibers that run in lock step between warp primitives
// (ballot / shfl / reduce / sync), with hand-rolled x86-64 context switches.  This lets the
// CPU-only test suite ("-m 'not gpu'") check the kernel logic -- table construction, taint
// handling, k-ary search, ordered folds, heap pushes, pairing -- against the oracle without a
// device.  It is NOT a product code path: libwaltb200.so contains none of this and fails
// loudly without a GPU; nothing under walt_b200/ loads this library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../walt_b200/csrc/walt_core.cuh"
#include "../../walt_b200/csrc/walt_stdsort.cuh"

using namespace waltcore;

// ------------------------------------------------------------------------------------------
// fibers
// ------------------------------------------------------------------------------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;

struct WarpEmu;
struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  WarpEmu* warp = nullptr;
  uint32_t lane = 0;
  bool done = false;
};

typedef void (*LaneFn)(WarpEmu*, uint32_t lane, void* arg);

struct WarpEmu {
  Fiber fib[32];
  void* main_sp = nullptr;
  uint32_t slots[2][32];
  uint32_t tags[2][32];
  uint32_t cur = 0;       // running lane
  uint32_t nf = 32;       // lanes in the emulated group (8, 16 or 32)
  uint32_t n_done = 0;
  LaneFn fn = nullptr;
  void* arg = nullptr;
  bool diverged = false;

  WarpEmu() {
    for (int i = 0; i < 32; ++i) {
      fib[i].stack = (char*)aligned_alloc(64, STACK_BYTES);
      fib[i].warp = this;
      fib[i].lane = i;
    }
  }
  ~WarpEmu() { for (int i = 0; i < 32; ++i) free(fib[i].stack); }

  static void trampoline();
  void yield_next() {
    uint32_t from = cur;
    uint32_t to = (cur + 1) % nf;
    cur = to;
    emu_switch(&fib[from].sp, fib[to].sp);
  }
  void run(LaneFn f, void* a) {
    fn = f; arg = a; n_done = 0; cur = 0;
    for (uint32_t i = 0; i < nf; ++i) {
      fib[i].done = false;
      // initial frame: 6 callee-saved registers + return address (trampoline)
      void** top = (void**)(fib[i].stack + STACK_BYTES);
      top -= 1;                      // after `ret` rsp == 8 (mod 16), as at a normal function entry
      *--top = (void*)&WarpEmu::trampoline;
      for (int r = 0; r < 6; ++r) *--top = nullptr;
      fib[i].sp = top;
    }
    emu_switch(&main_sp, fib[0].sp);
  }
};

thread_local WarpEmu* g_current = nullptr;

void WarpEmu::trampoline() {
  WarpEmu* w = g_current;
  uint32_t lane = w->cur;
  w->fn(w, lane, w->arg);
  w->fib[lane].done = true;
  w->n_done++;
  if (w->n_done == w->nf) {
    void* dummy;
    emu_switch(&dummy, w->main_sp);
  } else {
    // lanes finish in order; the next one is parked at its last primitive
    uint32_t to = (lane + 1) % w->nf;
    w->cur = to;
    void* dummy;
    emu_switch(&dummy, w->fib[to].sp);
  }
  abort();  // never resumed
}

// the group policy handed to walt_core.cuh: WD lanes own one read
template <uint32_t WD>
struct EmuWarp {
  static constexpr uint32_t WIDTH = WD;
  WarpEmu* w;
  uint32_t my_lane;
  uint32_t n_prim = 0;
  uint32_t lane() const { return my_lane; }
  // every lane deposits v, one full round later all WD values are visible
  const uint32_t* exchange(uint32_t v, uint32_t tag) {
    uint32_t buf = n_prim & 1u;
    ++n_prim;
    w->slots[buf][my_lane] = v;
    w->tags[buf][my_lane] = tag;
    w->yield_next();
    for (uint32_t i = 0; i < WD; ++i)
      if (w->tags[buf][i] != tag) w->diverged = true;
    return w->slots[buf];
  }
  uint32_t ballot(bool p) {
    const uint32_t* s = exchange(p ? 1u : 0u, 1u);
    uint32_t m = 0;
    for (uint32_t i = 0; i < WD; ++i) m |= (s[i] & 1u) << i;
    return m;
  }
  uint32_t shfl(uint32_t v, int src) { return exchange(v, 2u)[(uint32_t)src % WD]; }
  uint32_t reduce_add(uint32_t v) {
    const uint32_t* s = exchange(v, 3u);
    uint32_t a = 0;
    for (uint32_t i = 0; i < WD; ++i) a += s[i];
    return a;
  }
  uint32_t reduce_min(uint32_t v) {
    const uint32_t* s = exchange(v, 4u);
    uint32_t a = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < WD; ++i) a = std::min(a, s[i]);
    return a;
  }
  void sync() { exchange(0u, 5u); }
};

// ------------------------------------------------------------------------------------------
// host copy of the device-side index structures, built with the same per-element functions
// the GPU kernels use
// ------------------------------------------------------------------------------------------
struct EmuSubIndex {
  std::vector<uint64_t> genome;
  std::vector<uint32_t> index, table, taint_bits, taint_rank, taint_start, taint_key, taint_pos, taint_len, taint_slot;
  std::vector<Entry> entries;
  uint32_t depth = 0, ag = 0;
  uint32_t unsorted = 0;
  SubIndexView view() const {
    SubIndexView v;
    v.genome = genome.data(); v.entries = entries.data(); v.table = table.data();
    v.taint_bits = taint_bits.data(); v.taint_rank = taint_rank.data(); v.taint_start = taint_start.data();
    v.taint_pos = taint_pos.data(); v.taint_len = taint_len.data(); v.taint_slot = taint_slot.data();
    v.n_taint = (uint32_t)taint_key.size(); v.index_size = (uint32_t)index.size();
    v.depth = depth; v.ag = ag;
    return v;
  }
};

struct EmuEngine {
  std::vector<uint32_t> starts;
  uint32_t n_chr = 0, genome_len = 0;
  Pow3 p3;
  EmuSubIndex sub[4];
  ChromView cv() const { ChromView c; c.starts = starts.data(); c.n_chr = n_chr; c.genome_len = genome_len; return c; }
};

uint32_t choose_depth(uint32_t index_size) {
  uint32_t d = KEY_WEIGHT;
  uint64_t p = 531441;
  while (d < MAX_DEPTH && p < index_size) { p *= 3; ++d; }
  return d;
}

}  // namespace

extern "C" {

void* emu_engine_create(uint32_t n_chr, const uint32_t* lengths) {
  EmuEngine* e = new EmuEngine;
  e->n_chr = n_chr;
  e->starts.resize(n_chr + 1);
  e->starts[0] = 0;
  for (uint32_t i = 0; i < n_chr; ++i) e->starts[i + 1] = e->starts[i] + lengths[i];
  e->genome_len = e->starts[n_chr];
  uint32_t p = 1;
  for (uint32_t i = 0; i <= MAX_DEPTH; ++i) { e->p3.v[i] = p; p *= 3u; }
  return e;
}
void emu_engine_destroy(void* h) { delete (EmuEngine*)h; }

// returns 0 ok, 1 bad letter, 2 index not in makedb order
int emu_engine_load_subindex(void* h, int which, const char* seq, const uint32_t* index,
                             uint32_t index_size, int force_depth) {
  EmuEngine* e = (EmuEngine*)h;
  EmuSubIndex& s = e->sub[which];
  s.ag = which >= 2;
  const uint64_t L = e->genome_len;
  const uint64_t words = (L + PAD_BASES + 31) / 32 + TAIL_PAD_WORDS;
  s.genome.assign(words, 0);
  for (uint64_t i = 0; i < L; ++i) {
    uint32_t c = (uint8_t)seq[i];
    if (!ascii_is_acgt(c)) return 1;
    uint32_t code = ascii_code(c);
    if (s.ag ? code == 2u : code == 1u) return 1;
    uint64_t p = i + PAD_BASES;
    s.genome[p >> 5] |= (uint64_t)code << (62 - 2 * (p & 31));
  }
  s.index.assign(index, index + index_size);
  s.index.resize(index_size + 32, 0);  // readable pad like the device buffer
  s.index.resize(index_size);
  s.depth = force_depth > 0 ? (uint32_t)force_depth : choose_depth(index_size);
  const uint32_t n_keys = e->p3.v[s.depth];
  s.table.assign((size_t)n_keys + 1, 0);
  s.entries.assign((size_t)index_size + 64, Entry{0u, 0u});
  // table fill == build_table_kernel
  ChromView cv = e->cv();
  uint32_t prev_key = 0;
  s.unsorted = 0;
  for (uint64_t i = 0; i <= index_size; ++i) {
    uint32_t key;
    if (i < index_size) {
      uint32_t en = s.index[i];
      uint32_t chr = chrom_of(cv.starts, cv.n_chr, en);
      key = entry_table_key(s.genome.data(), en, cv.starts[chr + 1], s.depth, s.ag, e->p3);
      s.entries[i].pos = en;
      s.entries[i].fp = entry_fingerprint(s.genome.data(), en, cv.starts[chr + 1], s.depth, s.ag, e->p3);
    } else {
      key = n_keys;
    }
    uint32_t from = i ? prev_key + 1 : 0;
    if (i && i < index_size && key < prev_key) s.unsorted++;
    for (uint64_t k = from; k <= key; ++k) s.table[k] = (uint32_t)i;
    if (i < index_size) prev_key = std::max(prev_key, key);
  }
  // taint list == taint_kernel + host sort
  std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> t;
  for (uint32_t c = 0; c < e->n_chr; ++c) {
    uint32_t cs = e->starts[c], ce = e->starts[c + 1];
    for (uint32_t d = MIN_SEED_LEN + 1; d <= TAINT_SPAN; ++d) {
      if (ce - cs < d) break;
      uint32_t pos = ce - d;
      t.push_back({entry_key12(s.genome.data(), pos, s.ag, e->p3), {pos, d}});
    }
  }
  std::sort(t.begin(), t.end());
  s.taint_key.clear(); s.taint_pos.clear(); s.taint_len.clear();
  for (auto& x : t) {
    s.taint_key.push_back(x.first); s.taint_pos.push_back(x.second.first); s.taint_len.push_back(x.second.second);
  }
  build_taint_directory(s.taint_key, s.taint_bits, s.taint_rank, s.taint_start);
  // slots of the tainted positions == what table_keys_kernel records on the device
  s.taint_slot.assign(s.taint_pos.size() + 1, 0xFFFFFFFFu);
  for (uint32_t i = 0; i < index_size; ++i) {
    const uint32_t en = s.index[i];
    const uint32_t chr = chrom_of(cv.starts, cv.n_chr, en);
    if (cv.starts[chr + 1] - en > TAINT_SPAN) continue;
    for (size_t t = 0; t < s.taint_pos.size(); ++t)
      if (s.taint_pos[t] == en) s.taint_slot[t] = i;
  }
  return s.unsorted ? 2 : 0;
}

uint32_t emu_engine_depth(void* h, int which) { return ((EmuEngine*)h)->sub[which].depth; }

struct emu_best { uint32_t genome_pos, times, mismatch; char strand; char pad[3]; };
struct emu_cand { uint32_t genome_pos, mismatch; char strand; char pad[3]; };
struct emu_pair { uint32_t best_times; int32_t best_i, best_j, frag_len; };

}  // extern "C"

namespace {

struct SeJob {
  const EmuEngine* e; const char* seqs; const uint64_t* offs; uint32_t lo, hi;
  int ag; uint32_t m, b; int literal; emu_best* out; uint32_t max_len;
  uint64_t* scratch; uint32_t* cached_len; Counters* ctr; int* bad;
  int packed;   // seqs is the 2-bit form of walt_pack_reads (include/walt_host.h)
};

template <uint32_t WD>
void se_lane(WarpEmu* w, uint32_t lane, void* arg) {
  SeJob* j = (SeJob*)arg;
  EmuWarp<WD> W{w, lane};
  SubIndexView ix2[2] = {j->e->sub[j->ag ? 2 : 0].view(), j->e->sub[j->ag ? 3 : 1].view()};
  ChromView cv = j->e->cv();
  MapConfig cfg; cfg.b = j->b; cfg.literal_all = j->literal;
  const uint32_t nwmax = (j->max_len + 31) / 32;
  ReadScratch sc = carve_scratch(j->scratch, nwmax ? nwmax : 1);
  uint32_t cached = *j->cached_len;  // lane-private copy, kept uniform
  Counters ctr{0, 0, 0};
  for (uint32_t r = j->lo; r < j->hi; ++r) {
    BestState st;
    uint32_t len = (uint32_t)(j->offs[r + 1] - j->offs[r]);
    const char* seq = j->packed ? j->seqs + (j->offs[r] >> 2) + r : j->seqs + j->offs[r];
    bool ok = MAP_OK == (j->packed
                  ? map_read_se<EmuWarp<WD>, true>(W, ix2, cv, j->e->p3, cfg, seq, len, j->ag != 0, j->m, sc, cached, st, ctr)
                  : map_read_se<EmuWarp<WD>, false>(W, ix2, cv, j->e->p3, cfg, seq, len, j->ag != 0, j->m, sc, cached, st, ctr));
    if (lane == 0) {
      if (!ok) *j->bad = 1;
      j->out[r].genome_pos = st.pos; j->out[r].times = st.times; j->out[r].mismatch = st.mm;
      j->out[r].strand = (char)st.strand; memset(j->out[r].pad, 0, 3);
    }
  }
  {  // counters are lane-local: sum them over the group
    uint32_t a = W.reduce_add(ctr.lookups), b = W.reduce_add(ctr.candidates), c = W.reduce_add(ctr.literal);
    if (lane == 0) { j->ctr->lookups = a; j->ctr->candidates = b; j->ctr->literal = c; }
  }
}

struct PeJob {
  const EmuEngine* e; const char* seqs; const uint64_t* offs; uint32_t lo, hi;
  int ag; uint32_t m, b, top_k; int literal; emu_cand* ranked; uint32_t* n_ranked; uint32_t max_len;
  uint64_t* scratch; HeapEntry* heap; int* bad;
  int logged; HeapEntry* log; uint32_t* hist;   // two-phase form: LogSink + replay_heap_log
};

template <uint32_t WD>
void pe_lane(WarpEmu* w, uint32_t lane, void* arg) {
  PeJob* j = (PeJob*)arg;
  EmuWarp<WD> W{w, lane};
  SubIndexView ix2[2] = {j->e->sub[j->ag ? 2 : 0].view(), j->e->sub[j->ag ? 3 : 1].view()};
  ChromView cv = j->e->cv();
  MapConfig cfg; cfg.b = j->b; cfg.literal_all = j->literal;
  const uint32_t nwmax = (j->max_len + 31) / 32;
  ReadScratch sc = carve_scratch(j->scratch, nwmax ? nwmax : 1);
  uint32_t cached = 0;
  Counters ctr{0, 0, 0};
  for (uint32_t r = j->lo; r < j->hi; ++r) {
    uint32_t len = (uint32_t)(j->offs[r + 1] - j->offs[r]);
    uint32_t hsize = 0;
    bool ok;
    if (j->logged) {
      uint32_t n_log = 0;
      ok = MAP_OK == map_read_pe_logged(W, ix2, cv, j->e->p3, cfg, j->seqs + j->offs[r], len, j->ag != 0, j->m,
                                        j->top_k, sc, cached, j->log, j->hist, n_log, ctr);
      W.sync();
      if (lane == 0) replay_heap_log(j->log, n_log, j->top_k, j->heap, hsize);   // second kernel: one thread per read
    } else {
      ok = MAP_OK == map_read_pe(W, ix2, cv, j->e->p3, cfg, j->seqs + j->offs[r], len, j->ag != 0, j->m,
                                 j->top_k, sc, cached, j->heap, hsize, ctr);
    }
    if (lane == 0) {
      if (!ok) *j->bad = 1;
      // drain, paired.cpp:684-692
      uint32_t c = 0, sz = hsize;
      while (sz) {
        HeapEntry t = heap_pop(j->heap, sz);
        emu_cand& o = j->ranked[(size_t)r * j->top_k + c++];
        o.genome_pos = t.pos; o.mismatch = he_mm(t); o.strand = (t.mm_strand & 0x10000u) ? '-' : '+';
        memset(o.pad, 0, 3);
      }
      j->n_ranked[r] = c;
    }
    W.sync();
  }
}

struct PairJob {
  const EmuEngine* e; const emu_cand* r1; const uint32_t* n1; const uint64_t* offs1;
  const emu_cand* r2; const uint32_t* n2; const uint64_t* offs2; uint32_t lo, hi, top_k, m; int frag_range;
  emu_pair* out; uint32_t* sm;
};
struct GetEmuCand {
  const emu_cand* p;
  RankedCand operator()(uint32_t i) const { RankedCand c; c.pos = p[i].genome_pos; c.mm = p[i].mismatch; c.strand = (uint8_t)p[i].strand; return c; }
};
void pair_wide_lane(WarpEmu* w, uint32_t lane, void* arg) {
  PairJob* j = (PairJob*)arg;
  EmuWarp<32> W{w, lane};
  ChromView cv = j->e->cv();
  for (uint32_t p = j->lo; p < j->hi; ++p) {
    PairResult r = pair_candidates_wide(W, cv, GetEmuCand{j->r1 + (size_t)p * j->top_k}, j->n1[p],
                                        (uint32_t)(j->offs1[p + 1] - j->offs1[p]), GetEmuCand{j->r2 + (size_t)p * j->top_k},
                                        j->n2[p], (uint32_t)(j->offs2[p + 1] - j->offs2[p]), j->m, j->frag_range, j->sm);
    if (lane == 0) { j->out[p].best_times = r.best_times; j->out[p].best_i = r.best_i; j->out[p].best_j = r.best_j; j->out[p].frag_len = r.frag; }
    W.sync();
  }
}

template <class Job, class Fn>
int run_parallel(uint32_t n, int threads, uint32_t width, Fn make_and_run) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  std::atomic<int> diverged{0};
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      uint32_t lo = (uint32_t)((uint64_t)n * t / threads), hi = (uint32_t)((uint64_t)n * (t + 1) / threads);
      if (lo >= hi) return;
      WarpEmu* w = new WarpEmu;
      w->nf = width;
      g_current = w;
      make_and_run(w, lo, hi);
      if (w->diverged) diverged = 1;
      delete w;
    });
  }
  for (auto& th : pool) th.join();
  return diverged.load();
}

}  // namespace

extern "C" {

// returns 0 ok; 1 = warp primitives diverged (bug); 5 = non-ACGT read
int emu_map_se(void* h, const char* seqs, const uint64_t* offs, uint32_t n, int ag, uint32_t m,
               uint32_t b, int literal, emu_best* out, int threads, uint64_t* counters3, uint32_t width, int packed) {
  if (width != 8 && width != 16 && width != 32) return 1;
  EmuEngine* e = (EmuEngine*)h;
  uint32_t max_len = 1;
  for (uint32_t r = 0; r < n; ++r) max_len = std::max<uint32_t>(max_len, (uint32_t)(offs[r + 1] - offs[r]));
  if (max_len > MAX_READ_LEN) return 1;
  std::atomic<int> bad{0};
  std::atomic<uint64_t> c0{0}, c1{0}, c2{0};
  int div = run_parallel<SeJob>(n, threads, width, [&](WarpEmu* w, uint32_t lo, uint32_t hi) {
    std::vector<uint64_t> scratch(scratch_words((max_len + 31) / 32) + 8);
    uint32_t cached = 0; Counters ctr{0, 0, 0}; int b_ = 0;
    SeJob j{e, seqs, offs, lo, hi, ag, m, b, literal, out, max_len, scratch.data(), &cached, &ctr, &b_, packed};
    w->run(width == 8 ? se_lane<8> : width == 16 ? se_lane<16> : se_lane<32>, &j);
    if (b_) bad = 1;
    c0 += ctr.lookups; c1 += ctr.candidates; c2 += ctr.literal;
  });
  if (counters3) { counters3[0] = c0; counters3[1] = c1; counters3[2] = c2; }
  if (div) return 1;
  return bad ? 5 : 0;
}

int emu_map_pe_mate(void* h, const char* seqs, const uint64_t* offs, uint32_t n, int ag, uint32_t m,
                    uint32_t b, uint32_t top_k, int literal, emu_cand* ranked, uint32_t* n_ranked,
                    int threads, uint32_t width, int logged) {
  if (width != 8 && width != 16 && width != 32) return 1;
  if (logged && m > LOG_MAX_MM) return 1;
  EmuEngine* e = (EmuEngine*)h;
  uint32_t max_len = 1;
  for (uint32_t r = 0; r < n; ++r) max_len = std::max<uint32_t>(max_len, (uint32_t)(offs[r + 1] - offs[r]));
  if (max_len > MAX_READ_LEN) return 1;
  std::atomic<int> bad{0};
  int div = run_parallel<PeJob>(n, threads, width, [&](WarpEmu* w, uint32_t lo, uint32_t hi) {
    std::vector<uint64_t> scratch(scratch_words((max_len + 31) / 32) + 8);
    std::vector<HeapEntry> heap(top_k + 1), log(pe_log_slots(top_k, m) + 1);
    std::vector<uint32_t> hist(m + 2);
    int b_ = 0;
    PeJob j{e, seqs, offs, lo, hi, ag, m, b, top_k, literal, ranked, n_ranked, max_len, scratch.data(), heap.data(), &b_,
            logged, log.data(), hist.data()};
    w->run(width == 8 ? pe_lane<8> : width == 16 ? pe_lane<16> : pe_lane<32>, &j);
    if (b_) bad = 1;
  });
  if (div) return 1;
  return bad ? 5 : 0;
}

// pairing == pair_kernel (one thread per pair)
void emu_pair(void* h, const emu_cand* r1, const uint32_t* n1, const uint64_t* offs1,
              const emu_cand* r2, const uint32_t* n2, const uint64_t* offs2, uint32_t n,
              uint32_t top_k, uint32_t m, int frag_range, emu_pair* out) {
  EmuEngine* e = (EmuEngine*)h;
  ChromView cv = e->cv();
  for (uint32_t p = 0; p < n; ++p) {
    const emu_cand* a = r1 + (size_t)p * top_k;
    const emu_cand* b = r2 + (size_t)p * top_k;
    auto g1 = [a](uint32_t i) { RankedCand c; c.pos = a[i].genome_pos; c.mm = a[i].mismatch; c.strand = (uint8_t)a[i].strand; return c; };
    auto g2 = [b](uint32_t i) { RankedCand c; c.pos = b[i].genome_pos; c.mm = b[i].mismatch; c.strand = (uint8_t)b[i].strand; return c; };
    // pair_candidates takes one functor type; wrap through a common lambda type
    struct Get { const emu_cand* p; RankedCand operator()(uint32_t i) const { RankedCand c; c.pos = p[i].genome_pos; c.mm = p[i].mismatch; c.strand = (uint8_t)p[i].strand; return c; } };
    (void)g1; (void)g2;
    PairResult r = pair_candidates(cv, Get{a}, n1[p], (uint32_t)(offs1[p + 1] - offs1[p]), Get{b}, n2[p],
                                   (uint32_t)(offs2[p + 1] - offs2[p]), m, frag_range);
    out[p].best_times = r.best_times; out[p].best_i = r.best_i; out[p].best_j = r.best_j; out[p].frag_len = r.frag;
  }
}


// pairing by a whole (emulated) warp per pair == the long-list path of pair_kernel
int emu_pair_wide(void* h, const emu_cand* r1, const uint32_t* n1, const uint64_t* offs1,
                  const emu_cand* r2, const uint32_t* n2, const uint64_t* offs2, uint32_t n,
                  uint32_t top_k, uint32_t m, int frag_range, struct emu_pair* out, int threads) {
  EmuEngine* e = (EmuEngine*)h;
  if (m > 127) return 1;
  return run_parallel<PairJob>(n, threads, 32, [&](WarpEmu* w, uint32_t lo, uint32_t hi) {
    std::vector<uint32_t> sm(3u * top_k + 8u);
    PairJob j{e, r1, n1, offs1, r2, n2, offs2, lo, hi, top_k, m, frag_range, out, sm.data()};
    w->run(pair_wide_lane, &j);
  });
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// std::sort replay (walt_stdsort.cuh) against the real libstdc++ std::sort
// ------------------------------------------------------------------------------------------
extern "C" {

// cls[i] is the class rank of the element that starts in slot i (its "position" is i).  Writes
// the final arrangement of positions under our transcription and under std::sort with the same
// comparator; returns the number of heap-sort fallbacks our replay took.
int emu_stdsort_both(const uint32_t* cls, uint32_t n, uint32_t* out_ours, uint32_t* out_std) {
  std::vector<uint32_t> c(cls, cls + n), p(n);
  for (uint32_t i = 0; i < n; ++i) p[i] = i;
  const int fb = waltsort::std_sort(waltsort::PairSeq{p.data(), c.data()}, 0, (int64_t)n, waltsort::ByClass());
  memcpy(out_ours, p.data(), (size_t)n * 4u);
  std::vector<waltsort::Elem> v(n);
  for (uint32_t i = 0; i < n; ++i) v[i] = waltsort::Elem{i, cls[i]};
  std::sort(v.begin(), v.end(), [](const waltsort::Elem& a, const waltsort::Elem& b) { return a.cls < b.cls; });
  for (uint32_t i = 0; i < n; ++i) out_std[i] = v[i].pos;
  return fb;
}

// the same sort run the way the device builder runs it: level by level, one task at a time
void emu_stdsort_levels(const uint32_t* cls, uint32_t n, uint32_t* out_pos, uint32_t* n_levels) {
  std::vector<uint32_t> c(cls, cls + n), p(n);
  for (uint32_t i = 0; i < n; ++i) p[i] = i;
  const waltsort::PairSeq seq{p.data(), c.data()};
  std::vector<waltsort::SortTask> cur, nxt;
  if (n) cur.push_back(waltsort::SortTask{0u, n, waltsort::depth_limit_for(n)});
  uint32_t levels = 0;
  while (!cur.empty()) {
    nxt.clear();
    // any order inside a level is as good as any other: go backwards to make that point
    for (size_t i = cur.size(); i-- > 0;) {
      waltsort::SortTask kids[2];
      const int k = waltsort::sort_task_step(seq, cur[i], waltsort::ByClass(), kids);
      for (int j = 0; j < k; ++j) nxt.push_back(kids[j]);
    }
    cur.swap(nxt);
    ++levels;
  }
  memcpy(out_pos, p.data(), (size_t)n * 4u);
  if (n_levels) *n_levels = levels;
}

// the builder's "back to ascending position" step: heap sort by position of (pos, cls) pairs
void emu_heapsort_by_pos(uint32_t* pos, uint32_t* cls, uint32_t n) {
  waltsort::heap_sort_(waltsort::PairSeq{pos, cls}, 0, (int64_t)n, waltsort::ByPos());
}

// McIlroy's adversary ("A killer adversary for quicksort", 1999) run against the real std::sort:
// produces a permutation of 0..n-1 on which median-of-three introsort degenerates, so that the
// depth limit is hit and the heap-sort fallback of the replay gets exercised.
void emu_antiqsort(uint32_t n, uint32_t* out_vals) {
  std::vector<int> val(n, (int)n);   // n == "gas"
  std::vector<uint32_t> ptr(n);
  for (uint32_t i = 0; i < n; ++i) ptr[i] = i;
  int nsolid = 0, candidate = 0;
  const int gas = (int)n;
  std::sort(ptr.begin(), ptr.end(), [&](uint32_t x, uint32_t y) {
    if (val[x] == gas && val[y] == gas) {
      if ((int)x == candidate) val[x] = nsolid++; else val[y] = nsolid++;
    }
    if (val[x] == gas) candidate = (int)x;
    else if (val[y] == gas) candidate = (int)y;
    return val[x] < val[y];
  });
  for (uint32_t i = 0; i < n; ++i) out_vals[i] = (uint32_t)(val[i] == gas ? nsolid++ : val[i]);
}

}  // extern "C"
