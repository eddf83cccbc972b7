// CPU SIMT emulation of the sharded head step's exchange (TEST INFRASTRUCTURE ONLY; built and run by
// tests/test_xchg_emulated_cpu.py with g++, no GPU).
//
// Runs the UNMODIFIED device code of csrc/step_kernels.cuh (head_prep_kernel, head_finalize_kernel, head_finish_kernel),
// csrc/stepx_acquire.inc + stepx_cta0.inc (the prologue of the fused kernel's MODE_STEPX instantiation) and the helpers
// of csrc/step_xchg.cuh / xchg.cuh for `world` emulated ranks, one OS thread each, with the launch sequence of
// simt_head_step_sharded / simt_head_finish_sharded (csrc/head.cu: run_step) and the host logic of HeadRunner.step /
// finish (simt_b200/head.py).  Only the fused kernel's pixel loop is replaced: its outputs (per-CTA loss / count
// partials, per-SM dT tiles) are synthesised from a seed as dyadic rationals, so every sum the kernels form is exact
// and the expected loss / dT / stats of every step are known bit for bit.
// Peer stores are delayed and reordered by the shim (tests/cpu_simt/cuda_shim.h); ranks are skewed by random sleeps.
//
//   xchg_emul <world> <steps> <mode> <seed> [die_rank die_step]
//   mode: sync | pipelined | mixed | announce_sync | pipelined_noannounce | mixed_nodrain (negative control)
//   environment: XCHG_EMUL_WATCHDOG_S = seconds before unfinished ranks count as dead-locked (default 240)
//   die_rank / die_step: that rank stops before that step; the others must time out, poison their outputs with NaN
//   and raise SIMT_ERRBIT_XCHG_TIMEOUT (never continue with a partial sum).
// Exit code 0 = every check passed on every rank.
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>

#define SIMT_CPU_EMULATION 1
#include "step_kernels.cuh"

namespace cpusimt {
thread_local Rank* R = nullptr;
[[noreturn]] void die(const char* what) {
  fprintf(stderr, "EMULATION FAILURE: %s\n", what);
  fflush(stderr);
  _exit(3);
}
void debug_backtrace_and_exit() { die("debug backtrace requested"); }
}  // namespace cpusimt

using namespace simt;

// ---- problem (small, odd sizes) -------------------------------------------------------------------------------------
#ifndef EMU_C        /* -DEMU_C=19 -DEMU_CK=19: the bench geometry (13 finalize blocks, 2 finish blocks) */
#define EMU_C 4
#define EMU_CK 5
#endif
#ifndef EMU_CKP      /* row length of the kernel's dT tile: 20 for CK <= 20 (CPL 10 x LPR 2), 24 for CK <= 24 (12 x 2) */
#define EMU_CKP 20
#endif
static constexpr int C = EMU_C, CK = EMU_CK, CKP = EMU_CKP;
static_assert(CK <= CKP, "the dT tile holds CKP >= CK channels per row");
static constexpr int NSTATS = 2 + CK * C;
static constexpr int G = 3;                           // CTAs of the fused kernel (loss / count partials)
static constexpr int NTILES = 2;                      // per-SM dT tiles
static constexpr int NT = 64;                         // threads per CTA of the emulated prologue
static constexpr long long NPIX = 777;                // labels per rank (ragged: exercises the tails of the count pass)
static constexpr long long NDL = 101;
static constexpr int IGNORE = 255;
static constexpr float GRAD_OUT = 0.75f;

struct Local {                    // what one rank's fused kernel would have produced for one step
  std::vector<uint8_t> labels;
  double part_loss[G];
  long long part_cnt[G];
  float tiles[NTILES][C * CKP];
  long long count;                // valid labels
  double stats[NSTATS];           // the rank's local stats exactly as finalize forms them
};

static Local make_local(unsigned long long seed, int rank, unsigned long long step) {
  std::mt19937_64 g(seed * 1000003ULL + (unsigned long long)rank * 7919ULL + step * 104729ULL + 17ULL);
  Local L;
  L.labels.resize((size_t)NPIX);
  L.count = 0;
  for (auto& v : L.labels) {
    const unsigned r = (unsigned)(g() % 10);
    v = (r == 0) ? (uint8_t)IGNORE : (uint8_t)(g() % C);
    L.count += (v != IGNORE);
  }
  long long left = L.count;
  double lsum = 0.0;
  for (int i = 0; i < G; ++i) {
    L.part_loss[i] = (double)((long long)(g() % 4096) - 2048) / 64.0;     // dyadic: every sum is exact
    lsum += L.part_loss[i];
    L.part_cnt[i] = (i == G - 1) ? left : (long long)(g() % (unsigned long long)(left + 1));
    left -= L.part_cnt[i];
  }
  for (int t = 0; t < NTILES; ++t)
    for (int o = 0; o < C * CKP; ++o) {
      const int k = o % CKP;
      L.tiles[t][o] = (k < CK) ? (float)((long long)(g() % 2048) - 1024) / 32.0f : 0.f;   // padded channels carry zeros
    }
  L.stats[0] = -kLn2 * lsum;
  L.stats[1] = (double)L.count;
  for (int k = 0; k < CK; ++k)
    for (int y = 0; y < C; ++y) {
      double t = 0.0;
      for (int tl = 0; tl < NTILES; ++tl) t += (double)L.tiles[tl][y * CKP + k];
      L.stats[2 + k * C + y] = -t;
    }
  return L;
}

struct Expected { double stats[NSTATS]; float loss; float dT[CK * C]; float gs; };

static Expected expected_of(unsigned long long seed, int world, unsigned long long step) {
  Expected E;
  double cnt = 0.0;
  std::vector<Local> Ls;
  for (int r = 0; r < world; ++r) { Ls.push_back(make_local(seed, r, step)); cnt += (double)Ls.back().count; }
  for (int i = 0; i < NSTATS; ++i) {
    double t = 0.0;
    for (int r = 0; r < world; ++r) t += Ls[(size_t)r].stats[i];     // rank order, as the kernels sum
    E.stats[i] = t;
  }
  const double sc = (double)GRAD_OUT / cnt;
  E.loss = (float)(E.stats[0] / E.stats[1]);
  for (int i = 0; i < CK * C; ++i) E.dT[i] = (float)(E.stats[2 + i] * sc);
  E.gs = (float)((double)GRAD_OUT / cnt);
  return E;
}

// ---- kernel launch wrappers -----------------------------------------------------------------------------------------
struct PrepArgs { float* dl; long long n_dl; const uint8_t* lab; const uint8_t* next; long long npix; unsigned long long* ws; XchgArgs X; FinishArgs F; };
static void prep_body(void* p) {
  PrepArgs* a = (PrepArgs*)p;
  head_prep_kernel<uint8_t>(a->dl, a->n_dl, a->lab, a->next, a->npix, C, IGNORE, a->ws, a->X, a->F);
}

// the members of HeadArgs the prologue (.inc files) touches
struct EmuHeadArgs {
  XchgArgs X; const float* grad_out; int* err; double* count_global; unsigned long long* ws_hdr; FinishArgs fin;
  float* gs_out;    // [G]: s_gs of every CTA, for the check
};
static void prologue_body(void* p) {
  const EmuHeadArgs A = *(EmuHeadArgs*)p;
  constexpr int MODE = MODE_STEPX;
  __shared__ float s_gs;
  const int tid = (int)threadIdx.x;
#include "stepx_acquire.inc"
#include "stepx_cta0.inc"
  __syncthreads();
  if (tid == 0) A.gs_out[blockIdx.x] = s_gs;
}

struct FinArgs {
  float* part_dT; const double* part_loss; const long long* part_cnt; unsigned long long* counter; double* stats;
  float* loss; float* dT; int* err; const float* grad_out; const double* count_dev; unsigned long long* ws; XchgArgs X; int defer;
};
static void finalize_body(void* p) {
  FinArgs* a = (FinArgs*)p;
  head_finalize_kernel(a->part_dT, a->part_loss, a->part_cnt, G, NTILES, CK, CKP, C, MODE_STEP, 1.f, a->counter, a->stats,
                       a->loss, a->dT, a->err, a->grad_out, a->count_dev, a->ws, a->X, a->defer);
}
struct FinishKArgs { unsigned long long* ws; XchgArgs X; FinishArgs F; };
static void finish_body(void* p) {
  FinishKArgs* a = (FinishKArgs*)p;
  head_finish_kernel(a->ws, a->X, a->F);
}

// ---- head_prep_kernel alone: label dtypes, alignments, sizes --------------------------------------------------------
template <typename LabelT>
struct PrepArgsT { float* dl; long long n_dl; const LabelT* lab; long long npix; int Cc; int ignore; unsigned long long* ws; };
template <typename LabelT>
static void prep_only_body(void* p) {
  auto* a = (PrepArgsT<LabelT>*)p;
  head_prep_kernel<LabelT>(a->dl, a->n_dl, a->lab, nullptr, a->npix, a->Cc, a->ignore, a->ws, XchgArgs{}, FinishArgs{});
}
static int prep_cases(unsigned long long seed) {
  cpusimt::Rank emu;
  emu.rng.seed(seed);
  cpusimt::R = &emu;
  std::mt19937_64 g(seed + 99);
  int bad = 0;
  for (int it = 0; it < 40; ++it) {
    const long long npix = (long long)(g() % 3000);
    const long long n_dl = (long long)(g() % 500);
    const int off8 = (int)(g() % 17), offd = (int)(g() % 5);       // misaligned label / dLogits pointers
    const int Cc = 1 + (int)(g() % 40);
    const int ignore = (g() % 4 == 0) ? (int)(g() % 300) : 255;   // also ignore labels outside the uint8 range
    const bool i64 = g() % 2;
    std::vector<uint8_t> l8((size_t)npix + 64);
    std::vector<long long> l64((size_t)npix + 8);
    long long want = 0;
    for (long long i = 0; i < npix; ++i) {
      const long long v = i64 ? (long long)(g() % 300) - 20 : (long long)(g() % 256);
      l8[(size_t)(off8 + i)] = (uint8_t)v; l64[(size_t)(off8 % 8 + i)] = v;
      want += (v >= 0 && v < Cc && v != ignore);
    }
    std::vector<float> dl((size_t)n_dl + 8, 7.f);
    std::vector<unsigned long long> ws(32, 0ULL);
    const unsigned grid = 1 + (unsigned)(g() % 5), block = 32u * (1 + (unsigned)(g() % 8));
    if (i64) {
      PrepArgsT<long long> a{dl.data() + offd, n_dl, l64.data() + off8 % 8, npix, Cc, ignore, ws.data()};
      cpusimt::launch(grid, block, prep_only_body<long long>, &a);
    } else {
      PrepArgsT<uint8_t> a{dl.data() + offd, n_dl, l8.data() + off8, npix, Cc, ignore, ws.data()};
      cpusimt::launch(grid, block, prep_only_body<uint8_t>, &a);
    }
    const double got = *reinterpret_cast<double*>(&ws[WS_COUNT_LOCAL]);
    bool ok = got == (double)want && ws[WS_ACCUM] == 0 && ws[WS_TICKET] == 0;
    for (long long i = 0; i < n_dl; ++i) ok = ok && dl[(size_t)(offd + i)] == 0.f;
    for (int i = 0; i < offd; ++i) ok = ok && dl[(size_t)i] == 7.f;                 // nothing outside the range is touched
    for (size_t i = (size_t)(offd + n_dl); i < dl.size(); ++i) ok = ok && dl[i] == 7.f;
    if (!ok) {
      fprintf(stderr, "prep case %d: npix %lld n_dl %lld int64 %d C %d ignore %d grid %u block %u: count %g vs %lld\n", it, npix, n_dl,
              (int)i64, Cc, ignore, grid, block, got, want);
      ++bad;
    }
  }
  return bad;
}

// ---- one emulated rank ----------------------------------------------------------------------------------------------
struct Shared {
  int world, steps;
  std::string mode;
  unsigned long long seed;
  int die_rank = -1, die_step = -1;
  std::vector<unsigned char*> mail;
  std::atomic<int> failures{0};
  std::atomic<int> done{0};
  std::atomic<unsigned long long> switches{0};
};

static bool step_defers(const std::string& m, int s) {
  if (m == "pipelined" || m == "pipelined_noannounce") return true;
  if (m == "mixed" || m == "mixed_nodrain") return s % 3 != 0;
  return false;
}
static bool step_announces(const std::string& m, int s, int steps) {
  if (s >= steps) return false;
  if (m == "pipelined") return true;
  if (m == "mixed" || m == "mixed_nodrain") return s % 3 != 0;
  if (m == "announce_sync") return true;
  return false;
}

#define CHECK(cond, ...)                                                          \
  do {                                                                            \
    if (!(cond)) {                                                                \
      if (S->failures++ < 12) {                                                   \
        fprintf(stderr, "rank %d step %d: CHECK FAILED %s: ", rank, s, #cond);   \
        fprintf(stderr, __VA_ARGS__);                                             \
        fprintf(stderr, "\n");                                                    \
      }                                                                           \
    }                                                                             \
  } while (0)

static bool same_bits(float a, float b) { return memcmp(&a, &b, 4) == 0; }
static bool same_bits(double a, double b) { return memcmp(&a, &b, 8) == 0; }

static void rank_main(Shared* S, int rank) {
  cpusimt::Rank emu;
  emu.rng.seed(S->seed * 31ULL + (unsigned long long)rank * 1009ULL + 5ULL);
  emu.p_deliver = 0.05 + 0.1 * (double)(emu.rng() % 8);           // some ranks have a slow "NVLink"
  emu.p_flush_at_kernel_end = (emu.rng() % 2) ? 0.3 : 0.9;
  cpusimt::R = &emu;
  const bool fault = S->die_rank >= 0;

  const bool single = S->world == 1;      // simt_head_step on one GPU: XchgArgs{} (world 0), MODE_STEP, no exchange at all
  XchgArgs X{};
  if (!single) {
    for (int r = 0; r < S->world; ++r) X.mail[r] = S->mail[(size_t)r];
    X.rank = rank; X.world = S->world; X.n_stats = NSTATS;
    X.slot_entries = 2 + C * kXchgMaxCKP;
  }
  // no fault injected: wait for ever (a lost word shows up as the watchdog's dead-lock); the negative controls and the
  // fault runs use the kernels' own bounded waits, so a lost word shows up as poison + error bit within seconds
  const char* spins = getenv("XCHG_EMUL_MAX_SPINS");
  X.max_spins = spins ? atoll(spins) : ((fault || S->mode == "mixed_nodrain") ? 3000 : 0);

  // device buffers of this rank
  std::vector<unsigned long long> ws(kWsHeader / 8 + NSTATS + 16, 0ULL);
  std::vector<double> stats(NSTATS, -777.0);
  std::vector<float> dT(CK * C, -777.f), dl((size_t)NDL, 3.f), tiles((size_t)NTILES * C * CKP, 0.f), gs(G, 0.f);
  std::vector<double> part_loss(G, 0.0);
  std::vector<long long> part_cnt(G, 0);
  float loss = -777.f, grad_out = GRAD_OUT;
  int err = 0;
  FinishArgs F{stats.data(), &loss, dT.data(), &grad_out, &err, CK, C, CKP};

  // host-side mirror of the protocol state (what the outputs must hold after every call)
  bool deferred = false;
  unsigned long long pending[2] = {0, 0}, unsent = 0, last_finished = 0;
  int s = 0;

  auto run_finish = [&]() {
    FinishKArgs a{ws.data(), X, F};
    cpusimt::launch((unsigned)((NSTATS + 255) / 256), 256, finish_body, &a);
    for (int q = 0; q < 2; ++q) if (pending[q] > last_finished) last_finished = pending[q];
    if (unsent) last_finished = unsent;
    pending[0] = pending[1] = 0; unsent = 0;
    deferred = false;
  };
  auto check_outputs = [&](bool stats_too) {
    if (last_finished == 0 || fault) return;
    const Expected E = expected_of(S->seed, S->world, last_finished);
    CHECK(same_bits(loss, E.loss), "loss %.9g vs %.9g (results of step %llu)", (double)loss, (double)E.loss, last_finished);
    for (int i = 0; i < CK * C; ++i)
      CHECK(same_bits(dT[(size_t)i], E.dT[i]), "dT[%d] %.9g vs %.9g (step %llu)", i, (double)dT[(size_t)i], (double)E.dT[i], last_finished);
    if (stats_too)
      for (int i = 0; i < NSTATS; ++i)
        CHECK(same_bits(stats[(size_t)i], E.stats[i]), "stats[%d] %.17g vs %.17g (step %llu)", i, stats[(size_t)i], E.stats[i], last_finished);
  };

  for (s = 1; s <= S->steps; ++s) {
    if (fault && rank == S->die_rank && s == S->die_step) break;      // this rank "dies"
    usleep((useconds_t)(emu.rng() % 300));                            // host-side skew between the ranks
    const bool defer = step_defers(S->mode, s);
    const bool announce = step_announces(S->mode, s, S->steps);
    const Local L = make_local(S->seed, rank, (unsigned long long)s);
    const Local Ln = make_local(S->seed, rank, (unsigned long long)s + 1);
    // ---- HeadRunner.step ----
    // (mode mixed_nodrain is the NEGATIVE CONTROL: without this drain a synchronous step overwrites stats slots that
    // a peer has not reduced yet -- the emulation must catch it)
    if (!defer && deferred && S->mode != "mixed_nodrain") { run_finish(); check_outputs(true); }
    deferred = defer;
    // ---- run_step (csrc/head.cu): prep, fused kernel (prologue emulated, pixel loop synthesised), finalize ----
    std::fill(dl.begin(), dl.end(), 3.f);
    PrepArgs pa{dl.data(), NDL, L.labels.data(), announce ? Ln.labels.data() : nullptr, NPIX, ws.data(), X, F};
    cpusimt::launch(3, 64, prep_body, &pa);
    for (float v : dl) CHECK(v == 0.f, "dLogits not zeroed");
    CHECK(*reinterpret_cast<double*>(&ws[WS_COUNT_LOCAL]) == (double)L.count, "local count %g vs %lld",
          *reinterpret_cast<double*>(&ws[WS_COUNT_LOCAL]), L.count);
    if (emu.rng() % 2) usleep((useconds_t)(emu.rng() % 200));
    EmuHeadArgs ha{X, &grad_out, &err, reinterpret_cast<double*>(&ws[WS_COUNT_GLOBAL]), ws.data(), F, gs.data()};
    if (!single) cpusimt::launch(G, NT, prologue_body, &ha);
    else         // MODE_STEP (head_kernel.cuh): s_gs = grad_out / count_local
      for (int g = 0; g < G; ++g) gs[(size_t)g] = (float)((double)grad_out / *reinterpret_cast<double*>(&ws[WS_COUNT_LOCAL]));
    {   // mirror of stepx_cta0.inc
      const unsigned long long pend = pending[s & 1];
      if (pend != 0 && pend + 2 <= (unsigned long long)s) { last_finished = pend; pending[s & 1] = 0; }
      if (unsent) { pending[unsent & 1] = unsent; unsent = 0; }
    }
    if (!fault) {
      const Expected Es = expected_of(S->seed, S->world, (unsigned long long)s);
      for (int g = 0; g < G; ++g) CHECK(same_bits(gs[(size_t)g], Es.gs), "grad scale of CTA %d: %.9g vs %.9g", g, (double)gs[(size_t)g], (double)Es.gs);
    }
    check_outputs(false);
    // the pixel loop's outputs
    for (int g = 0; g < G; ++g) { part_loss[(size_t)g] = L.part_loss[g]; part_cnt[(size_t)g] = L.part_cnt[g]; }
    // (finalize forms ls = -kLn2 * sum(part_loss); make_local's stats[0] uses the same expression)
    for (int t = 0; t < NTILES; ++t) memcpy(&tiles[(size_t)t * C * CKP], L.tiles[t], sizeof(L.tiles[t]));
    ws[WS_COUNTER] = 12345ULL;                                        // the unit scheduler's counter after a launch
    FinArgs fa{tiles.data(), part_loss.data(), part_cnt.data(), &ws[WS_COUNTER], stats.data(), &loss, dT.data(), &err,
               &grad_out, reinterpret_cast<double*>(&ws[WS_COUNT_LOCAL]), ws.data(), X, defer ? 1 : 0};
    cpusimt::launch((unsigned)((C * CKP + 31) / 32 + 1), 1024, finalize_body, &fa);
    if (defer) unsent = (unsigned long long)s; else last_finished = (unsigned long long)s;
    CHECK(ws[WS_COUNTER] == 0ULL, "unit scheduler not re-armed");
    for (float v : tiles) CHECK(v == 0.f, "dT tiles not re-zeroed");
    if (!single) CHECK(*reinterpret_cast<volatile unsigned long long*>(X.mail[rank]) == (unsigned long long)s, "step counter");
    if (fault) {
      if (s >= S->die_step && S->mode == "sync") {      // a peer is gone: poison, never a partial sum
        CHECK((err & SIMT_ERRBIT_XCHG_TIMEOUT) != 0, "no timeout flagged (err %d)", err);
        CHECK(loss != loss, "loss is %.9g, not NaN", (double)loss);
        for (float v : dT) CHECK(v != v, "dT holds %.9g, not NaN", (double)v);
        for (int g = 0; g < G; ++g) CHECK(gs[(size_t)g] != gs[(size_t)g], "grad scale is %.9g, not NaN", (double)gs[(size_t)g]);
      }
    } else {
      check_outputs(!defer);
    }
  }
  if (fault && rank != S->die_rank) {
    // the pipelined forms notice a missing peer one or two steps later; after finish() every survivor must have flagged it
    --s;
    run_finish();
    CHECK((err & SIMT_ERRBIT_XCHG_TIMEOUT) != 0, "no timeout flagged after finish (err %d)", err);
    CHECK(loss != loss, "loss is %.9g after finish, not NaN", (double)loss);
  }
  if (!fault) {
    --s;
    if (!single) run_finish();         // HeadRunner.finish() after the last step (a no-op kernel when nothing is outstanding)
    check_outputs(true);
    CHECK(last_finished == (unsigned long long)S->steps, "last finished step %llu", last_finished);
    CHECK(err == 0, "error flag %d", err);
    CHECK(ws[WS_PENDING] == 0 && ws[WS_PENDING_ODD] == 0 && ws[WS_UNSENT] == 0 && ws[WS_COUNT_NEXT] == 0, "protocol words left set");
  }
  emu.flush();
  S->switches += emu.n_switch;
  S->done++;
}

int main(int argc, char** argv) {
  if (argc == 3 && std::string(argv[1]) == "prep") {
    const int bad = prep_cases(strtoull(argv[2], nullptr, 10));
    printf("%s prep: %d failed cases\n", bad ? "FAIL" : "OK", bad);
    return bad ? 1 : 0;
  }
  if (argc < 5) { fprintf(stderr, "usage: xchg_emul world steps mode seed [die_rank die_step] | xchg_emul prep seed\n"); return 2; }
  Shared S;
  S.world = atoi(argv[1]); S.steps = atoi(argv[2]); S.mode = argv[3]; S.seed = strtoull(argv[4], nullptr, 10);
  if (argc >= 7) { S.die_rank = atoi(argv[5]); S.die_step = atoi(argv[6]); }
  if (S.world < 1 || S.world > kMaxPeers) return 2;
  const size_t slot = 2 + (size_t)C * kXchgMaxCKP;
  const size_t bytes = kHdrBytes + kCountBytes + (size_t)2 * kMaxPeers * (2 * slot) * sizeof(unsigned long long);   // simt_xchg_bytes(C)
  for (int r = 0; r < S.world; ++r) S.mail.push_back((unsigned char*)calloc(bytes, 1));
  std::vector<std::thread> th;
  for (int r = 0; r < S.world; ++r) th.emplace_back(rank_main, &S, r);
  const int expect_done = S.world;
  const auto t0 = std::chrono::steady_clock::now();
  while (S.done.load() < expect_done) {
    usleep(20000);
    const char* wd = getenv("XCHG_EMUL_WATCHDOG_S");
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(wd ? atoi(wd) : 240)) cpusimt::die("watchdog: the ranks did not finish (a word was lost or a rank waits for ever)");
  }
  for (auto& t : th) t.join();
  const int f = S.failures.load();
  printf("%s world=%d steps=%d mode=%s seed=%llu%s: %d failed checks, %llu fiber switches\n", f ? "FAIL" : "OK", S.world, S.steps,
         S.mode.c_str(), S.seed, S.die_rank >= 0 ? " (fault injected)" : "", f, S.switches.load());
  return f ? 1 : 0;
}
