// Third-generation schedule of hot path A for N >= 256: ONE persistent launch per call.
//
// The second-generation kernels (ns2d_v2.cuh) run a substage as two grid-wide launches, so the
// intermediate fields H (4 spectra-sized arrays per sample) and advt make a round trip through HBM:
// 835 MB of DRAM traffic per substage at 512^2 x 64 against 242 MB of algorithmic state traffic.
// Here the same unit bodies (rows item = one double row pair of one sample, cols item = one quad of
// physical columns of one sample) are work items of a single persistent kernel:
//
//   * items are handed out in a fixed order by a global ticket counter; the order is CHUNK-MAJOR:
//     for each chunk of W samples -> [prologue rows] -> { [cols], [rows] } x (steps x stages), each
//     phase sample-major; the workspaces (state, H, advt) hold W slots that the next chunk re-uses.
//     (A small W keeps them in the 126 MB L2 -- DRAM traffic drops 7x -- but the step is not DRAM
//     bound and a short window is limited by the per-sample dependency chain: measured best W = 64,
//     see profiles/r04_flow_sweep.md);
//   * the dependencies (a cols item needs every rows unit of its sample and substage, a rows unit
//     every cols quad of the previous phase) are per-sample counters in global memory: producers
//     add 1 with release semantics, the consumer's thread 0 polls with acquire semantics.  An item
//     only ever waits for items with SMALLER tickets, which are already held by resident CTAs, so
//     the schedule cannot deadlock whatever the number of resident CTAs is (the polling loop is
//     nevertheless bounded by the SM clock: it raises a host-visible error word and traps instead of
//     hanging).  Every consumer read of another CTA's stores goes through the TMA / bulk-copy engine
//     (L2, never a possibly stale L1 line), behind acquire + fence.proxy.async;
//   * the next ticket is fetched while the current item is processed;
//   * an item is a GROUP of consecutive units of one sample (GR double rows / GC column quads): one
//     ticket, one dependency poll and one release per group, and inside the group the inputs of unit
//     g+1 (state, tables, advection row, H tile) are staged by the TMA engine while unit g is being
//     transformed.
//
// Arithmetic, layouts of H / advt / the unit-layout state and the table blocks: ns2d_v2.cuh.
#pragma once
#include "ns2d_v2.cuh"

namespace tcfd {

constexpr int FLOW_MAX_STAGES = 8;

template <class T>
struct FlowParams {
  NsParams<T> p;  // p.B = samples of this call; workspaces hold `slots` samples
  int nsub;       // substages of the call = steps * nstages
  int nstages;
  int W;          // samples per chunk (== slots of H / advt / wU / hU)
  T beta[FLOW_MAX_STAGES], gdt[FLOW_MAX_STAGES], mu[FLOW_MAX_STAGES];
  unsigned char rd_h[FLOW_MAX_STAGES], wr_h[FLOW_MAX_STAGES];
  cx<typename pack2<T>::type>* wU;  // [W] unit-layout state, updated in place
  cx<typename pack2<T>::type>* hU;  // [W]
  cx<typename pack2<T>::type>* w0U; // [W] unit-layout copy of the call's input state (dw/dt), or null
  int* sync;  // [0] ticket, [1] unused, [2 .. 2+B) rows counters, [2+B .. 2+2B) cols counters
  int* err;   // host-visible error word (0 = ok)
  long long wait_cycles;  // bound of a dependency wait in SM cycles (0 = unbounded); TCFD_FLOW_TIMEOUT_S, default 3 s
  unsigned long long* prof;  // [16] cycles of thread 0 per region, summed over CTAs (profiling variant only)
};

// ------------------------------------------------------------------------------------------ sync
TCFD_D int flow_fetch_add(int* p, int v) {
#ifndef TCFD_EMU
  // inline PTX, not atomicAdd(): the compiler's warp-aggregated form of the intrinsic ends in a shuffle that waits for
  // the returned value on the spot, while the ticket is only needed at the end of the item
  int o;
  asm volatile("atom.relaxed.gpu.global.add.s32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
#else
  const int o = *p;
  *p = o + v;
  return o;
#endif
}
TCFD_D int flow_ld_acquire(const int* p) {
#ifndef TCFD_EMU
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
}
// one thread, after a CTA barrier that follows the item's global stores
TCFD_D void flow_signal(int* p) {
#ifndef TCFD_EMU
  // release at gpu scope: cumulative over the stores of the whole CTA (ordered before by bar.sync)
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
#else
  __atomic_fetch_add(p, 1, __ATOMIC_RELEASE);
#endif
}
// one thread: wait until *p >= target.  A wait that exceeds the bound (default ~3 s of SM clock; TCFD_FLOW_TIMEOUT_S,
// 0 disables it for debuggers / sanitizers / time-sliced GPUs) is a lost dependency: the error word is raised and the
// grid is killed instead of hanging.
TCFD_D void flow_wait(const int* p, int target, int* err, long long wait_cycles) {
#ifndef TCFD_EMU
  if (flow_ld_acquire(p) < target) {
    const long long t0 = clock64();
    unsigned ns = 32;
    for (unsigned spin = 1;; ++spin) {
      __nanosleep(ns);
      if (ns < 1024) ns *= 2;  // back off: hundreds of CTAs polling one word would starve the producers' REDs
      if (flow_ld_acquire(p) >= target) break;
      if ((spin & 255u) == 0 && (*reinterpret_cast<volatile int*>(err) != 0 || (wait_cycles > 0 && clock64() - t0 > wait_cycles))) {
        *reinterpret_cast<volatile int*>(err) = 1;
        __threadfence_system();
        asm volatile("trap;");
      }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");  // bulk / TMA reads of the producer's stores follow
#else
  // emulation: CTAs run one after the other, so a dependency can only be pending while another thread of
  // THIS CTA is about to release the previous item; anything longer is a schedule bug
  for (long spin = 0; __atomic_load_n(p, __ATOMIC_ACQUIRE) < target; ++spin) {
    std::this_thread::yield();
    if (spin > 200000000L) {
      *err = 1;
      std::abort();
    }
  }
#endif
}
// the caller's arrays are touched once per call: streaming (evict-first) accesses keep them from
// displacing the workspace window that lives in L2
template <class T>
TCFD_D cx<T> flow_ld_stream(const cx<T>* p) {
#ifndef TCFD_EMU
  if constexpr (sizeof(T) == 4) {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
    return cx<T>{v.x, v.y};
  } else {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return cx<T>{v.x, v.y};
  }
#else
  return *p;
#endif
}
template <class T>
TCFD_D void flow_st_stream(cx<T>* p, cx<T> v) {
#ifndef TCFD_EMU
  if constexpr (sizeof(T) == 4) __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
  else __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
#else
  *p = v;
#endif
}
TCFD_D long long flow_clock() {
#ifndef TCFD_EMU
  return clock64();
#else
  return 0;
#endif
}
TCFD_D void flow_proxy_fence() {
#ifndef TCFD_EMU
  asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
}
// ------------------------------------------------------------------------------------------ smem
// One layout for both roles:
//   cols: [ tile ]                                              [ buf ] [ barriers ]
//   rows: [ W | H | ADV (advection row / dw/dt reference) | TAB0 | MASK0 | TAB1 | MASK1 ]
template <class T, int N, int VB = 1>  // VB: transforms a thread carries side by side (exchange buffer size)
struct FlowSmem {
  typedef typename pack2<T>::type L;
  typedef RowsSmem<T, N> R;
  static constexpr int NH = N / 2 + 1;
  static constexpr int IB = 4 * (int)sizeof(cx<L>);
  typedef TileGeom<N, IB> G;
  static constexpr int TABM = R::TAB_BYTES + 2 * R::MASK_ROW;  // one table block + its mask rows
  static constexpr int OFF_W = 0;
  static constexpr int OFF_H = R::UNIT_BYTES;
  static constexpr int OFF_ADV = (2 * R::UNIT_BYTES + 127) / 128 * 128;  // UNIT_BYTES >= N entries; tiled TMA destination: 128 B aligned
  static constexpr int OFF_TAB = (OFF_ADV + R::UNIT_BYTES + 15) / 16 * 16;
  static constexpr int ROWS_STAGE = OFF_TAB + 2 * TABM;
  static constexpr int AREA = (G::BYTES > ROWS_STAGE ? G::BYTES : ROWS_STAGE);
  static constexpr int OFF_BUF = (AREA + 127) / 128 * 128;
  static constexpr int OFF_BAR = OFF_BUF + VB * N * (int)sizeof(cx<L>);
  // per-CTA copies of the small per-row tables every rows unit reads (kappa_x[N], forced-row flags[N]): shared-memory
  // reads instead of four dependent global loads at the head of every unit (long-scoreboard stalls in the ncu source view)
  static constexpr int OFF_KX = OFF_BAR + 64;
  static constexpr int OFF_FR = OFF_KX + N * (int)sizeof(T);
  static constexpr int BYTES = OFF_FR + (N + 15) / 16 * 16 + 1024;  // + slack for the 1 KB alignment of the area
};

template <class L>
TCFD_D cx<typename lane_traits<L>::scalar> lane_rt(cx<L> v, int lane) {
  typedef typename lane_traits<L>::scalar T;
  return lane ? cx<T>{v.x.hi, v.y.hi} : cx<T>{v.x.lo, v.y.lo};
}

// ------------------------------------------------------------------------------------------ kernel
// Timing experiment (-DTCFD_FLOW_VARIANTS builds, MODE 1): the transforms are replaced by one barrier, so
// the launch measures everything BUT the FFTs (results are meaningless).
// MODE bit 6: thread 0 of every CTA attributes its cycles (clock64) to regions -- 0 control (ticket, decode,
// release), 1 dependency wait, 2 staging wait (mbarrier), 3 transforms, 4 cols arithmetic + stores, 5 rows
// unit head (scalars, advection row), 6 rows spectral fields + H stores, 7 end-of-item barrier, 8 rows prologue
// loads, 9 rows RK/CN update proper, 10 rows post-update barrier + staging of the next unit -- summed into fp.prof.
#define FLOW_MARK(R)                                                       \
  do {                                                                     \
    if constexpr ((MODE & 64) != 0) {                                      \
      if (ctl) {                                                        \
        const long long now_ = flow_clock();                               \
        prof_acc[prof_cur] += now_ - prof_last;                            \
        prof_last = now_;                                                  \
        prof_cur = (R);                                                    \
      }                                                                    \
    }                                                                      \
  } while (0)
#define FLOW_FFT(DIR, ARR)                                                \
  do {                                                                    \
    const int pr_ = prof_cur;                                             \
    FLOW_MARK(3);                                                         \
    if constexpr ((MODE & 1) != 0) __syncthreads();                       \
    else fft_run<L, N, DIR, 1, PP, N>(ARR, tw, buf, parity, t, sync);     \
    FLOW_MARK(pr_);                                                       \
  } while (0)
#define FLOW_LOADWAIT(BAR, PH)  \
  do {                          \
    const int pr_ = prof_cur;   \
    FLOW_MARK(2);               \
    tile_load_wait(BAR, PH);    \
    FLOW_MARK(pr_);             \
  } while (0)
#define FLOW_DEPWAIT(PTR, TGT)        \
  do {                                \
    const int pr_ = prof_cur;         \
    FLOW_MARK(1);                     \
    flow_wait(PTR, TGT, fp.err, fp.wait_cycles); \
    FLOW_MARK(pr_);                   \
  } while (0)
// GR: double rows per rows item, GC: column quads per cols item (GC divides N/4).
template <class T, int N, int MINB, int GR, int GC, int MODE = 0>
__global__ void __launch_bounds__(N / 8, MINB)
ns2d_flow_kernel(const FlowParams<T> fp, const
#ifndef TCFD_EMU
                 __grid_constant__
#endif
                 TileMaps maps) {
  typedef typename pack2<T>::type L;
  // MODE bit 1: ping-pong exchange buffers (one barrier per exchange instead of two; +N entries of shared memory)
  constexpr bool PP = (MODE & 2) != 0;
  typedef FlowSmem<T, N, PP ? 2 : 1> S;
  typedef typename S::G G;
  typedef typename S::R R;
  constexpr int NT = N / 8, NH = N / 2 + 1, ND = N / 4 + 1, NQ = N / 4;
  constexpr int JB = (ND + ADV_BLOCK - 1) / ADV_BLOCK;  // blocks of advection rows per slot
  constexpr int IR = (ND + GR - 1) / GR;  // rows items per sample and phase
  constexpr int IC = NQ / GC;             // cols items per sample and phase
  static_assert(NQ % GC == 0, "GC must divide N/4");
  constexpr int IB = S::IB;
  const NsParams<T>& p = fp.p;
  TCFD_DYN_SMEM(smem_raw);
  unsigned char* area = smem_raw + ((1024u - (smem_offset(smem_raw) & 1023u)) & 1023u);
  unsigned char* tile = area;
  cx<L>* buf = reinterpret_cast<cx<L>*>(area + S::OFF_BUF);
  const cx<L>* wst = reinterpret_cast<const cx<L>*>(area + S::OFF_W);
  const cx<L>* hst = reinterpret_cast<const cx<L>*>(area + S::OFF_H);
  const cx<L>* advst = reinterpret_cast<const cx<L>*>(area + S::OFF_ADV);
  unsigned long long* bar_s = reinterpret_cast<unsigned long long*>(area + S::OFF_BAR);  // state / tile
  unsigned long long* bar_t = bar_s + 1;                                                  // tables
  unsigned long long* bar_a = bar_s + 2;                                                  // advection row
  unsigned long long* bar_o = bar_s + 3;                                                  // dw/dt reference
  volatile int* sh = reinterpret_cast<volatile int*>(bar_s + 4);                          // [0] ticket

  const int t = threadIdx.x;
  // the CONTROL thread (tickets, dependency polls, staging, release).  (Moving it to the last warp, away from
  // thread 0's extra self-conjugate entries, was measured 6 % slower: the two single-thread sections sit
  // between different barriers, so splitting them over the warps only adds waiting.)
  const bool ctl = t == 0;
  FftTwiddles<T, N> tw;
  tw.load(p.tw, t);
  CtaSync sync;
  int parity = 0;
  T kyv[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) kyv[m] = p.kappa_y[m < 4 ? t + m * NT : N - t - m * NT];
  const T ky0 = p.kappa_y[0], kyh = p.kappa_y[N / 2];
  T* kxs = reinterpret_cast<T*>(area + S::OFF_KX);
  unsigned char* frs = area + S::OFF_FR;
  for (int i = t; i < N; i += NT) {
    kxs[i] = p.kappa_x[i];
    frs[i] = p.fhat ? p.frow[i] : (unsigned char)0;
  }

  int* ticket = fp.sync;
  int* cnt_rows = fp.sync + 2;
  int* cnt_cols = fp.sync + 2 + p.B;
  const int W = fp.W;
  const int nsub = fp.nsub;
  constexpr int per_pair = IR + IC;  // items of one {cols, rows} pair per sample
  const int nchunks = (p.B + W - 1) / W;
  // items of a full chunk (the host keeps the total below 2^30: 32-bit ticket arithmetic)
  const int per_chunk = W * (IR + nsub * per_pair);
  const bool want_dwdt = p.dwdt != nullptr;

  long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, prof_last = flow_clock();  // (profiling variant only)
  int prof_cur = 0;
  int next_tk = 0;  // thread 0: the ticket of the next item (requested one item ahead)
  if (ctl) {
    stage_barrier_init(bar_s);
    stage_barrier_init(bar_t);
    stage_barrier_init(bar_a);
    stage_barrier_init(bar_o);
#ifndef TCFD_EMU
    tma_prefetch_desc(&maps.main);
    tma_prefetch_desc(&maps.adv);
#endif
    next_tk = flow_fetch_add(ticket, 1);
    sh[0] = next_tk;
    sh[2] = 0;
  }
  // items processed by this CTA: the mailbox (ticket: sh[0|1], "inputs already staged" flag: sh[2|3])
  // alternates between two slots
  int it = 0;
  unsigned tabsel = 0;  // table block (of two) the next rows unit uses
  unsigned phase_s = 0, phase_t = 0, phase_a = 0, phase_o = 0;
  // counter of the item whose stores were just ordered by the item's closing barrier, released by the
  // control thread at the top of the next iteration.  (Releasing from a thread of the other warp, so that the
  // fence overlaps thread 0's ticket / poll / staging work, was measured neutral: 974 vs 975 steps/s.)
  int* pending = nullptr;
  const bool sig = ctl;
  __syncthreads();

  for (;;) {
    // here: the previous item is complete (its stores are ordered before the last barrier, shared
    // memory is free) and sh[0] holds this item's ticket
    FLOW_MARK(0);
    const int tk = sh[it & 1];
    const bool staged = sh[2 + (it & 1)] != 0;  // thread 0 issued this item's first loads during the previous item
    ++it;
    if (sig && pending) flow_signal(pending);
    pending = nullptr;
    if (ctl) next_tk = flow_fetch_add(ticket, 1);  // consumed at the end of this item
    // ---- decode (chunks are equal-sized except the last)
    // ticket -> chunk c, its size Wc, kind, substage j (-1 = prologue), item index u inside the phase
    auto decode = [&](int tkt, int& c_, int& Wc_, bool& rows_, int& j_, int& u_) -> bool {
      c_ = nchunks == 1 ? (tkt < per_chunk ? 0 : 1) : tkt / per_chunk;
      if (c_ >= nchunks) return false;
      Wc_ = (c_ == nchunks - 1) ? p.B - c_ * W : W;
      int r = tkt - c_ * per_chunk;
      if (r >= Wc_ * (IR + nsub * per_pair)) return false;  // past the end (last chunk)
      if (r < Wc_ * IR) {
        rows_ = true; j_ = -1; u_ = r;
      } else {
        r -= Wc_ * IR;
        j_ = r / (Wc_ * per_pair);
        const int q = r % (Wc_ * per_pair);
        rows_ = q >= Wc_ * IC;
        u_ = rows_ ? q - Wc_ * IC : q;
      }
      return true;
    };
    int c, Wc, j, u;
    bool is_rows;
    if (!decode(tk, c, Wc, is_rows, j, u)) break;

    // one thread: stage the inputs of unit d of slot sl_ -- its table block always, state and advection
    // row for substage items
    auto issue_unit = [&](int sl_, int d, unsigned sel, bool prologue_, bool rd_h_) {
      unsigned char* tabdst = area + S::OFF_TAB + sel * S::TABM;
      stage_expect(bar_t, (unsigned)S::TABM);
      bulk_load(tabdst, reinterpret_cast<const unsigned char*>(p.tabU) + (size_t)d * R::TAB_BYTES, (unsigned)R::TAB_BYTES, bar_t);
      bulk_load(tabdst + R::TAB_BYTES, p.maskU + (size_t)d * 2 * R::MASK_ROW, (unsigned)(2 * R::MASK_ROW), bar_t);
      if (!prologue_) {
        const size_t ub_ = ((size_t)sl_ * ND + d) * 2 * NH;
        stage_expect(bar_s, (unsigned)R::UNIT_BYTES * (1u + (rd_h_ ? 1u : 0u)));
        bulk_load(area + S::OFF_W, fp.wU + ub_, (unsigned)R::UNIT_BYTES, bar_s);
        if (rd_h_) bulk_load(area + S::OFF_H, fp.hU + ub_, (unsigned)R::UNIT_BYTES, bar_s);
        if (d < p.NDF) {
          stage_expect(bar_a, (unsigned)(N * sizeof(cx<L>)));
          adv_row_issue<N, (int)sizeof(cx<L>)>(area + S::OFF_ADV, maps, d % ADV_BLOCK, sl_ * JB + d / ADV_BLOCK, bar_a);
        }
      }
    };
    // one thread, at a point of the current item where the staging area of the next item is free
    // (cols: tile consumed; rows: state / advection stages consumed and the other table block idle):
    // if the next ticket's dependency is ALREADY satisfied, start its first loads now.  Never waits.
    int staged_next = 0;
    auto try_stage_next = [&](bool cur_is_rows) {
      int c2, W2, j2, u2;
      bool rows2;
      if (!decode(next_tk, c2, W2, rows2, j2, u2)) return;
      if (cur_is_rows && !rows2) return;  // the tile would overwrite the table block still in use
      if (!rows2) {
        const int sl2 = u2 / IC, s2 = c2 * W + sl2;
        if (flow_ld_acquire(&cnt_rows[s2]) < IR * (j2 + 1)) return;
        flow_proxy_fence();
        tile_load_issue<N, IB>(tile, maps, (u2 % IC) * GC * 4, sl2, bar_s);
      } else {
        const int sl2 = u2 / IR, s2 = c2 * W + sl2;
        const bool pro2 = j2 < 0;
        if (pro2) {
          if (s2 >= W && flow_ld_acquire(&cnt_rows[s2 - W]) < IR * (nsub + 1)) return;
        } else {
          if (flow_ld_acquire(&cnt_cols[s2]) < IC * (j2 + 1)) return;
        }
        flow_proxy_fence();
        issue_unit(sl2, (u2 % IR) * GR, cur_is_rows ? (tabsel ^ 1u) : tabsel, pro2, !pro2 && fp.rd_h[j2 % fp.nstages]);
      }
      staged_next = 1;
    };

    if (!is_rows) {
      // ================================================================= cols item: GC quads
      FLOW_MARK(4);
      const int sl = u / IC, q0 = (u % IC) * GC;  // slot inside the chunk, first quad
      const int s = c * W + sl;
      if (ctl && !staged) {
        FLOW_DEPWAIT(&cnt_rows[s], IR * (j + 1));
        tile_load_issue<N, IB>(tile, maps, q0 * 4, sl, bar_s);
      }
#pragma unroll 1
      for (int g = 0; g < GC; ++g) {
        const int y0 = (q0 + g) * 4;
        cx<L> cc[1][8];
        FLOW_LOADWAIT(bar_s, phase_s);
        phase_s ^= 1u;
        // fp32: column 3 is read together with column 2, so the tile is free (and the next one requested) one transform
        // earlier; fp64 has no registers to spare for it
        constexpr bool PRE3 = false;  // measured: 994 vs 1018 steps/s with the early read of column 3 (12 more registers, no gain from the earlier request)
        constexpr int LASTC = PRE3 ? 2 : 3;  // the column whose transform follows the last tile read
        cx<L> z3[PRE3 ? 8 : 1];
        auto column = [&](int cidx) {
          cx<L> z[1][8];
          if (PRE3 && cidx == 3) {
#pragma unroll
            for (int m = 0; m < 8; ++m) z[0][m] = z3[PRE3 ? m : 0];
          } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) z[0][m] = tile_ld<T, G>(tile, t + m * NT, cidx);
          }
          if (PRE3 && cidx == 2) {
#pragma unroll
            for (int m = 0; m < 8; ++m) z3[PRE3 ? m : 0] = tile_ld<T, G>(tile, t + m * NT, 3);
          }
          FLOW_FFT(+1, z);
          if (cidx == LASTC) {
            // every thread has passed a barrier after its last tile read: the tile is free
            if (ctl) {
              if (g + 1 < GC) tile_load_issue<N, IB>(tile, maps, y0 + 4, sl, bar_s);
              else try_stage_next(false);
            }
          }
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const T adv = -(z[0][m].x.hi * z[0][m].x.lo + z[0][m].y.hi * z[0][m].y.lo);
            if (cidx == 0) cc[0][m].x.lo = adv;
            if (cidx == 1) cc[0][m].y.lo = adv;
            if (cidx == 2) cc[0][m].x.hi = adv;
            if (cidx == 3) cc[0][m].y.hi = adv;
          }
        };
        if constexpr ((MODE & 8) != 0) {  // experiment: one copy of the column body (instruction-cache footprint)
#pragma unroll 1
          for (int cidx = 0; cidx < 4; ++cidx) column(cidx);
        } else {
#pragma unroll
          for (int cidx = 0; cidx < 4; ++cidx) column(cidx);
        }
        FLOW_FFT(-1, cc);
        // separation buffer: with ping-pong exchanges it is the half the next exchange would write (the other
        // half may still be read by slower threads of the transform's last exchange)
        cx<L>* sbuf = buf + (PP ? parity * N : 0);
        if (PP) parity ^= 1;
#pragma unroll
        for (int m = 0; m < 8; ++m) sbuf[t + m * NT] = cc[0][m];
        __syncthreads();
        // blocked layout [slot][jj / 8][y][jj % 8] (tma.cuh): eight consecutive lanes write one 128-byte line
        cx<L>* dst = reinterpret_cast<cx<L>*>(p.advt2) + ((size_t)sl * JB * N + y0) * ADV_BLOCK;
        for (int jj = t; jj < p.NDF; jj += NT) {
          const int k = 2 * jj;
          const cx<L> c0 = sbuf[k], c1 = sbuf[k + 1], n0 = sbuf[(N - k) % N], n1 = sbuf[N - k - 1];
          const L hf(T(0.5));
          const L Ea = hf * (c0.x + n0.x), Fa = hf * (c0.y - n0.y), Ga = hf * (c0.y + n0.y), Ha = hf * (n0.x - c0.x);
          const L Eb = hf * (c1.x + n1.x), Fb = hf * (c1.y - n1.y), Gb = hf * (c1.y + n1.y), Hb = hf * (n1.x - c1.x);
          cx<L>* o = dst + (size_t)(jj / ADV_BLOCK) * N * ADV_BLOCK + (jj % ADV_BLOCK);
          o[0 * ADV_BLOCK] = cx<L>{L(Ea.lo, Eb.lo), L(Fa.lo, Fb.lo)};
          o[1 * ADV_BLOCK] = cx<L>{L(Ga.lo, Gb.lo), L(Ha.lo, Hb.lo)};
          o[2 * ADV_BLOCK] = cx<L>{L(Ea.hi, Eb.hi), L(Fa.hi, Fb.hi)};
          o[3 * ADV_BLOCK] = cx<L>{L(Ga.hi, Gb.hi), L(Ha.hi, Hb.hi)};
        }
        if (g + 1 < GC && !PP) __syncthreads();  // buf is re-used by the next quad's transforms
      }
      if (ctl) {
        sh[it & 1] = next_tk;
        sh[2 + (it & 1)] = staged_next;
      }
      pending = &cnt_cols[s];
      FLOW_MARK(7);
      __syncthreads();
      continue;
    }

    // =================================================================== rows item: up to GR units
    FLOW_MARK(5);
    const int sl = u / IR, d0 = (u % IR) * GR;
    const int gn = (ND - d0) < GR ? (ND - d0) : GR;
    const int s = c * W + sl;
    const bool prologue = j < 0;
    const int k = prologue ? 0 : j % fp.nstages;
    const bool last_sub = (j == nsub - 1);
    const bool do_inv = !last_sub;
    const bool rd_h = !prologue && fp.rd_h[k], wr_h = !prologue && fp.wr_h[k];
    const bool ld_old = last_sub && want_dwdt;  // dw/dt needs the call's input state
    const size_t sb = (size_t)s * N * NH;  // reference layout: sample base
    const T beta = fp.beta[k], gdt = fp.gdt[k], mu = fp.mu[k];

    if (ctl && !staged) {
      if (prologue) {
        if (s >= W) FLOW_DEPWAIT(&cnt_rows[s - W], IR * (nsub + 1));
      } else {
        FLOW_DEPWAIT(&cnt_cols[s], IC * (j + 1));
      }
      issue_unit(sl, d0, tabsel, prologue, rd_h);
    }

#pragma unroll 1
    for (int g = 0; g < gn; ++g) {
      FLOW_MARK(5);
      const int d = d0 + g;
      const size_t ub = ((size_t)sl * ND + d) * 2 * NH;  // unit layout: block base (slot)
      const bool valid1 = 2 * d + 1 <= N / 2;
      const int r1a = 2 * d, r2a = (N - r1a) % N;
      const int r1b = valid1 ? 2 * d + 1 : r1a, r2b = (N - r1b) % N;
      const bool selfa = r1a == r2a, selfb = r1b == r2b;
      const T kx1a = kxs[r1a], kx2a = kxs[r2a], kx1b = kxs[r1b], kx2b = kxs[r2b];
      const unsigned char* tabsrc = area + S::OFF_TAB + tabsel * S::TABM;
      const L* linst = reinterpret_cast<const L*>(tabsrc);
      const T* nilst = reinterpret_cast<const T*>(tabsrc) + 2 * NH;
      const unsigned char* maskst = tabsrc + R::TAB_BYTES;

      cx<L> wv[8], e0, e1;  // e0 = entries (r2, 0), e1 = entries (r1, N/2): owned by thread 0
      if (prologue) {
        FLOW_MARK(8);
        const cx<T>* w = p.w_in + sb;
        const int lo_a = r1a * NH + t, hi_a = r2a * NH + N - t, lo_b = r1b * NH + t, hi_b = r2b * NH + N - t;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const bool lo = m < 4;
          const cx<T> w0 = flow_ld_stream(w + (lo ? lo_a + m * NT : hi_a - m * NT)),
                      w1 = flow_ld_stream(w + (lo ? lo_b + m * NT : hi_b - m * NT));
          wv[m] = cx<L>{L(w0.x, w1.x), L(w0.y, w1.y)};
        }
        if (t == 0) {
          const cx<T> a0 = flow_ld_stream(w + r2a * NH), a1 = flow_ld_stream(w + r2b * NH);
          e0 = cx<L>{L(a0.x, a1.x), L(a0.y, a1.y)};
          const cx<T> b0 = flow_ld_stream(w + r1a * NH + N / 2), b1 = flow_ld_stream(w + r1b * NH + N / 2);
          e1 = cx<L>{L(b0.x, b1.x), L(b0.y, b1.y)};
        }
        // the table block was issued by thread 0 AFTER its dependency wait: its arrival also tells
        // every other thread that the slot's previous occupant is done (stores below re-use the slot)
        FLOW_LOADWAIT(bar_t, phase_t);
        phase_t ^= 1u;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const size_t o = ub + (m < 4 ? 0 : NH) + (m < 4 ? t + m * NT : N - t - m * NT);
          fp.wU[o] = wv[m];
          if (want_dwdt) fp.w0U[o] = wv[m];
        }
        if (t == 0) {
          fp.wU[ub + NH + 0] = e0;
          fp.wU[ub + N / 2] = e1;
          if (want_dwdt) {
            fp.w0U[ub + NH + 0] = e0;
            fp.w0U[ub + N / 2] = e1;
          }
        }
      } else {
        const bool forced = (frs[r1a] | frs[r2a] | frs[r1b] | frs[r2b]) != 0;
        cx<L> a[1][8];
        if (d < p.NDF) {
          FLOW_LOADWAIT(bar_a, phase_a);
          phase_a ^= 1u;
#pragma unroll
          for (int m = 0; m < 8; ++m) a[0][m] = advst[t + m * NT];
          FLOW_FFT(-1, a);
        } else {
#pragma unroll
          for (int m = 0; m < 8; ++m) a[0][m] = cx<L>{L(T(0)), L(T(0))};
          if (ld_old) __syncthreads();  // the ADV stage was last read by the previous unit's update
        }
        if (ld_old) {
          // the advection row is in registers (a barrier of the transform lies behind its reads): the
          // ADV stage takes this unit's block of the call's input state
          if (ctl) {
            stage_expect(bar_o, (unsigned)R::UNIT_BYTES);
            bulk_load(area + S::OFF_ADV, fp.w0U + ub, (unsigned)R::UNIT_BYTES, bar_o);
          }
        }
        FLOW_LOADWAIT(bar_t, phase_t);
        phase_t ^= 1u;
        FLOW_LOADWAIT(bar_s, phase_s);
        phase_s ^= 1u;
        if (ld_old) {
          FLOW_LOADWAIT(bar_o, phase_o);
          phase_o ^= 1u;
        }
        // RK / CN update of entry (half, col) in both lanes; returns the new w
        FLOW_MARK(9);
        auto update = [&](int half, int col, cx<L> A, bool own_a, bool own_b) -> cx<L> {
          if constexpr ((MODE & 16) != 0) {  // timing experiment: loads and stores only
            const cx<L> w_ = wst[half * NH + col];
            const cx<L> wn_ = A + w_;
            fp.wU[ub + half * NH + col] = wn_;
            if (wr_h) fp.hU[ub + half * NH + col] = A;
            return wn_;
          }
          const int ra = half ? r2a : r1a, rb = half ? r2b : r1b;
          const L lin = linst[col];
          const unsigned mk = maskst[half * R::MASK_ROW + col];
          const L f((mk & 1u) ? T(1) : T(0), (mk & 2u) ? T(1) : T(0));
          cx<L> F{f * A.x, f * A.y};
          if (forced) {
            const cx<T> f0 = p.fhat[ra * NH + col], f1 = p.fhat[rb * NH + col];
            F = F + cx<L>{L(f0.x, f1.x), L(f0.y, f1.y)};
          }
          const cx<L> w = wst[half * NH + col];
          // fused multiply-adds (explicit: the build disables contraction): 7 packed instructions fewer per entry
          cx<L> h = F;
          if (rd_h) {
            const cx<L> ho = hst[half * NH + col];
            h = cx<L>{fma_rn(ho.x, beta, F.x), fma_rn(ho.y, beta, F.y)};
          }
          if (wr_h) fp.hU[ub + half * NH + col] = h;
          const L den = fma_rn(lin, -mu, L(T(1)));
          const L inv(rcp_cn(den.lo), rcp_cn(den.hi));
          const cx<L> lw{lin * w.x, lin * w.y};
          const cx<L> x{fma_rn(lw.x, mu, fma_rn(h.x, gdt, w.x)), fma_rn(lw.y, mu, fma_rn(h.y, gdt, w.y))};
          const cx<L> wn = inv * x;
          if (last_sub) {
            if (own_a) flow_st_stream(p.w_out + sb + ra * NH + col, cx<T>{wn.x.lo, wn.y.lo});
            if (own_b) flow_st_stream(p.w_out + sb + rb * NH + col, cx<T>{wn.x.hi, wn.y.hi});
            if (want_dwdt) {
              const cx<L> o = advst[half * NH + col];
              if (own_a) flow_st_stream(p.dwdt + sb + ra * NH + col, cx<T>{p.inv_tdt * (wn.x.lo - o.x.lo), p.inv_tdt * (wn.y.lo - o.y.lo)});
              if (own_b) flow_st_stream(p.dwdt + sb + rb * NH + col, cx<T>{p.inv_tdt * (wn.x.hi - o.x.hi), p.inv_tdt * (wn.y.hi - o.y.hi)});
            }
          } else {
            fp.wU[ub + half * NH + col] = wn;
          }
          return wn;
        };
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const bool lo = m < 4;
          const int col = lo ? t + m * NT : N - t - m * NT;
          const bool first = lo || (m == 4 && t == 0);
          wv[m] = update(lo ? 0 : 1, col, lo ? a[0][m] : conj(a[0][m]), first || !selfa, valid1 && (first || !selfb));
        }
        if (t == 0) {
          e0 = update(1, 0, conj(a[0][0]), !selfa, valid1 && !selfb);
          e1 = update(0, N / 2, a[0][4], !selfa, valid1 && !selfb);
        }
      }
      // the W / H / ADV stages are consumed (and the other table block has been free since the
      // previous unit's inverse transforms): stage the next unit under this unit's inverse transforms
      FLOW_MARK(10);
      __syncthreads();
      if (ctl) {
        if (g + 1 < gn) issue_unit(sl, d + 1, tabsel ^ 1u, prologue, rd_h);
        else try_stage_next(true);
      }
      tabsel ^= 1u;

      if (do_inv) {
        FLOW_MARK(6);
        // per lane (row pair) two transforms: type 0 ("P" = A + i B) gives row r1 of z, type 1 ("Q" = A - i B,
        // stored conjugated) row r2 = N - r1 (ns2d_v2.cuh: ns_fields_z); self-paired rows need P only
        const int nq = valid1 ? 4 : 2;
#pragma unroll 1
        for (int q = 0; q < nq; ++q) {
          const int lane = q >> 1, type = q & 1;
          if (type && (lane ? selfb : selfa)) continue;  // CTA-uniform
          const T kx1 = lane ? kx1b : kx1a, kx2 = lane ? kx2b : kx2a;
          const int row = lane ? (type ? r2b : r1b) : (type ? r2a : r1a);
          const T* nl = nilst + lane * NH;
          const T sg = type ? T(-1) : T(1);
          cx<L> z[1][8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const bool lo = m < 4;
            const T nil = nl[lo ? t + m * NT : N - t - m * NT];
            if constexpr ((MODE & 4) != 0) {  // timing experiment: no field construction
              z[0][m] = wv[m];
              continue;
            }
            const cx<L> f = ns_fields_z<T>(lane_rt(wv[m], lane), nil, lo ? kx1 : kx2, kyv[m], lo ? sg : -sg);
            z[0][m] = lo ? f : conj(f);
          }
          if (t == 0 && (MODE & 4) == 0) {
            // self-conjugate columns ky = 0 and ky = N/2: Hermitian part of the two rows (C2R semantics)
            // (thread 0's generic entries m = 0 and m = 4 above ARE f(w[r1][0]) and conj f(w[r2][N/2]): only the two
            // partner entries are evaluated here -- this block runs on one lane while the other warp waits)
            const T n0 = nl[0], nh = nl[N / 2];
            const cx<L> f2_ = ns_fields_z<T>(lane_rt(e0, lane), n0, kx2, ky0, -sg);
            const cx<L> g1 = ns_fields_z<T>(lane_rt(e1, lane), nh, kx1, kyh, sg);
            z[0][0] = L(T(0.5)) * (z[0][0] + conj(f2_));
            z[0][4] = L(T(0.5)) * (g1 + z[0][4]);
          }
          FLOW_FFT(+1, z);
          cx<L>* Hrow = reinterpret_cast<cx<L>*>(p.H2) + ((size_t)sl * N + row) * (size_t)N + t;
          if (type) {
#pragma unroll
            for (int m = 0; m < 8; ++m) Hrow[m * NT] = conj(z[0][m]);
          } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) Hrow[m * NT] = z[0][m];
          }
        }
      }
    }
    if (ctl) {
      sh[it & 1] = next_tk;
      sh[2 + (it & 1)] = staged_next;
    }
    pending = &cnt_rows[s];
    FLOW_MARK(7);
    __syncthreads();
  }
  if constexpr ((MODE & 64) != 0) {
    FLOW_MARK(0);
    if (ctl && fp.prof) {
#ifndef TCFD_EMU
      for (int r = 0; r < 16; ++r) atomicAdd(fp.prof + r, (unsigned long long)prof_acc[r]);
#endif
    }
  }
}

}  // namespace tcfd
