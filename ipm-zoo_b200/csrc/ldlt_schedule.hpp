// ipm-zoo_b200/csrc/ldlt_schedule.hpp -- host-side task list of the persistent dataflow LDL^T.
//
// The factorization of LinearSolvers::ldlt_decomposition (LinearSolvers.cpp:14-42) is cut into
// tile tasks on a 128 x 128 tile grid (tile row i, tile column j, i >= j):
//   DIAG(k)            LDL^T of the diagonal tile (k,k)                    after k updates of (k,k)
//   TRSM(i,k,h)        64-row half h of tile (i,k):  X (D_k L_kk^T) = A    after DIAG(k) and k updates of (i,k)
//   UPD(i,j,k0,k1)     C_ij -= sum_{k0<=k<k1} L_ik D_k L_jk^T              after TRSM(i,k), TRSM(j,k), cnt(i,j)==k0
//   DIAGU(k,k0)        UPD(k,k,k0,k) and DIAG(k) fused: the last updates of a diagonal tile go straight from
//                      the accumulators into the shared-memory tile that is factored (no round trip
//                      through global memory and no task switch on the critical chain)
// The device kernel hands these out through ONE ticket counter, in the order of this list, and
// each task spins on its dependencies (flags in global memory).  Any topological order is
// deadlock-free (a waiting ticket only waits on lower tickets, all of which are held by running
// CTAs); a GOOD order keeps the spinning short.  The order here is the start order of a greedy
// list schedule simulated on `workers` SMs with a duration model of the three task kinds:
// priority goes to the left-most tile column (the critical DIAG -> TRSM -> UPD chain runs as
// far ahead of the bulk updates as the dependencies allow -- dynamic look-ahead of any depth),
// and a trailing tile is only updated once `kb` panels have accumulated for it (fewer, larger
// tasks: less C traffic), unless its column is within `la` columns of the chain front.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <queue>
#include <vector>

namespace ipmz {

constexpr int DF_TILE = 128;  // tile edge = panel width
constexpr int DF_HALF = 64;   // rows per TRSM task

enum DfType { DF_DIAG = 0, DF_TRSM = 1, DF_UPD = 2, DF_DIAGU = 3, DF_DONE = 4 };

struct DfTask {
  int type;  // DfType | half << 8
  int i, j;  // tile row, tile column (DIAG: i = j = k, TRSM: j = k)
  int k01;   // UPD: k0 | k1 << 16
};

struct DfModel {
  int workers = 148;
  double diag_us = 34.0;       // one 128 x 128 LDL^T
  double trsm_us = 12.5;       // one 64-row half
  bool fuse_diag = true;       // emit DIAGU when all remaining updates of a diagonal tile are available
  double upd_base_us = 7.0;    // C tile in/out + pipeline fill (measured: K=1 24.5 us, K=4 77 us)
  double upd_panel_us = 17.5;  // 128 x 128 x 128 on one SM at ~95 % of its DMMA rate
  int kb = 4;                  // panels accumulated before a non-urgent tile is updated
  int la = 2;                  // columns right of the chain front that are updated eagerly
  int kmax = 8;                // panels per UPD task at most (n=8192 sweep: kb 4 / kmax 8 best)
};

struct DfSchedule {
  std::vector<DfTask> tasks;
  int nt = 0;
  double makespan_us = 0.0;  // of the simulated schedule
  double work_us = 0.0;      // sum of modelled task durations
};

inline DfSchedule df_build_schedule(int N, const DfModel& m) {
  DfSchedule out;
  const int nt = (N + DF_TILE - 1) / DF_TILE;
  out.nt = nt;
  if (nt <= 0) return out;
  auto rows_of = [&](int i) { return std::min(DF_TILE, N - i * DF_TILE); };
  auto need_of = [&](int i) { return rows_of(i) > DF_HALF ? 2 : 1; };

  std::vector<int> rdy((size_t)nt * nt, 0);   // [k*nt+i] finished halves of panel tile (i,k)
  std::vector<int> cnt((size_t)nt * nt, 0);   // [i*nt+j] panels applied to tile (i,j)
  std::vector<char> busy((size_t)nt * nt, 0), queued((size_t)nt * nt, 0), final_pushed((size_t)nt * nt, 0);
  std::vector<char> diag_done(nt, 0);
  int front = -1;  // latest started DIAG

  struct Cand { uint64_t key; int kind, i, j, h; };
  auto cmp = [](const Cand& a, const Cand& b) { return a.key > b.key; };
  std::priority_queue<Cand, std::vector<Cand>, decltype(cmp)> ready(cmp);
  auto key_of = [](int col, int rank, int row) { return ((uint64_t)col << 40) | ((uint64_t)rank << 32) | (uint64_t)row; };

  auto panel_ready = [&](int k, int i) { return rdy[(size_t)k * nt + i] == need_of(i); };
  auto avail_k = [&](int i, int j) {
    int k = cnt[(size_t)i * nt + j], n = 0;
    while (k + n < j && n < m.kmax && panel_ready(k + n, i) && panel_ready(k + n, j)) ++n;
    return n;
  };
  auto upd_eligible = [&](int i, int j, int& nk) {
    if (busy[(size_t)i * nt + j]) return false;
    nk = avail_k(i, j);
    if (nk <= 0) return false;
    const int left = j - cnt[(size_t)i * nt + j];
    return nk >= std::min(std::min(m.kb, m.kmax), left) || j <= front + m.la;
  };
  // (re)examine tile (i,j): push whatever became startable
  auto touch = [&](int i, int j) {
    const size_t t = (size_t)i * nt + j;
    if (busy[t]) return;
    if (cnt[t] == j) {  // fully updated: final operation of the tile
      if (final_pushed[t]) return;
      if (i == j) {
        final_pushed[t] = 1;
        ready.push(Cand{key_of(j, 0, i), DF_DIAG, i, j, 0});
      } else if (diag_done[j]) {
        final_pushed[t] = 1;
        for (int h = 0; h < need_of(i); ++h) ready.push(Cand{key_of(j, 1, i), DF_TRSM, i, j, h});
      }
      return;
    }
    int nk;
    if (m.fuse_diag && i == j && rows_of(i) == DF_TILE && !queued[t]) {
      // every remaining panel of this diagonal tile is there: update and factor in one task
      nk = avail_k(i, j);
      if (nk > 0 && nk == j - cnt[t]) {
        queued[t] = 1;
        ready.push(Cand{key_of(j, 0, i), DF_DIAGU, i, j, 0});
        return;
      }
    }
    if (!queued[t] && upd_eligible(i, j, nk)) {
      queued[t] = 1;
      ready.push(Cand{key_of(j, 2, i), DF_UPD, i, j, 0});
    }
  };

  struct Ev { double t; int idx; };
  auto ecmp = [](const Ev& a, const Ev& b) { return a.t > b.t; };
  std::priority_queue<Ev, std::vector<Ev>, decltype(ecmp)> events(ecmp);
  std::priority_queue<double, std::vector<double>, std::greater<double>> workers;
  for (int w = 0; w < std::max(1, m.workers); ++w) workers.push(0.0);

  auto complete = [&](int idx) {
    const DfTask& tk = out.tasks[idx];
    const int type = tk.type & 0xff, i = tk.i, j = tk.j;
    if (type == DF_DIAG || type == DF_DIAGU) {
      if (type == DF_DIAGU) {
        cnt[(size_t)i * nt + j] = j;
        busy[(size_t)i * nt + j] = 0;
        final_pushed[(size_t)i * nt + j] = 1;
      }
      diag_done[j] = 1;
      rdy[(size_t)j * nt + j] = 1;
      for (int r = j + 1; r < nt; ++r) touch(r, j);
    } else if (type == DF_TRSM) {
      rdy[(size_t)j * nt + i] += 1;
      if (panel_ready(j, i)) {
        for (int c = j + 1; c <= i; ++c) touch(i, c);
        for (int r = i + 1; r < nt; ++r) touch(r, i);
      }
    } else {
      const size_t t = (size_t)i * nt + j;
      cnt[t] = tk.k01 >> 16;
      busy[t] = 0;
      touch(i, j);
    }
  };

  touch(0, 0);
  double now = 0.0;
  size_t total_final = 0;
  for (int i = 0; i < nt; ++i) total_final += 1 + (size_t)(nt - 1 - i);  // tiles; loop ends when events drain
  (void)total_final;
  for (;;) {
    const double tfree = workers.top();
    if (tfree > now) now = tfree;
    while (!events.empty() && events.top().t <= now) {
      const int idx = events.top().idx;
      events.pop();
      complete(idx);
    }
    // best startable candidate at `now`
    bool started = false;
    while (!ready.empty()) {
      const Cand c = ready.top();
      ready.pop();
      DfTask tk{};
      double dur = 0.0;
      if (c.kind == DF_UPD) {
        const size_t t = (size_t)c.i * nt + c.j;
        queued[t] = 0;
        int nk;
        if (!upd_eligible(c.i, c.j, nk)) continue;
        busy[t] = 1;
        tk = DfTask{DF_UPD, c.i, c.j, cnt[t] | ((cnt[t] + nk) << 16)};
        dur = m.upd_base_us + m.upd_panel_us * nk;
      } else if (c.kind == DF_DIAGU) {
        const size_t t = (size_t)c.i * nt + c.j;
        queued[t] = 0;
        const int nk = busy[t] ? 0 : avail_k(c.i, c.j);
        if (nk <= 0 || nk != c.j - cnt[t]) {  // stale: re-examine the tile
          touch(c.i, c.j);
          continue;
        }
        busy[t] = 1;
        tk = DfTask{DF_DIAGU, c.i, c.j, cnt[t] | (c.j << 16)};
        dur = 0.5 * m.upd_base_us + m.upd_panel_us * nk + m.diag_us;
        front = c.j;
        for (int col = front + 1; col <= std::min(nt - 1, front + m.la); ++col)
          for (int r = col; r < nt; ++r) touch(r, col);
      } else if (c.kind == DF_TRSM) {
        tk = DfTask{DF_TRSM | (c.h << 8), c.i, c.j, 0};
        dur = m.trsm_us;
      } else {
        tk = DfTask{DF_DIAG, c.i, c.j, 0};
        dur = m.diag_us;
        front = c.j;
        for (int col = front + 1; col <= std::min(nt - 1, front + m.la); ++col)
          for (int r = col; r < nt; ++r) touch(r, col);
      }
      out.tasks.push_back(tk);
      out.work_us += dur;
      events.push(Ev{now + dur, (int)out.tasks.size() - 1});
      workers.pop();
      workers.push(now + dur);
      if (now + dur > out.makespan_us) out.makespan_us = now + dur;
      started = true;
      break;
    }
    if (started) continue;
    if (events.empty()) break;  // nothing running, nothing ready: done
    now = std::max(now, events.top().t);
  }
  return out;
}

// Topological check of a task list (tests; also guards the kernel against a scheduler bug, which
// would otherwise show up as a device-side spin until the watchdog fires).
inline bool df_validate_schedule(int N, const DfSchedule& s) {
  const int nt = s.nt;
  auto rows_of = [&](int i) { return std::min(DF_TILE, N - i * DF_TILE); };
  auto need_of = [&](int i) { return rows_of(i) > DF_HALF ? 2 : 1; };
  std::vector<int> rdy((size_t)nt * nt, 0), cnt((size_t)nt * nt, 0);
  std::vector<char> fin((size_t)nt * nt, 0);
  for (const DfTask& tk : s.tasks) {
    const int type = tk.type & 0xff, i = tk.i, j = tk.j;
    if (i < j || i >= nt || j < 0) return false;
    if (type == DF_DIAG) {
      if (i != j || cnt[(size_t)i * nt + j] != j || rdy[(size_t)j * nt + j]) return false;
      rdy[(size_t)j * nt + j] = 1;
      fin[(size_t)i * nt + j] = 1;
    } else if (type == DF_TRSM) {
      if (i == j || !rdy[(size_t)j * nt + j] || cnt[(size_t)i * nt + j] != j) return false;
      if (++rdy[(size_t)j * nt + i] > need_of(i)) return false;
      fin[(size_t)i * nt + j] = 1;
    } else if (type == DF_UPD || type == DF_DIAGU) {
      const int k0 = tk.k01 & 0xffff, k1 = tk.k01 >> 16;
      if (k0 >= k1 || k1 > j || cnt[(size_t)i * nt + j] != k0) return false;
      for (int k = k0; k < k1; ++k)
        if (rdy[(size_t)k * nt + i] != need_of(i) || rdy[(size_t)k * nt + j] != need_of(j)) return false;
      cnt[(size_t)i * nt + j] = k1;
      if (type == DF_DIAGU) {
        if (i != j || k1 != j || rows_of(i) != DF_TILE || rdy[(size_t)j * nt + j]) return false;
        rdy[(size_t)j * nt + j] = 1;
        fin[(size_t)i * nt + j] = 1;
      }
    } else {
      return false;
    }
  }
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j <= i; ++j) {
      if (!fin[(size_t)i * nt + j]) return false;
      if (i != j && rdy[(size_t)j * nt + i] != need_of(i)) return false;
    }
  return true;
}

// ---- condensed assembly as UPD tasks (dataflow.cu, launch_assembly_dataflow) -------------------------------------
// Task list of K(i,j) -= A(i, :) B(j, :)^T over the lower-triangular tiles of an n x n matrix with an inner dimension
// of m: the K range of every tile is cut into `chunks` consecutive pieces, emitted chunk-major so that the pieces of
// one tile (ordered by the tile counter `cnt`) are thousands of tickets apart.
inline std::vector<DfTask> df_build_assembly_tasks(int n, int m, int chunks) {
  std::vector<DfTask> tasks;
  const int nt = (n + DF_TILE - 1) / DF_TILE, P = (m + DF_TILE - 1) / DF_TILE;
  if (nt <= 0 || P <= 0) return tasks;
  if (chunks < 1) chunks = 1;
  if (chunks > P) chunks = P;
  for (int c = 0; c < chunks; ++c) {
    const int k0 = (int)((long long)P * c / chunks), k1 = (int)((long long)P * (c + 1) / chunks);
    for (int i = 0; i < nt; ++i)
      for (int j = 0; j <= i; ++j) tasks.push_back(DfTask{DF_UPD, i, j, k0 | (k1 << 16)});
  }
  return tasks;
}

// Every lower tile receives its K range [0, P) exactly once, in increasing order (the order the kernel's tile counter
// enforces); P <= nt because the preset `rdy` flags are indexed [panel * nt + tile row].
inline bool df_validate_assembly_tasks(int n, int m, const std::vector<DfTask>& tasks) {
  const int nt = (n + DF_TILE - 1) / DF_TILE, P = (m + DF_TILE - 1) / DF_TILE;
  if (P > nt) return false;
  std::vector<int> cnt((size_t)nt * nt, 0);
  for (const DfTask& tk : tasks) {
    if ((tk.type & 0xff) != DF_UPD || tk.i < tk.j || tk.i >= nt || tk.j < 0) return false;
    const int k0 = tk.k01 & 0xffff, k1 = tk.k01 >> 16;
    if (k0 >= k1 || k1 > P || cnt[(size_t)tk.i * nt + tk.j] != k0) return false;
    cnt[(size_t)tk.i * nt + tk.j] = k1;
  }
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j <= i; ++j)
      if (cnt[(size_t)i * nt + j] != P) return false;
  return true;
}

}  // namespace ipmz
