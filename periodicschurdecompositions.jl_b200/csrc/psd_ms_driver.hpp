// Host driver of the large-N multishift periodic QR iteration (pure C++, no CUDA): the sweep
// loop, the shift strategy and the window schedule.  It talks to a backend - the CUDA kernels
// (psd_ms.cu) in the product, the CPU emulation of the same kernels in tests/ms_emul/ - so the
// control logic that decides convergence is tested on the CPU as well.
//
// Loop (replaces the outer iteration of pschur!(H1, Hs), PeriodicSchurDecompositions.jl:442-1060,
// for N >= 192):
//   scan      negligible subdiagonal entries of H_1 are set to zero; the lowest unreduced diagonal
//             block of order > W becomes the active block [ilo, ihi]
//   shifts    eigenvalues of the trailing ns x ns block of the active block (reference algorithm on
//             one CTA, eigenvalues only)
//   sweep     packets of NB bulges each are introduced at ilo every second round and chased to the
//             bottom, D rows per round, every packet inside its own diagonal window; every round
//             ends with the tensor-core updates of the off-window parts of H_j and of Z_j
//   finish    all remaining blocks (order <= W) are reduced independently by one CTA each
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "psd_ms_core.cuh"

namespace psd {
namespace ms {

struct DriverStats {
  int sweeps = 0;
  long long rounds = 0;
  long long windows = 0;     // window-rounds (chase CTAs launched)
  long long shift_pairs = 0;
  int exceptional = 0;
  int final_blocks = 0;
  double apply_flops = 0.0;  // 2 * wl^2 * (extent) summed over the tensor-core updates
};

struct DriverConfig {
  int n = 0, p = 0;
  int wantT = 1, wantZ = 1;
  int nsw = 64;      // order of the shift window (<= 64, limited by shared memory for large p)
  int rep_max = 2;   // each shift pair is used up to this many times per sweep
  int max_sweeps = 0;  // 0: 40 + 30 n / nsw
};

// status: 0 = reduced to blocks of order <= W and finished; 1 = no convergence (the factors are
// still a valid Hessenberg-triangular form with Z accumulated; the caller may fall back).
template <class Backend>
int drive(Backend& be, const DriverConfig& cfg, DriverStats& st) {
  const Geom g = geom_for(cfg.p);
  const int n = cfg.n;
  const int max_sweeps = cfg.max_sweeps > 0 ? cfg.max_sweeps : 40 + 30 * n / std::max(2, cfg.nsw);
  int last_ilo = -1, last_ihi = -1, stagnant = 0;
  std::vector<WinDesc> plan;
  std::vector<int> round_off;
  for (;;) {
    int ilo = 0, ihi = -1, done = 0, nzero = 0;
    be.scan(g.W, ilo, ihi, done, nzero);
    if (done) break;
    if (st.sweeps >= max_sweeps) return 1;
    if (ilo == last_ilo && ihi == last_ihi && nzero == 0) stagnant++; else stagnant = 0;
    last_ilo = ilo; last_ihi = ihi;
    const int m = ihi - ilo + 1;
    int ns = std::min(cfg.nsw, 2 * (m / 3));
    ns = std::max(2, ns & ~1);
    double perturb = 0.0;
    if (stagnant >= 4 && stagnant % 4 == 0) {
      perturb = 0.5;  // exceptional shifts: spread the stale set
      st.exceptional++;
    }
    if (stagnant >= 40) return 1;
    const int npairs = be.shifts(ihi - ns + 1, ns, perturb);
    if (getenv("PSD_MS_TRACE"))
      fprintf(stderr, "[psd ms] sweep %d: block [%d, %d] zeroed %d stagnant %d ns %d -> %d pairs\n", st.sweeps, ilo, ihi,
              nzero, stagnant, ns, npairs);
    if (npairs <= 0) return 1;
    // how often each pair is used: enough packets to keep the diagonal busy, at most rep_max
    const int npk1 = (npairs + g.NB - 1) / g.NB;
    int rep = std::max(1, std::min(cfg.rep_max, (m / g.W) / std::max(1, 2 * npk1)));
    const int total_pairs = npairs * rep;
    const int npk = (total_pairs + g.NB - 1) / g.NB;
    plan.clear();
    round_off.clear();
    for (int rd = 0;; rd++) {
      const size_t before = plan.size();
      for (int q = 0; q < npk; q++) {
        WinDesc w;
        if (packet_window(g, ilo, ihi, total_pairs, q, rd, w)) {
          w.npairs = npairs;
          plan.push_back(w);
        }
      }
      if (plan.size() == before) {
        if (rd >= 2 * (npk - 1)) break;  // every packet has been introduced and has left
        round_off.push_back((int)before);  // an empty round between introductions
        continue;
      }
      round_off.push_back((int)before);
    }
    round_off.push_back((int)plan.size());
    be.upload_plan(plan);
    for (size_t rd = 0; rd + 1 < round_off.size(); rd++) {
      const int off = round_off[rd], cnt = round_off[rd + 1] - off;
      if (cnt == 0) continue;
      be.round(off, cnt);
      st.rounds++;
      st.windows += cnt;
      for (int i = 0; i < cnt; i++) {
        const WinDesc& w = plan[off + i];
        const double left = (cfg.wantT ? n : w.ihi + 1) - (w.s + w.wl);
        const double right = w.s - (cfg.wantT ? 0 : w.ilo);
        const double z = cfg.wantZ ? n : 0;
        st.apply_flops += 2.0 * w.wl * w.wl * (left + right + z) * cfg.p;
      }
    }
    st.sweeps++;
    st.shift_pairs += total_pairs;
  }
  int nblocks = 0;
  be.finish(nblocks);
  st.final_blocks = nblocks;
  return 0;
}

}  // namespace ms
}  // namespace psd
