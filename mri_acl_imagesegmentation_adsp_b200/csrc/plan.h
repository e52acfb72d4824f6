// plan.h -- host-side planners (plain C++, no CUDA calls): which columns are sampled, how the
// row pass distributes its pruned first-stage DFTs over warps, twiddle tables, radix lists.
// Shared by the product library (mriacl_recon.cu) and the CPU emulation tests.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#if !defined(__CUDACC__) && !defined(__host__)
#define __host__
#define __device__
#endif

namespace mriacl {

struct HostCf { float x, y; };

// w_N^k = exp(sign * 2 pi i k / N), computed in double, rounded once to float
inline std::vector<HostCf> make_twiddles(int n, int sign) {
  std::vector<HostCf> t(n);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < n; ++k) {
    // exact values on the axes so that quarter turns carry no rounding noise
    const double ang = two_pi * (double)k / (double)n;
    double c = std::cos(ang), s = std::sin(ang);
    if ((4 * k) % n == 0) {
      const int q = (4 * k) / n;
      c = (q == 0) ? 1.0 : (q == 2) ? -1.0 : 0.0;
      s = (q == 1) ? 1.0 : (q == 3) ? -1.0 : 0.0;
    }
    t[k].x = (float)c;
    t[k].y = (float)(sign * s);
  }
  return t;
}

// radix list for the generic Stockham kernel: 4s, then a 2, then odd primes ascending
inline std::vector<int> generic_radices(int n) {
  std::vector<int> r;
  while (n % 4 == 0) { r.push_back(4); n /= 4; }
  if (n % 2 == 0) { r.push_back(2); n /= 2; }
  for (int p = 3; (long long)p * p <= n; p += 2)
    while (n % p == 0) { r.push_back(p); n /= p; }
  if (n > 1) r.push_back(n);
  return r;
}

inline int crop_start(int n, int out) { return (n - out) / 2; }

// ---------------------------------------------------------------------------------------
// Fused plan: sampled-column list for the column pass and the warp schedule of the row pass.
// ---------------------------------------------------------------------------------------
struct FusedPlanHost {
  int H = 0, W = 0, pad_left = 0, Wp = 0, oh = 0, ow = 0, row0 = 0, col0 = 0;
  int P = 0, Q = 0;
  std::vector<int> act_w;       // physical (unpadded) column index of active column j, ascending
  std::vector<float> act_m;     // its mask value
  std::vector<int> sched;       // row-pass schedule (layout below)
  std::vector<HostCf> sptw;     // twiddles of the sparse residues: [entry][SPTW_PITCH], w_N^{n_e k1}
  std::vector<double> warp_cost;
};

// Schedule layout (int32):
//   sched[w], w < n_warps          offset of warp w's list
//   list: n_units, then per unit   n2, type, nnz, payload
//     type 1 (dense):  payload = P ints: active-column index j of n1 = 0..P-1, or -1
//     type 2 / 3 (dense, first / second half of the outputs): same payload; used when
//                      split_dense so that one dense residue is shared by two warps
//     type 0 (sparse): payload = offset of the unit's first row in sptw, then nnz column indices j;
//                      sptw row e holds w_N^{n_e k1}, k1 = 0..P-1 (n_e = Q n1 + n2 logical index)
constexpr int SPTW_PITCH = 24;   // complex elements per sptw row for P = 23 (32-row kernel)
// row pitch of the twiddle tables for a P-point first stage: P + 1 rounded up to even (16-byte rows): 23 -> 24, 31 -> 32
__host__ __device__ constexpr int sptw_pitch(int P) { return (P + 2) & ~1; }
inline void build_fused_plan(int H, int W, int pad_left, int Wp, int oh, int ow, const float* mask,
                             int P, int Q, int n_warps, int max_sparse, bool split_dense, FusedPlanHost& pl) {
  pl.H = H; pl.W = W; pl.pad_left = pad_left; pl.Wp = Wp; pl.oh = oh; pl.ow = ow;
  pl.row0 = crop_start(H, oh); pl.col0 = crop_start(Wp, ow);
  pl.P = P; pl.Q = Q;
  pl.act_w.clear(); pl.act_m.clear();
  std::vector<int> j_of_w(W, -1);
  for (int w = 0; w < W; ++w) {
    const float m = mask ? mask[w] : 1.0f;
    if (m != 0.0f) { j_of_w[w] = (int)pl.act_w.size(); pl.act_w.push_back(w); pl.act_m.push_back(m); }
  }
  struct Unit { int n2, type, nnz; std::vector<int> payload; double cost; };
  std::vector<Unit> units;
  const std::vector<HostCf> twN = make_twiddles(Wp, +1);
  pl.sptw.clear();
  for (int n2 = 0; n2 < Q; ++n2) {
    Unit u; u.n2 = n2;
    std::vector<int> jn1(P, -1);
    std::vector<int> pairs;   // (n, j)
    int nnz = 0;
    for (int n1 = 0; n1 < P; ++n1) {
      const int n = Q * n1 + n2;
      int ph = n + Wp / 2; if (ph >= Wp) ph -= Wp;        // logical -> physical (ifftshift)
      const int w = ph - pad_left;
      if (w >= 0 && w < W && j_of_w[w] >= 0) {
        jn1[n1] = j_of_w[w]; pairs.push_back(n); pairs.push_back(j_of_w[w]); ++nnz;
      }
    }
    u.nnz = nnz;
    if (nnz > max_sparse) {
      u.payload = jn1;
      const double hp = (P - 1) / 2;
      const double full = 2.0 * P + 2.2 * (P - 1) + hp * (4.0 * hp + 26.0);
      if (split_dense) {
        u.type = 2; u.cost = 0.5 * full + 30.0;
        units.push_back(u);
        u.type = 3;
        units.push_back(u);
      } else {
        u.type = 1; u.cost = full;
        units.push_back(u);
      }
    } else {
      u.type = 0; u.cost = 20.0 + P * (5.0 * nnz + 2.0);
      u.payload.clear();
      u.payload.push_back((int)pl.sptw.size());
      for (int e = 0; e < nnz; ++e) {
        const int n = pairs[2 * e];
        u.payload.push_back(pairs[2 * e + 1]);
        for (int k1 = 0; k1 < sptw_pitch(P); ++k1)
          pl.sptw.push_back(k1 < P ? twN[(size_t)(((long long)n * k1) % Wp)] : HostCf{0.f, 0.f});
      }
      units.push_back(u);
    }
  }
  // longest-processing-time-first assignment to warps
  std::vector<int> order(units.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return units[a].cost > units[b].cost; });
  std::vector<std::vector<int>> per_warp(n_warps);
  pl.warp_cost.assign(n_warps, 0.0);
  for (int ui : order) {
    int best = 0;
    for (int w = 1; w < n_warps; ++w) if (pl.warp_cost[w] < pl.warp_cost[best]) best = w;
    per_warp[best].push_back(ui);
    pl.warp_cost[best] += units[ui].cost;
  }
  pl.sched.assign(n_warps, 0);
  for (int w = 0; w < n_warps; ++w) {
    pl.sched[w] = (int)pl.sched.size();
    pl.sched.push_back((int)per_warp[w].size());
    for (int ui : per_warp[w]) {
      const Unit& u = units[ui];
      pl.sched.push_back(u.n2); pl.sched.push_back(u.type); pl.sched.push_back(u.nnz);
      pl.sched.insert(pl.sched.end(), u.payload.begin(), u.payload.end());
    }
  }
}

// ---------------------------------------------------------------------------------------
// Pair schedule for the 16-row row pass (rowpass16.cuh): the two half-warps of a warp run two units of the
// same kind (dense piece type / sparse nnz) side by side.
//   sched[w], w < n_warps   offset of warp w's list
//   list: n_pairs, then per pair: type, nnz, offA, offB   (offB = -1: second half-warp idles)
//   unit payload at offA / offB: n2, then  dense: twiddle-row offset, P column indices (n_act = the zero column);
//                                          sparse: twiddle-row offset, nnz column indices
//   sptw16 = the sparse twiddle rows followed by one row w_N^{n2 k1} per dense residue
//   an idle second half-warp (offB = -1) repeats unit A but stores into the spare residue column n2 = Q
// Built from the 32-row schedule's units (same columns, same sptw rows), so both kernels do the same arithmetic.
// The staged tile of the 16-row kernel is RESIDUE-MAJOR: a dense residue owns P consecutive slots (slot = base + n1,
// unsampled positions stay zero), the sparse residues' columns follow; `slot_of_j` maps active column j to its slot
// (used by the cp.async prefetch), so every first-stage operand sits at an immediate offset from one base.
//   dense payload: n2, twiddle-row offset, base slot        sparse payload: n2, twiddle-row offset, nnz slots
inline void build_pair_schedule(const FusedPlanHost& pl, int n_warps, std::vector<int>& out, std::vector<HostCf>& sptw16,
                                std::vector<int>* slot_of_j = nullptr, int* n_slots_out = nullptr) {
  // recover the units from the 32-row schedule
  struct U { int n2, type, nnz; std::vector<int> payload; double cost; };
  std::vector<U> units;
  const int nw32 = pl.sched.empty() ? 0 : pl.sched[0];   // first list starts right after the offset table
  for (int w = 0; w < nw32; ++w) {
    int off = pl.sched[w];
    const int n_units = pl.sched[off++];
    for (int u = 0; u < n_units; ++u) {
      U x; x.n2 = pl.sched[off]; x.type = pl.sched[off + 1]; x.nnz = pl.sched[off + 2]; off += 3;
      const int len = x.type != 0 ? pl.P : 1 + x.nnz;
      x.payload.assign(pl.sched.begin() + off, pl.sched.begin() + off + len);
      off += len;
      const double hp = (pl.P - 1) / 2;
      const double shared = 2.0 * pl.P + 2.2 * (pl.P - 1);            // loads + a_n / b_n, repeated by every part
      const double per_pair = 4.0 * hp + 26.0;
      if (x.type == 0) {
        x.cost = 20.0 + pl.P * (2.5 * x.nnz + 1.5);
        units.push_back(x);
      } else if (x.type != 3) {     // one entry per dense residue (the 32-row schedule lists halves 2 and 3)
        // 16-row kernel: thirds of the output pairs, types 2, 3, 4
        const int n0 = (int)hp / 3, n1 = ((int)hp - n0 + 1) / 2, n2 = (int)hp - n0 - n1;
        const int cnt[3] = {n0, n1, n2};
        for (int part = 0; part < 3; ++part) {
          U y = x; y.type = 2 + part; y.cost = shared + per_pair * cnt[part] + (part == 0 ? 30.0 : 0.0);
          units.push_back(y);
        }
      }
    }
  }
  std::stable_sort(units.begin(), units.end(), [](const U& a, const U& b) {
    if (a.type != b.type) return a.type > b.type;
    if (a.nnz != b.nnz) return a.nnz > b.nnz;
    return a.n2 < b.n2;
  });
  struct Pair { int a, b; double cost; };
  std::vector<Pair> pairs;
  for (size_t i = 0; i < units.size();) {
    if (i + 1 < units.size() && units[i].type == units[i + 1].type && units[i].nnz == units[i + 1].nnz) {
      pairs.push_back({(int)i, (int)i + 1, units[i].cost}); i += 2;
    } else { pairs.push_back({(int)i, -1, units[i].cost}); i += 1; }
  }
  std::vector<int> order(pairs.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pairs[a].cost > pairs[b].cost; });
  std::vector<std::vector<int>> per_warp(n_warps);
  std::vector<double> load(n_warps, 0.0);
  for (int pi : order) {
    int best = 0;
    for (int w = 1; w < n_warps; ++w) if (load[w] < load[best]) best = w;
    per_warp[best].push_back(pi);
    load[best] += pairs[pi].cost;
  }
  // twiddle rows of the dense residues: dtw[n2][k1] = w_N^{n2 k1}, appended to a copy of sptw
  sptw16 = pl.sptw;
  const std::vector<HostCf> twN = make_twiddles(pl.Wp, +1);
  std::vector<int> dense_row(pl.Q, -1);
  for (auto& u : units)
    if (u.type != 0 && dense_row[u.n2] < 0) {
      dense_row[u.n2] = (int)sptw16.size();
      for (int k1 = 0; k1 < sptw_pitch(pl.P); ++k1)
        sptw16.push_back(k1 < pl.P ? twN[(size_t)(((long long)u.n2 * k1) % pl.Wp)] : HostCf{0.f, 0.f});
    }
  out.assign(n_warps, 0);
  std::vector<std::pair<int, int>> patch;   // (position in out, unit index)
  for (int w = 0; w < n_warps; ++w) {
    out[w] = (int)out.size();
    out.push_back((int)per_warp[w].size());
    for (int pi : per_warp[w]) {
      const Pair& pr = pairs[pi];
      out.push_back(units[pr.a].type); out.push_back(units[pr.a].nnz);
      patch.push_back({(int)out.size(), pr.a}); out.push_back(0);
      if (pr.b >= 0) { patch.push_back({(int)out.size(), pr.b}); out.push_back(0); } else out.push_back(-1);
    }
  }
  // slot layout (identical for every n_warps: it only depends on the units)
  std::vector<int> slots(pl.act_w.size(), -1), dense_base(pl.Q, -1);
  int n_slots = 0;
  for (int n2 = 0; n2 < pl.Q; ++n2)
    for (auto& u : units)
      if (u.type != 0 && u.n2 == n2 && dense_base[n2] < 0) {
        dense_base[n2] = n_slots;
        for (int n1 = 0; n1 < pl.P; ++n1) if (u.payload[n1] >= 0) slots[u.payload[n1]] = n_slots + n1;
        n_slots += pl.P;
      }
  for (auto& u : units)
    if (u.type == 0)
      for (int e = 0; e < u.nnz; ++e) if (slots[u.payload[1 + e]] < 0) slots[u.payload[1 + e]] = n_slots++;
  if (slot_of_j) *slot_of_j = slots;
  if (n_slots_out) *n_slots_out = n_slots;
  std::vector<int> unit_off(units.size(), -1);
  for (auto& pu : patch) {
    if (unit_off[pu.second] < 0) {
      unit_off[pu.second] = (int)out.size();
      out.push_back(units[pu.second].n2);
      if (units[pu.second].type != 0) {
        out.push_back(dense_row[units[pu.second].n2]);
        out.push_back(dense_base[units[pu.second].n2]);
      } else {
        out.push_back(units[pu.second].payload[0]);
        for (int e = 0; e < units[pu.second].nnz; ++e) out.push_back(slots[units[pu.second].payload[1 + e]]);
      }
    }
    out[pu.first] = unit_off[pu.second];
  }
}

// ---------------------------------------------------------------------------------------
// Plan of the pair row pass (rowpair.cuh).  The staged tile is residue-major:
//   slot (n2 / step) * P + n1                 for the dense residues n2 = 0 mod step (unsampled positions stay zero),
//   nd * P + si * ne + e                      entry e of the si-th sparse residue (absent entries stay zero).
// The logical index may be rotated by `shift` (n' = n - shift mod N) so that the dense residues are the multiples of
// `step`; a rotation multiplies X[k] by a unit phase and the path only keeps magnitudes.
// Tables (float): coef[12][hp] (cos, sin)(2 pi n pair / P) | dtw[nd][12] (w_N^{n2 k1a}, w_N^{n2 k1b})
//                 | sptw[(Q - nd) * ne][12] (w_N^{n' k1a}, w_N^{n' k1b}),  k1a = pair, k1b = (P - pair) mod P.
// ---------------------------------------------------------------------------------------
struct RowPairPlanHost {
  bool ok = false;
  int step = 0, ne = 0, nd = 0, n_slots = 0, shift = 0;
  std::vector<int> slot_of_j;
  std::vector<int> zero_slots;     // dense-region slots without a column
  std::vector<float> tables;
};

inline void build_rowpair_plan(const FusedPlanHost& pl, int step, int ne, RowPairPlanHost& rp) {
  const int P = pl.P, Q = pl.Q, N = P * Q, hp = (P - 1) / 2, n_pair = (P + 1) / 2;
  rp = RowPairPlanHost();
  rp.step = step; rp.ne = ne;
  if (Q % step != 0 || N != pl.Wp || step < 2) return;
  const int nd = Q / step, nsp = Q - nd, n_act = (int)pl.act_w.size();
  rp.nd = nd; rp.n_slots = nd * P + nsp * ne;
  std::vector<int> logical(n_act);
  for (int j = 0; j < n_act; ++j) {
    int n = pl.act_w[j] + pl.pad_left - pl.Wp / 2;     // physical (padded) -> logical: inverse of the ifftshift
    if (n < 0) n += pl.Wp;
    logical[j] = n;
  }
  const double two_pi = 6.283185307179586476925286766559;
  auto tw = [&](long long e) {            // w_N^e, exact on the axes
    e %= N; if (e < 0) e += N;
    double c = std::cos(two_pi * (double)e / N), s = std::sin(two_pi * (double)e / N);
    if ((4 * e) % N == 0) { const int q = (int)((4 * e) / N); c = (q == 0) ? 1.0 : (q == 2) ? -1.0 : 0.0; s = (q == 1) ? 1.0 : (q == 3) ? -1.0 : 0.0; }
    return HostCf{(float)c, (float)s};
  };
  for (int shift = 0; shift < step && !rp.ok; ++shift) {
    std::vector<int> cnt(Q, 0), slot(n_act, -1), nprime(n_act, 0);
    std::vector<char> used(rp.n_slots, 0);
    bool fits = true;
    for (int j = 0; j < n_act && fits; ++j) {
      int np = logical[j] - shift; if (np < 0) np += N;
      nprime[j] = np;
      const int n1 = np / Q, n2 = np % Q;
      if (n2 % step == 0) slot[j] = (n2 / step) * P + n1;
      else {
        const int si = n2 - n2 / step - 1;             // index among the non-multiples of step
        if (cnt[n2] >= ne) { fits = false; break; }
        slot[j] = nd * P + si * ne + cnt[n2]++;
      }
      used[slot[j]] = 1;
    }
    if (!fits) continue;
    rp.ok = true; rp.shift = shift; rp.slot_of_j = slot;
    rp.zero_slots.clear();
    for (int sl = 0; sl < nd * P; ++sl) if (!used[sl]) rp.zero_slots.push_back(sl);
    rp.tables.assign((size_t)n_pair * hp * 2 + (size_t)nd * n_pair * 4 + (size_t)nsp * ne * n_pair * 4, 0.f);
    float* coef = rp.tables.data();
    float* dtw = coef + (size_t)n_pair * hp * 2;
    float* sptw = dtw + (size_t)nd * n_pair * 4;
    for (int pr = 0; pr < n_pair; ++pr) {
      const int k1a = pr, k1b = (P - pr) % P;
      for (int n = 1; n <= hp; ++n) {
        const double ang = two_pi * (double)((n * pr) % P) / (double)P;
        coef[(pr * hp + n - 1) * 2] = (float)std::cos(ang);
        coef[(pr * hp + n - 1) * 2 + 1] = (float)std::sin(ang);
      }
      for (int d = 0; d < nd; ++d) {
        const HostCf a = tw((long long)d * step * k1a), b = tw((long long)d * step * k1b);
        float* q = dtw + ((size_t)d * n_pair + pr) * 4;
        q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
      }
    }
    for (int j = 0; j < n_act; ++j) {
      if (slot[j] < nd * P) continue;
      const int entry = slot[j] - nd * P;
      for (int pr = 0; pr < n_pair; ++pr) {
        const int k1a = pr, k1b = (P - pr) % P;
        const HostCf a = tw((long long)nprime[j] * k1a), b = tw((long long)nprime[j] * k1b);
        float* q = sptw + ((size_t)entry * n_pair + pr) * 4;
        q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Plan of the 640-wide row pass (rowpass640.cuh): logical index n = 80 n1 + pos; for every butterfly position
// pos the sampled inputs as entries n1 | (j << 3), sorted by n1.
// ---------------------------------------------------------------------------------------
struct Row640PlanHost {
  std::vector<int> pos_off;   // [81]
  std::vector<int> ent;
  std::vector<int> perm;      // [160] first-pass position of thread tid: warps see one kind of butterfly where possible
  std::vector<int> upos;      // non-empty positions, most entries first: first-pass unit u = (upos[u / 8], line u % 8)
};

inline void build_row640_plan(const FusedPlanHost& pl, Row640PlanHost& rp) {
  const int N = 640;
  std::vector<std::vector<int>> per_pos(80);
  for (int j = 0; j < (int)pl.act_w.size(); ++j) {
    int n = pl.act_w[j] + pl.pad_left - N / 2;
    if (n < 0) n += N;
    per_pos[n % 80].push_back((n / 80) | (j << 3));
  }
  rp.pos_off.assign(81, 0);
  rp.ent.clear();
  for (int pos = 0; pos < 80; ++pos) {
    std::sort(per_pos[pos].begin(), per_pos[pos].end(), [](int a, int b) { return (a & 7) < (b & 7); });
    rp.pos_off[pos] = (int)rp.ent.size();
    rp.ent.insert(rp.ent.end(), per_pos[pos].begin(), per_pos[pos].end());
  }
  rp.pos_off[80] = (int)rp.ent.size();
  // First-pass thread -> position map.  Thread tid = 80 sub + slot works on lines sub, sub + 2, ...; its slot picks the
  // position.  A warp runs every code path its lanes need, so positions are grouped by kind (all eight inputs sampled /
  // a few / none) and the groups are started on warp boundaries where the empty positions leave room for that.
  rp.perm.assign(160, 0);
  std::vector<int> full, part, none;
  for (int pos = 0; pos < 80; ++pos) {
    const int c = rp.pos_off[pos + 1] - rp.pos_off[pos];
    (c == 8 ? full : c == 0 ? none : part).push_back(pos);
  }
  for (int sub = 0; sub < 2; ++sub) {
    std::vector<int> order(80, -1);
    std::vector<int> pad = none;
    int slot = 0;
    auto place = [&](const std::vector<int>& v) { for (int x : v) order[slot++] = x; };
    auto align = [&]() {     // advance to the next warp boundary of thread index 80 sub + slot, filling with empty positions
      while (((80 * sub + slot) & 31) != 0 && slot < 80 && !pad.empty() &&
             (int)pad.size() > 0 && slot + (int)part.size() < 80) { order[slot++] = pad.back(); pad.pop_back(); }
    };
    place(full);
    if (!full.empty() && !part.empty()) align();
    place(part);
    for (int x : pad) if (slot < 80) order[slot++] = x;
    // (if the padding ran out before a boundary the groups simply follow each other)
    std::vector<char> seen(80, 0);
    for (int i = 0; i < 80; ++i) if (order[i] >= 0) seen[order[i]] = 1;
    int fill = 0;
    for (int i = 0; i < 80; ++i) if (order[i] < 0) { while (seen[fill]) ++fill; order[i] = fill; seen[fill] = 1; }
    for (int i = 0; i < 80; ++i) rp.perm[80 * sub + i] = order[i];
  }
  // Balanced first pass (undersampled plans): only the non-empty positions are work, and there are few of them (8x mask
  // + padding: 26 of 80), so (position, line) units are dealt round-robin to ALL threads instead of a fixed position per
  // thread; sorted by entry count so that the threads of a warp run the same number of inner iterations.
  rp.upos.clear();
  for (int pos = 0; pos < 80; ++pos) if (rp.pos_off[pos + 1] > rp.pos_off[pos]) rp.upos.push_back(pos);
  std::stable_sort(rp.upos.begin(), rp.upos.end(), [&](int a, int b) {
    return rp.pos_off[a + 1] - rp.pos_off[a] > rp.pos_off[b + 1] - rp.pos_off[b];
  });
}

// FNV-1a over the plan-defining inputs: cache key for device-resident plans
inline uint64_t plan_key(const int* dims, int n_dims, const float* mask, int mask_len) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  };
  mix(dims, sizeof(int) * n_dims);
  if (mask) mix(mask, sizeof(float) * mask_len);
  else { const int none = -1; mix(&none, sizeof(none)); }
  return h;
}

}  // namespace mriacl
