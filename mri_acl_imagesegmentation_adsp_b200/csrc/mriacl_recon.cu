// mriacl_recon.cu -- C ABI of libmriacl_recon.so (see include/mriacl_recon.h) and the host
// orchestration of the kernels.  Built for sm_100a only; no cuFFT, no CPU fallback.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mriacl_recon.h"
#include "common.cuh"
#include "rt.h"
#include "plan.h"
#include "hostpack.h"
#include "generic_kernels.cuh"
#include "colpass640.cuh"
#include "rowpass.cuh"
#include "rowpass16.cuh"
#include "rowpass640.cuh"
#include "rowpass_generic.cuh"
#ifndef MRIACL_EMU
#include "post_kernels.cuh"
#include "grappa_kernels.cuh"
#endif
#ifdef MRIACL_EXPERIMENTAL
// schedules that were built, measured and found slower than `sequential`: kept as negative
// results for the emulator tests and for A/B runs, not part of the product library
#include "rowpair.cuh"
#include "fused640x368.cuh"
#include "colpass640_tma.cuh"
#include "coresident640x368.cuh"
#endif

using namespace mriacl;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

std::mutex g_mu;

// ---- per-device facts ----------------------------------------------------------------
struct DeviceInfo { int sms = 0; bool smem_set = false; };
std::map<int, DeviceInfo> g_dev;

int device_sms(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.sms) d.sms = rt_sm_count(dev);
  return d.sms;
}

#ifdef MRIACL_EXPERIMENTAL
int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#endif

constexpr int FUSED_P = 23, FUSED_Q = 16;      // 368-wide knee plans
constexpr int W372_P = 31, W372_Q = 12;        // 372-wide knee plans: same row-pass kernel, 31-point first stage, 12-point second
constexpr int W372_NW = 8;                     // 16 output pairs over 8 warps (two pairs each), 2 CTAs per SM
constexpr int W400_P = 25, W400_Q = 16;        // 400-wide knee plans: 25-point first stage (symmetric direct DFT, any odd length)
constexpr int W400_NW = 7;                     // 13 output pairs over 7 warps
constexpr int SMEM_MAX = 227 * 1024 - 512;     // dynamic shared memory a B200 CTA may opt in to (227 KB minus the kernels' static part)
// Row-pass CTA shapes: 16 warps (one CTA owns the SM) for the sequential schedule, 8 warps for the
// overlapped schedule, where one row-pass CTA shares each SM with column-pass CTAs.
constexpr int RP_NW_SEQ = 16;
#ifdef MRIACL_EXPERIMENTAL
constexpr int RP_NW_OVL = 8;
#endif
// Pair row pass (rowpair.cuh): dense residues every RPP_STEP-th, at most RPP_NE extra columns per other residue.
// That is the 4x-equispaced + low-frequency-block mask family; other masks keep the cooperative row pass.
#ifdef MRIACL_EXPERIMENTAL
constexpr int RPP_STEP = 4, RPP_NE = 2;
#endif



constexpr int GEN_SMEM_BYTES = 2 * MRIACL_GEN_SMEM_ELEMS * 8;

int ensure_smem_attrs(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceInfo& d = g_dev[dev];
  if (d.smem_set) return 0;
  int bad = 0;
  bad |= rt_allow_smem((const void*)generic_fft_kernel, GEN_SMEM_BYTES);
  bad |= rt_allow_smem((const void*)colpass640_ws_kernel, CP_SMEM_BYTES_WS);   // default carveout: 2 CTAs x 46 KB -> 100 KB, 156 KB of L1 left for the gather
  bad |= rt_allow_smem((const void*)rowpass16_kernel<FUSED_P, FUSED_Q, 12, 2>, SMEM_MAX / 2);
  bad |= rt_allow_smem((const void*)rowpass16_kernel<W372_P, W372_Q, W372_NW, 2>, SMEM_MAX / 2);
  bad |= rt_allow_smem((const void*)rowpass16_kernel<W400_P, W400_Q, W400_NW, 2>, SMEM_MAX / 2);
  bad |= rt_allow_smem((const void*)rowpass640_kernel<false>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)rowpass640_kernel<true>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)rowpass_generic_kernel, SMEM_MAX);
#ifdef MRIACL_EXPERIMENTAL
  const int cp_carve = env_int("MRIACL_CP_CARVEOUT", -1);
  if (cp_carve >= 0) bad |= rt_allow_smem((const void*)colpass640_ws_kernel, CP_SMEM_BYTES_WS, cp_carve);
  bad |= rt_allow_smem((const void*)colpass640_ws_g_kernel<8, 2>, CP_SMEM_BYTES_DB, cp_carve);
  bad |= rt_allow_smem((const void*)colpass640_kernel<true>, CP_SMEM_BYTES_DB, cp_carve);
  bad |= rt_allow_smem((const void*)colpass640_kernel<false>, CP_SMEM_BYTES_SB, 100);   // co-resident with rowpass<8>: same carveout
  bad |= rt_allow_smem((const void*)rowpass_kernel<FUSED_P, FUSED_Q, RP_NW_SEQ>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)rowpass_kernel<FUSED_P, FUSED_Q, RP_NW_OVL>, SMEM_MAX, 100);
  bad |= rt_allow_smem((const void*)fused640_kernel<FUSED_P, FUSED_Q>, SMEM_MAX / 2);
  bad |= rt_allow_smem((const void*)rowpass16_kernel<FUSED_P, FUSED_Q, 12, 1>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)rowpass16_kernel<FUSED_P, FUSED_Q, 16, 1>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)rowpass16_kernel<FUSED_P, FUSED_Q, 8, 3>, SMEM_MAX / 3);
  bad |= rt_allow_smem((const void*)knee_coresident_kernel<FUSED_P, FUSED_Q>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)knee_coresident_split_kernel<FUSED_P, FUSED_Q>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)knee_coresident_pair_kernel<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE>, SMEM_MAX);
  bad |= rt_allow_smem((const void*)rowpair_kernel<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE, 1>, SMEM_MAX / 2);
  bad |= rt_allow_smem((const void*)rowpair_kernel<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE, 2>, SMEM_MAX / 2);
#endif
  if (!bad) d.smem_set = true;
  return bad;
}

// ---- device-resident tables, built once and cached (every cache is bounded: a data loader that draws a new random
// mask or meets a new width per volume must not leak device memory) --------------------------------------------------
struct DeviceBuf {
  void* p = nullptr;
  DeviceBuf() = default;
  DeviceBuf(const DeviceBuf&) = delete;
  DeviceBuf& operator=(const DeviceBuf&) = delete;
  ~DeviceBuf() { if (p) rt_free(p); }      // rt_free synchronises the device: no kernel of an earlier call still reads it
};
typedef std::shared_ptr<DeviceBuf> DevPtr;

// allocate + upload `n` elements; an empty vector gets a one-element allocation so that the pointer is never null
template <class T> DevPtr dev_upload(const T* host, size_t n) {
  auto b = std::make_shared<DeviceBuf>();
  if (rt_malloc(&b->p, sizeof(T) * std::max<size_t>(1, n))) return nullptr;
  if (n && rt_upload(b->p, host, sizeof(T) * n)) return nullptr;
  return b;
}
template <class T> DevPtr dev_upload(const std::vector<T>& v) { return dev_upload(v.data(), v.size()); }

uint64_t g_clock = 0;                      // LRU stamps (under g_mu)
constexpr size_t MAX_CACHED_TWIDDLES = 64, MAX_CACHED_MASKS = 64, MAX_CACHED_PLANS = 64;

// forward twiddles w_N^k = exp(-2 pi i k/N) for the generic kernel, and inverse-sign tables
// for the fused kernels; key = (device, N, sign)
struct TwEntry { DevPtr buf; uint64_t stamp; };
std::map<std::tuple<int, int, int>, TwEntry> g_tw;

DevPtr get_twiddles(int dev, int n, int sign) {
  auto key = std::make_tuple(dev, n, sign);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) { it->second.stamp = ++g_clock; return it->second.buf; }
  }
  std::vector<HostCf> t = make_twiddles(n, sign);
  DevPtr d = dev_upload(t);
  if (!d) return nullptr;
  DevPtr evicted;                          // released after the lock is dropped (cudaFree synchronises)
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) { it->second.stamp = ++g_clock; return it->second.buf; }   // another thread was faster
    g_tw[key] = TwEntry{d, ++g_clock};
    if (g_tw.size() > MAX_CACHED_TWIDDLES) {
      auto victim = g_tw.end();
      for (auto jt = g_tw.begin(); jt != g_tw.end(); ++jt)
        if (jt->first != key && (victim == g_tw.end() || jt->second.stamp < victim->second.stamp)) victim = jt;
      if (victim != g_tw.end()) { evicted = victim->second.buf; g_tw.erase(victim); }
    }
  }
  return d;
}

struct MaskDev { std::vector<float> host; DevPtr dev; int device; uint64_t stamp; };
std::vector<MaskDev> g_masks;

// device copy of a host mask (generic path); nullptr mask -> nullptr.  `keep` holds the buffer for the caller.
int get_device_mask(int dev, const float* mask, int w, const float** out, DevPtr& keep) {
  *out = nullptr;
  if (!mask) return 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& m : g_masks)
      if (m.device == dev && (int)m.host.size() == w && !memcmp(m.host.data(), mask, sizeof(float) * w)) {
        m.stamp = ++g_clock; keep = m.dev; *out = (const float*)keep->p; return 0;
      }
  }
  DevPtr d = dev_upload(mask, (size_t)w);
  if (!d) return 1;
  DevPtr evicted;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_masks.push_back(MaskDev{std::vector<float>(mask, mask + w), d, dev, ++g_clock});
    if (g_masks.size() > MAX_CACHED_MASKS) {
      size_t victim = 0;
      for (size_t i = 1; i + 1 < g_masks.size(); ++i) if (g_masks[i].stamp < g_masks[victim].stamp) victim = i;
      evicted = g_masks[victim].dev;
      g_masks.erase(g_masks.begin() + (long)victim);
    }
  }
  keep = d; *out = (const float*)d->p;
  return 0;
}

struct FusedPlanDev {
  FusedPlanHost host;            // column list + the 32-row schedule the pair schedules are derived from
  std::vector<int> pairs12, pairs8;   // 16-row kernel: pair schedules for 12 and 8 (7 for 400-wide plans) warps
  std::vector<HostCf> sptw16;
  std::vector<int> rp16_slot_of_j;     // residue-major staged tile of the 16-row row pass
  int rp16_slots = 0;
  Row640PlanHost r640;             // 640-wide row pass (Wp == 640 plans)
  std::vector<float> mask_copy;
  bool has_mask = false;
  bool unit_mask = true;
  // device tables: owned through the DevPtrs below, released when the plan is evicted from the cache and the last
  // call using it has returned
  std::vector<DevPtr> owned;
  int* act_w = nullptr; float* act_m = nullptr;
  int* act_ident = nullptr;        // 0 .. n_act-1: the column list of packed k-space (MRIACL_PACKED_COLUMNS)
  std::vector<int> band_of_item, j0_of_item;   // TMA column pass: bands of 8 raw columns holding a sampled column, and each
  int* item_band = nullptr; int* item_j0 = nullptr;   //            band's first active column (j0_of_item has one entry more)
  int* sched_p8 = nullptr; int* sched_p12 = nullptr; cf* sptw16_dev = nullptr; int* rp16_slot_dev = nullptr;
  int* act_logical = nullptr;      // pruned generic row pass: logical index of active column j in the padded line
  int* r640_off = nullptr; int* r640_ent = nullptr; int* r640_perm = nullptr; int* r640_upos = nullptr;
  cf* twH = nullptr; cf* twW = nullptr; cf* twW_fwd = nullptr;     // (shared twiddle cache entries, kept alive by `owned`)
#ifdef MRIACL_EXPERIMENTAL
  FusedPlanHost host_ovl;        // schedule for RP_NW_OVL warps (same columns, same sptw)
  std::vector<int> pairs16;
  RowPairPlanHost rpp;           // pair row pass (ok = the mask fits its template)
  int* sched = nullptr; cf* sptw = nullptr; int* sched_ovl = nullptr; int* sched_p16 = nullptr;
  int* rpp_slot = nullptr; int* rpp_zero = nullptr; float* rpp_tab = nullptr;
#endif
  uint64_t stamp = 0;              // last use, for the cache bound
  template <class T, class U> bool put(const std::vector<U>& v, T*& member) {
    DevPtr d = dev_upload(v);
    if (!d) return false;
    owned.push_back(d);            // ownership first: an early return never leaks
    member = (T*)d->p;
    return true;
  }
  bool hold(const DevPtr& d, cf*& member) { if (!d) return false; owned.push_back(d); member = (cf*)d->p; return true; }
};
typedef std::shared_ptr<FusedPlanDev> PlanPtr;
std::map<std::pair<int, uint64_t>, std::vector<PlanPtr>> g_fused;
size_t g_plan_count = 0;

bool plan_matches(const FusedPlanDev& pl, int H, int W, int pad_left, int Wp, int oh, int ow, const float* mask, bool device_side) {
  const FusedPlanHost& h = pl.host;
  return h.H == H && h.W == W && h.pad_left == pad_left && h.Wp == Wp && h.oh == oh && h.ow == ow &&
         pl.has_mask == (mask != nullptr) && (!mask || !memcmp(pl.mask_copy.data(), mask, sizeof(float) * W)) &&
         (pl.act_w || !device_side);
}

PlanPtr get_fused_plan(int dev, int H, int W, int pad_left, int Wp, int oh, int ow, const float* mask, bool device_side) {
  const int dims[6] = {H, W, pad_left, Wp, oh, ow};
  const uint64_t key = plan_key(dims, 6, mask, W);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& pl : g_fused[{dev, key}])
      if (plan_matches(*pl, H, W, pad_left, Wp, oh, ow, mask, device_side)) { pl->stamp = ++g_clock; return pl; }
  }
  auto pl = std::make_shared<FusedPlanDev>();
  const int planP = Wp == W372_P * W372_Q ? W372_P : Wp == W400_P * W400_Q ? W400_P : FUSED_P;
  const int planQ = Wp == W372_P * W372_Q ? W372_Q : Wp == W400_P * W400_Q ? W400_Q : FUSED_Q;
  build_fused_plan(H, W, pad_left, Wp, oh, ow, mask, planP, planQ, RP_NW_SEQ, RP_MAX_SPARSE, /*split_dense=*/true, pl->host);
  pl->has_mask = mask != nullptr;
  if (mask) pl->mask_copy.assign(mask, mask + W);
  if (device_side) {
    const FusedPlanHost& h = pl->host;
    if (!pl->put(h.act_w, pl->act_w) || !pl->put(h.act_m, pl->act_m)) return nullptr;
    {
      std::vector<int> ident(h.act_w.size());
      for (size_t j = 0; j < ident.size(); ++j) ident[j] = (int)j;
      if (!pl->put(ident, pl->act_ident)) return nullptr;
    }
#if defined(MRIACL_EXPERIMENTAL) && !defined(MRIACL_EMU)
    for (size_t j = 0; j < h.act_w.size(); ++j) {      // TMA column pass: bands of CT_BW raw columns that hold a sampled column
      const int band = h.act_w[j] / CT_BW;
      if (pl->band_of_item.empty() || pl->band_of_item.back() != band) { pl->band_of_item.push_back(band); pl->j0_of_item.push_back((int)j); }
    }
    pl->j0_of_item.push_back((int)h.act_w.size());
    if (!pl->band_of_item.empty() && (!pl->put(pl->band_of_item, pl->item_band) || !pl->put(pl->j0_of_item, pl->item_j0))) return nullptr;
#endif
    build_pair_schedule(pl->host, 12, pl->pairs12, pl->sptw16, &pl->rp16_slot_of_j, &pl->rp16_slots);
    build_pair_schedule(pl->host, Wp == W400_P * W400_Q ? W400_NW : 8, pl->pairs8, pl->sptw16);   // (8 warps; 7 for the 400-wide plans)
    if (!pl->put(pl->rp16_slot_of_j, pl->rp16_slot_dev) || !pl->put(pl->sptw16, pl->sptw16_dev) ||
        !pl->put(pl->pairs12, pl->sched_p12) || !pl->put(pl->pairs8, pl->sched_p8)) return nullptr;
    if (Wp != CP_N && Wp != FUSED_P * FUSED_Q) {      // (372 too: its dense plans fall back to the pruned generic row pass)
      std::vector<int> lg(std::max<size_t>(1, h.act_w.size()), 0);
      for (size_t j = 0; j < h.act_w.size(); ++j) lg[j] = logical_of_phys(h.act_w[j] + pad_left, Wp);
      if (!pl->put(lg, pl->act_logical) || !pl->hold(get_twiddles(dev, Wp, -1), pl->twW_fwd)) return nullptr;
    }
    if (Wp == CP_N) {
      build_row640_plan(pl->host, pl->r640);
      if (!pl->put(pl->r640.pos_off, pl->r640_off) || !pl->put(pl->r640.ent, pl->r640_ent) ||
          !pl->put(pl->r640.perm, pl->r640_perm)) return nullptr;
      if (!pl->r640.upos.empty() && !pl->put(pl->r640.upos, pl->r640_upos)) return nullptr;
    }
#ifdef MRIACL_EXPERIMENTAL
    build_fused_plan(H, W, pad_left, Wp, oh, ow, mask, planP, planQ, RP_NW_OVL, RP_MAX_SPARSE, /*split_dense=*/true, pl->host_ovl);
    build_pair_schedule(pl->host, 16, pl->pairs16, pl->sptw16);
    if (!pl->put(h.sched, pl->sched) || !pl->put(h.sptw, pl->sptw) || !pl->put(pl->host_ovl.sched, pl->sched_ovl) ||
        !pl->put(pl->pairs16, pl->sched_p16)) return nullptr;
    build_rowpair_plan(pl->host, RPP_STEP, RPP_NE, pl->rpp);
    if (pl->rpp.ok && (!pl->put(pl->rpp.slot_of_j, pl->rpp_slot) || !pl->put(pl->rpp.zero_slots, pl->rpp_zero) ||
                       !pl->put(pl->rpp.tables, pl->rpp_tab))) return nullptr;
#endif
    for (float v : h.act_m) if (v != 1.0f) pl->unit_mask = false;
    if (!pl->hold(get_twiddles(dev, H, +1), pl->twH) || !pl->hold(get_twiddles(dev, Wp, +1), pl->twW)) return nullptr;
  }
  PlanPtr evicted;                           // destroyed (cudaFree, device synchronisation) after the lock is dropped
  std::lock_guard<std::mutex> lk(g_mu);
  auto& bucket = g_fused[{dev, key}];
  for (auto& other : bucket)                 // two threads missed on the same key: keep the first plan
    if (plan_matches(*other, H, W, pad_left, Wp, oh, ow, mask, device_side)) { other->stamp = ++g_clock; evicted = pl; return other; }
  pl->stamp = ++g_clock;
  bucket.push_back(pl);
  if (++g_plan_count > MAX_CACHED_PLANS) {   // drop the least recently used plan (the caller's PlanPtr keeps its own alive)
    std::vector<PlanPtr>* best_bucket = nullptr;
    size_t best_i = 0;
    uint64_t best = ~0ull;
    for (auto& kv : g_fused)
      for (size_t i = 0; i < kv.second.size(); ++i)
        if (kv.second[i] != pl && kv.second[i]->stamp < best) { best = kv.second[i]->stamp; best_bucket = &kv.second; best_i = i; }
    if (best_bucket) { evicted = (*best_bucket)[best_i]; best_bucket->erase(best_bucket->begin() + (long)best_i); --g_plan_count; }
  }
  return pl;
}

bool fused_shape(int H, int Wp) {
  return H == CP_N && (Wp == FUSED_P * FUSED_Q || Wp == W372_P * W372_Q || Wp == W400_P * W400_Q || Wp == CP_N);
}
// H = 640 with any other width: the fused column pass feeds the pruned generic row pass (rowpass_generic.cuh)
bool pruned_shape(int H, int Wp) { return H == CP_N && !fused_shape(H, Wp) && Wp <= MRIACL_MAX_LINE; }
// ... provided one line of it fits shared memory: two line buffers + the per-stage index tables + the output tiles
// (Wp = 3072 or 4096 do not: those shapes take the generic kernels, which hold lines up to MRIACL_MAX_LINE)
bool pruned_fits(int Wp, int ow) {
  const int stages = (int)generic_radices(Wp).size();
  return stages <= RG_MAX_STAGES && rowgen_smem_bytes(Wp, 1, ow, stages) <= SMEM_MAX;
}

struct ReconGeom {
  bool fused;
  int n_act, n_tiles;   // n_tiles: 32-row tiles (T row pitch = 32 * n_tiles)
  int n_tiles16, n_tiles8;
  size_t per_slice;     // workspace bytes per slice in flight
  size_t t_bytes;       // intermediate bytes per slice
};

int recon_geom(int A, int C, int H, int W, int pad_left, int Wp, int oh, int ow, const float* mask,
               unsigned flags, ReconGeom& g) {
  g.fused = (fused_shape(H, Wp) || (pruned_shape(H, Wp) && pruned_fits(Wp, ow))) && !(flags & MRIACL_FORCE_GENERIC);
  g.n_tiles = (oh + RP_ROWS - 1) / RP_ROWS;
  g.n_tiles16 = (oh + RP16_ROWS - 1) / RP16_ROWS;
  g.n_tiles8 = (oh + R640_ROWS - 1) / R640_ROWS;
  if (g.fused) {
    int n_act = 0;
    for (int w = 0; w < W; ++w) n_act += (!mask || mask[w] != 0.0f) ? 1 : 0;
    g.n_act = n_act;
    const size_t ohp = (size_t)g.n_tiles * RP_ROWS;
    g.t_bytes = align_up((size_t)A * C * (size_t)(n_act > 0 ? n_act : 1) * ohp * sizeof(cf), 256);
    g.per_slice = g.t_bytes + align_up((size_t)oh * 3 * sizeof(float), 256) + 256;   // partials (finest tiling: one row per tile) + completion counter
  } else {
    g.n_act = W;
    g.t_bytes = align_up((size_t)A * C * H * Wp * sizeof(cf), 256);
    g.per_slice = g.t_bytes;
  }
  return 0;
}

int validate_recon(int B, int A, int C, int H, int W, int pad_left, int Wp, int oh, int ow) {
  if (B < 0 || A < 1 || C < 1 || H < 1 || W < 1) return fail(MRIACL_ERR_INVALID, "bad dims B=%d A=%d C=%d H=%d W=%d", B, A, C, H, W);
  if (pad_left < 0 || Wp < W + pad_left) return fail(MRIACL_ERR_INVALID, "W_padded=%d < W=%d + pad_left=%d", Wp, W, pad_left);
  if (oh < 1 || ow < 1 || oh > H || ow > Wp) return fail(MRIACL_ERR_INVALID, "Invalid shapes: crop %dx%d of %dx%d", oh, ow, H, Wp);
  if (H > MRIACL_MAX_LINE || Wp > MRIACL_MAX_LINE) return fail(MRIACL_ERR_UNSUPPORTED, "line length above %d", MRIACL_MAX_LINE);
  return 0;
}

// one generic centred 1-D pass over [n_frames] frames
int launch_generic_pass(GenFftParams gp, int dev, rt_stream_t st) {
  std::vector<int> rad = generic_radices(gp.N);
  if ((int)rad.size() > MRIACL_GEN_MAX_STAGES) return fail(MRIACL_ERR_UNSUPPORTED, "too many FFT stages for N=%d", gp.N);
  gp.n_stages = (int)rad.size();
  for (int i = 0; i < gp.n_stages; ++i) gp.radix[i] = rad[i];
  const DevPtr tw = get_twiddles(dev, gp.N, -1);     // (an evicted table is freed with a device synchronisation, after this launch)
  if (!tw) return fail(MRIACL_ERR_CUDA, "twiddle table allocation failed: %s", rt_last_error_string());
  gp.tw = (const cf*)tw->p;
  int L = MRIACL_GEN_SMEM_ELEMS / gp.N;
  L = L < 1 ? 1 : (L > 8 ? 8 : L);
  if (L > gp.lines_per_frame) L = gp.lines_per_frame;
  gp.lines_per_block = L;
  const int blocks = ((gp.lines_per_frame + L - 1) / L) * gp.n_frames;
  if (blocks == 0) return 0;
  const size_t smem = (size_t)2 * L * gp.N * sizeof(cf);
  MRIACL_LAUNCH(generic_fft_kernel, blocks, MRIACL_GEN_THREADS, smem, st, gp);
  return 0;
}

// centred 2-D transform of n_frames frames [H][W(+pad)] -> out [n_frames][H][Wp]
int generic_fft2c(const cf* in, long long sb, long long sa, int A, int C, cf* out, int n_frames, int H, int W,
                  int pad_left, int Wp, const float* mask_dev, int inverse, int dev, rt_stream_t st) {
  GenFftParams r{};
  r.in = in; r.out = out; r.mask = mask_dev;
  r.N = Wp; r.in_len = W; r.in_pad = pad_left;
  r.lines_per_frame = H; r.n_frames = n_frames; r.lines_contig = 0;
  r.in_es = 1; r.in_ls = W; r.in_sb = sb; r.in_sa = sa; r.in_sc = (long long)H * W; r.A = A; r.C = C;
  r.out_es = 1; r.out_ls = Wp; r.out_fs = (long long)H * Wp;
  r.inverse = inverse; r.scale = (float)(1.0 / std::sqrt((double)Wp));
  if (int rc = launch_generic_pass(r, dev, st)) return rc;
  GenFftParams c{};
  c.in = out; c.out = out; c.mask = nullptr;
  c.N = H; c.in_len = H; c.in_pad = 0;
  c.lines_per_frame = Wp; c.n_frames = n_frames; c.lines_contig = 1;
  c.in_es = Wp; c.in_ls = 1; c.in_sb = (long long)H * Wp; c.in_sa = 0; c.in_sc = 0; c.A = 1; c.C = 1;
  c.out_es = Wp; c.out_ls = 1; c.out_fs = (long long)H * Wp;
  c.inverse = inverse; c.scale = (float)(1.0 / std::sqrt((double)H));
  return launch_generic_pass(c, dev, st);
}

#if defined(MRIACL_EXPERIMENTAL) && !defined(MRIACL_EMU)
// ---- TMA descriptor of a k-space batch: (2 W floats) x (C H rows) x A x B, boxes of one band x 128 rows ---------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (TensorMapEncodeFn)p;
  }();
  return fn;
}
// false: this batch cannot be described (alignment, strides) -- the caller takes the cp.async gather
bool kspace_tensor_map(CUtensorMap* map, const cf* ksp, long long sb, long long sa, int B, int A, int C, int H, int W) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc || H != CP_N || (W & 1) || ((uintptr_t)ksp & 15)) return false;
  const long long frame = (long long)C * H * W;
  if (A == 1) sa = frame;
  if (B == 1) sb = frame * A;
  if (sa <= 0 || sb <= 0 || (sa & 1) || (sb & 1)) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)C * H, (cuuint64_t)A, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 8, (cuuint64_t)sa * 8, (cuuint64_t)sb * 8};
  const cuuint32_t box[4] = {2 * CT_BW, CT_BOX_ROWS, 1, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ksp, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

struct FusedArgs {
  const cf* ksp; long long slice_stride, avg_stride; const float* mask; float* out; float* mean_std;
  int B, A, C, H, W, pad_left, Wp, oh, ow; unsigned flags; float eps;
  void* workspace; size_t workspace_bytes; rt_stream_t st; int dev, sms;
};

int run_fused640(const FusedArgs& a, const ReconGeom& g);
#ifdef MRIACL_EXPERIMENTAL
int run_fused_experimental(const FusedArgs& a, const ReconGeom& g);
#endif

// The 640 x 372 and 640 x 400 knee plans: column pass -> 16-row row pass with a P-point first stage and a Q-point second
// stage (372 = 31 x 12, 400 = 25 x 16; same kernel template and plan builder as the 368-wide plans) -> normalise.
template <int PP, int QQ, int NWW>
int run_fused_pq(const FusedArgs& a, const ReconGeom& g) {
  PlanPtr pl = get_fused_plan(a.dev, a.H, a.W, a.pad_left, a.Wp, a.oh, a.ow, a.mask, true);
  if (!pl) return fail(MRIACL_ERR_CUDA, "plan upload failed: %s", rt_last_error_string());
  const int n_act = (int)pl->host.act_w.size();
  const int n_groups = (n_act + CP_GW - 1) / CP_GW;
  const int ohp = g.n_tiles * RP_ROWS;
  const int row0 = crop_start(a.H, a.oh), col0 = crop_start(a.Wp, a.ow);
  const bool want_norm = (a.flags & MRIACL_NORM_INSTANCE) != 0;
  const int flip = (a.flags & MRIACL_FLIP_ROWS) ? 1 : 0;
  const int chunk = (int)std::min<size_t>((size_t)a.B, a.workspace_bytes / g.per_slice);
  int n_buf372 = 2;
  int smem16 = rowpass16_smem_bytes(PP, QQ, (int)pl->sptw16.size(), (int)pl->pairs8.size(), pl->rp16_slots, n_buf372, a.ow, a.A);
  if (smem16 > SMEM_MAX / 2) { n_buf372 = 1; smem16 = rowpass16_smem_bytes(PP, QQ, (int)pl->sptw16.size(), (int)pl->pairs8.size(), pl->rp16_slots, 1, a.ow, a.A); }
  if (smem16 > SMEM_MAX / 2) return run_fused640(a, g);     // (nearly) fully sampled: the staged tile is too large, take the pruned generic row pass
  for (int s0 = 0; s0 < a.B; s0 += chunk) {
    const int ns = std::min(chunk, a.B - s0);
    char* base = (char*)a.workspace;
    cf* T = (cf*)base;
    float* partials = (float*)(base + g.t_bytes * (size_t)chunk);
    ColPassParams cp{};
    cp.ksp = a.ksp; cp.sb = a.slice_stride; cp.sa = a.avg_stride; cp.A = a.A; cp.C = a.C; cp.W = (a.flags & MRIACL_PACKED_COLUMNS) ? n_act : a.W;
    cp.act_w = (a.flags & MRIACL_PACKED_COLUMNS) ? pl->act_ident : pl->act_w; cp.act_m = pl->act_m; cp.unit_mask = pl->unit_mask ? 1 : 0; cp.n_act = n_act; cp.n_groups = n_groups;
    cp.tw = pl->twH; cp.T = T; cp.oh = a.oh; cp.ohp = ohp; cp.row0 = row0; cp.flip = flip;
    cp.frame0 = s0 * a.A * a.C; cp.n_frames = ns * a.A * a.C; cp.done = nullptr;
    const long long col_items = (long long)cp.n_frames * n_groups;
    if (col_items > 0) {
      const int grid = (int)std::min<long long>(col_items, (long long)a.sms * 2);
      MRIACL_LAUNCH(colpass640_ws_kernel, grid, CP_WS_T, CP_SMEM_BYTES_WS, a.st, cp);
    }
    RowPass16Params q{};
    q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
    q.sched = pl->sched_p8; q.sched_len = (int)pl->pairs8.size();
    q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
    q.out = a.out + (size_t)s0 * a.oh * a.ow; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C;
    q.scale = (float)(1.0 / std::sqrt((double)a.H * (double)a.Wp));
    q.n_slices = ns; q.n_tiles = g.n_tiles16; q.done = nullptr; q.done_target = 0; q.error_flag = nullptr;
    q.n_buf = n_buf372;
    q.reverse = 1;
    auto kfn = rowpass16_kernel<PP, QQ, NWW, 2>;
    MRIACL_LAUNCH(kfn, std::min(ns * g.n_tiles16, 2 * a.sms), NWW * 32, smem16, a.st, q);
    if (want_norm || a.mean_std) {
      NormParams np{};
      np.in = q.out; np.out = q.out; np.mean_std = a.mean_std ? a.mean_std + 2 * (size_t)s0 : nullptr;
      np.partials = partials; np.n_part = g.n_tiles16; np.n = (long long)a.oh * a.ow; np.eps = a.eps;
      np.normalize = want_norm ? 1 : 0;
      np.n_split = want_norm ? std::max(1, std::min(16, (int)(np.n / 8192))) : 1;
      MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    }
  }
  return 0;
}

// The 640 x 640 plans (prostate-shape) and every other width behind the H = 640 column pass:
// column pass -> 640-wide row pass / pruned generic row pass -> normalise, back to back.
int run_fused640(const FusedArgs& a, const ReconGeom& g) {
  PlanPtr pl = get_fused_plan(a.dev, a.H, a.W, a.pad_left, a.Wp, a.oh, a.ow, a.mask, true);
  const bool wide640 = a.Wp == CP_N;
  if (!pl || (wide640 ? !pl->r640_off : !pl->act_logical)) return fail(MRIACL_ERR_CUDA, "plan upload failed: %s", rt_last_error_string());
  const int n_act = (int)pl->host.act_w.size();
  const int n_groups = (n_act + CP_GW - 1) / CP_GW;
  const int ohp = g.n_tiles * RP_ROWS;
  const int row0 = crop_start(a.H, a.oh), col0 = crop_start(a.Wp, a.ow);
  const bool want_norm = (a.flags & MRIACL_NORM_INSTANCE) != 0;
  const int flip = (a.flags & MRIACL_FLIP_ROWS) ? 1 : 0;
  const int n_ent = (int)pl->r640.ent.size();
  int rg_lines = 8;        // pruned generic row pass: lines per item, halved until the two line buffers + tiles fit
  const std::vector<int> rad = generic_radices(a.Wp);
  while (!wide640 && rg_lines > 1 && rowgen_smem_bytes(a.Wp, rg_lines, a.ow, (int)rad.size()) > SMEM_MAX / 3) rg_lines /= 2;
  // balanced first pass when at most three quarters of the 80 butterfly positions hold samples (undersampled + padded plans)
  const int n_upos = wide640 && pl->r640_upos && pl->r640.upos.size() <= 60 ? (int)pl->r640.upos.size() : 0;
  const int smem = wide640 ? row640_smem_bytes(std::max(1, n_act), n_ent, a.ow, n_upos) : rowgen_smem_bytes(a.Wp, rg_lines, a.ow, (int)rad.size());
  if (smem > SMEM_MAX) return fail(MRIACL_ERR_UNSUPPORTED, "row pass does not fit shared memory (Wp=%d n_act=%d ow=%d)", a.Wp, n_act, a.ow);
  if (!wide640 && (int)rad.size() > RG_MAX_STAGES) return fail(MRIACL_ERR_UNSUPPORTED, "too many FFT stages for N=%d", a.Wp);
  const int n_tiles_row = wide640 ? g.n_tiles8 : (a.oh + rg_lines - 1) / rg_lines;
  const int chunk = (int)std::min<size_t>((size_t)a.B, a.workspace_bytes / g.per_slice);
  const size_t part_bytes = align_up((size_t)a.oh * 3 * sizeof(float), 256);
  for (int s0 = 0; s0 < a.B; s0 += chunk) {
    const int ns = std::min(chunk, a.B - s0);
    char* base = (char*)a.workspace;
    cf* T = (cf*)base;
    float* partials = (float*)(base + g.t_bytes * (size_t)chunk);
    (void)part_bytes;
    ColPassParams cp{};
    cp.ksp = a.ksp; cp.sb = a.slice_stride; cp.sa = a.avg_stride; cp.A = a.A; cp.C = a.C; cp.W = (a.flags & MRIACL_PACKED_COLUMNS) ? n_act : a.W;
    cp.act_w = (a.flags & MRIACL_PACKED_COLUMNS) ? pl->act_ident : pl->act_w; cp.act_m = pl->act_m; cp.unit_mask = pl->unit_mask ? 1 : 0; cp.n_act = n_act; cp.n_groups = n_groups;
    cp.tw = pl->twH; cp.T = T; cp.oh = a.oh; cp.ohp = ohp; cp.row0 = row0; cp.flip = flip;
    cp.frame0 = s0 * a.A * a.C; cp.n_frames = ns * a.A * a.C; cp.done = nullptr;
    const long long col_items = (long long)cp.n_frames * n_groups;
    if (col_items > 0) {
      const int grid = (int)std::min<long long>(col_items, (long long)a.sms * 2);
      MRIACL_LAUNCH(colpass640_ws_kernel, grid, CP_WS_T, CP_SMEM_BYTES_WS, a.st, cp);
    } else if (rt_memset_async(T, 0, g.t_bytes * (size_t)ns, a.st)) {
      return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
    }
    float* out_s0 = a.out + (size_t)s0 * a.oh * a.ow;
    const float scale = (float)(1.0 / std::sqrt((double)a.H * (double)a.Wp));
    const int per_sm = std::max(1, std::min(3, SMEM_MAX / smem));
    if (wide640) {
      Row640Params q{};
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.pos_off = pl->r640_off; q.ent = pl->r640_ent; q.n_ent = n_ent; q.perm = pl->r640_perm; q.tw = pl->twW;
      q.upos = pl->r640_upos; q.n_upos = n_upos;
      q.out = out_s0; q.partials = partials; q.ow = a.ow; q.col0 = col0;
      q.A = a.A; q.C = a.C; q.scale = scale;
      q.n_slices = ns; q.n_tiles = n_tiles_row;
      if (n_upos > 0) MRIACL_LAUNCH(rowpass640_kernel<true>, std::min(ns * n_tiles_row, per_sm * a.sms), R640_T, smem, a.st, q);
      else MRIACL_LAUNCH(rowpass640_kernel<false>, std::min(ns * n_tiles_row, per_sm * a.sms), R640_T, smem, a.st, q);
    } else {
      RowGenParams q{};
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.act_logical = pl->act_logical; q.tw = pl->twW_fwd; q.N = a.Wp; q.L = rg_lines;
      q.out = out_s0; q.partials = partials; q.ow = a.ow; q.col0 = col0;
      q.A = a.A; q.C = a.C; q.scale = scale;
      q.n_slices = ns; q.n_tiles = n_tiles_row;
      q.n_stages = (int)rad.size();
      for (int i = 0; i < q.n_stages; ++i) q.radix[i] = rad[i];
      MRIACL_LAUNCH(rowpass_generic_kernel, std::min(ns * n_tiles_row, per_sm * a.sms), RG_T, smem, a.st, q);
    }
    if (want_norm || a.mean_std) {
      NormParams np{};
      np.in = out_s0; np.out = out_s0; np.mean_std = a.mean_std ? a.mean_std + 2 * (size_t)s0 : nullptr;
      np.partials = partials; np.n_part = n_tiles_row; np.n = (long long)a.oh * a.ow; np.eps = a.eps;
      np.normalize = want_norm ? 1 : 0;
      np.n_split = want_norm ? std::max(1, std::min(16, (int)(np.n / 8192))) : 1;
      MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    }
  }
  return 0;
}

#ifdef MRIACL_EXPERIMENTAL
// ---- side stream + events of the overlapped schedule, one set per (device, caller stream) ----
struct OverlapRes { rt_stream_t side = nullptr; rt_event_t ev_start[2], ev_row[2]; int* error_flag = nullptr; bool ok = false; };
std::map<std::pair<int, void*>, OverlapRes> g_ovl;

OverlapRes* get_overlap_res(int dev, rt_stream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  OverlapRes& r = g_ovl[{dev, (void*)st}];
  if (!r.ok) {
    int bad = rt_stream_create_high_priority(&r.side);
    for (int i = 0; i < 2; ++i) { bad |= rt_event_create(&r.ev_start[i]); bad |= rt_event_create(&r.ev_row[i]); }
    void* ef = nullptr;
    const int zero = 0;
    bad |= rt_malloc(&ef, sizeof(int));
    if (!bad) bad |= rt_upload(ef, &zero, sizeof(int));
    if (bad) return nullptr;
    r.error_flag = (int*)ef;
    r.ok = true;
  }
  return &r;
}

// The fused 640x368 plan.  Two schedules:
//  sequential  column pass (persistent, double-buffered gather) -> row pass (16 warps, one CTA per SM)
//              -> normalise, back to back on the caller's stream, `chunk` slices per group;
//  overlapped  (default on the device) the row pass is ONE persistent launch per group on a side stream
//              with 8-warp CTAs that leave room for column-pass CTAs on every SM; it consumes slices as the
//              concurrently running column pass (caller's stream) publishes them through per-slice counters,
//              so the HBM-bound gather and the issue-bound row transform share the SMs and the intermediate
//              is read back while still in L2.  The side stream is joined to the caller's stream at the end.
int run_fused_experimental(const FusedArgs& a, const ReconGeom& g) {
  PlanPtr pl = get_fused_plan(a.dev, a.H, a.W, a.pad_left, a.Wp, a.oh, a.ow, a.mask, true);
  if (!pl) return fail(MRIACL_ERR_CUDA, "plan upload failed: %s", rt_last_error_string());
  const int n_act = (int)pl->host.act_w.size();
  const int n_groups = (n_act + CP_G - 1) / CP_G;
  const int ohp = g.n_tiles * RP_ROWS;
  const int row0 = crop_start(a.H, a.oh), col0 = crop_start(a.Wp, a.ow);
  const bool want_norm = (a.flags & MRIACL_NORM_INSTANCE) != 0;
  const int flip = (a.flags & MRIACL_FLIP_ROWS) ? 1 : 0;
  const unsigned only = a.flags & (MRIACL_ONLY_COLPASS | MRIACL_ONLY_ROWPASS | MRIACL_ONLY_NORM);
  const bool do_col = !only || (only & MRIACL_ONLY_COLPASS);
  const bool do_row = !only || (only & MRIACL_ONLY_ROWPASS);
  const bool do_norm = !only || (only & MRIACL_ONLY_NORM);
  // schedule: flags win, then MRIACL_SCHEDULE=sequential|fused|overlapped, default sequential (the fastest measured;
  // DESIGN.md section 4.5 has the numbers for the two experimental schedules)
  static const int sched_env = [] {
    const char* e = getenv("MRIACL_SCHEDULE");
    if (e && !strcmp(e, "fused")) return 1;
    if (e && !strcmp(e, "overlapped")) return 2;
    if (e && !strcmp(e, "pair")) return 3;
    if (e && !strcmp(e, "coresident")) return 4;
    if (e && !strcmp(e, "pipelined")) return 5;
    return 0;
  }();
  int sched = sched_env;
  if (a.flags & MRIACL_SEQUENTIAL) sched = 0;
  else if (a.flags & MRIACL_SCHED_FUSED) sched = 1;
  else if (a.flags & MRIACL_SCHED_OVERLAP) sched = 2;
  else if (a.flags & MRIACL_SCHED_PAIR) sched = 3;
  else if (a.flags & MRIACL_SCHED_CORESIDENT) sched = 4;
  else if (a.flags & MRIACL_SCHED_PIPELINED) sched = 5;
  if (n_groups == 0 || (only && sched != 3)) sched = 0;
  if (sched == 3 && (a.A != 1 || !pl->rpp.ok)) sched = 0;      // the pair row pass serves single-average plans of its mask family
  const bool pair_rows = sched == 3;
  const bool coresident = sched == 4;
  // pipelined: the sequential kernels on small chunks, column pass on the caller's stream and row pass + normalise on a
  // side stream, two alternating T buffers that stay in L2 (eviction hints), kernel tails filled by the other stream
  const bool pipelined = sched == 5 && !only;
  const bool fused_mode = sched == 1;
  const bool overlap = sched == 2;

  const FusedPlanHost& hp = overlap ? pl->host_ovl : pl->host;
  const int sched_len = (int)hp.sched.size(), sptw_len = (int)hp.sptw.size();
  // overlapped: single prefetch buffer unless told otherwise, so that two column-pass CTAs fit beside the row pass
  int n_buf = overlap ? env_int("MRIACL_OVL_ROWBUF", 1) : 2;
  if (n_buf < 1 || n_buf > 2) n_buf = 1;
  if (rowpass_smem_bytes(FUSED_P, FUSED_Q, sptw_len, sched_len, n_act, n_buf, a.ow, a.A) > SMEM_MAX) n_buf = 1;
  const int rp_smem = rowpass_smem_bytes(FUSED_P, FUSED_Q, sptw_len, sched_len, n_act, n_buf, a.ow, a.A);
  if (rp_smem > SMEM_MAX)
    return fail(MRIACL_ERR_UNSUPPORTED, "row-pass tile does not fit shared memory (n_act=%d ow=%d A=%d)", n_act, a.ow, a.A);

  const size_t cap_slices = a.workspace_bytes / g.per_slice;
  OverlapRes* ov = nullptr;
  int chunk, n_bufs_ws;
  if (overlap || coresident || pipelined) {
    ov = get_overlap_res(a.dev, a.st);
    if (!ov) return fail(MRIACL_ERR_CUDA, "side stream / event creation failed: %s", rt_last_error_string());
    static const int pipe_chunk = std::max(1, env_int("MRIACL_PIPE_CHUNK", 8));
    if (coresident) { chunk = (int)std::min<size_t>((size_t)a.B, cap_slices); n_bufs_ws = 1; }
    else if (pipelined) {
      if (cap_slices >= 2) { chunk = (int)std::min<size_t>({(size_t)a.B, (size_t)pipe_chunk, cap_slices / 2}); n_bufs_ws = 2; }
      else { chunk = 1; n_bufs_ws = 1; }
    }
    else if (cap_slices >= (size_t)a.B) { chunk = a.B; n_bufs_ws = 1; }
    else if (cap_slices >= 2) { chunk = (int)(cap_slices / 2); n_bufs_ws = 2; }
    else { chunk = 1; n_bufs_ws = 1; }
  } else {
    chunk = (int)std::min<size_t>((size_t)a.B, cap_slices);
    n_bufs_ws = 1;
  }
  const size_t buf_bytes = g.per_slice * (size_t)chunk;
  const size_t part_bytes = align_up((size_t)a.oh * 3 * sizeof(float), 256);
  // row-pass kernel: 0 = 32-row tiles (16 warps), 1 = 16-row tiles 12 warps x 2 CTAs/SM, 2 = 12 warps x 1, 3 = 16 warps x 1
  static const int rp16_cfg = env_int("MRIACL_RP16_CFG", 1);

  int group = 0;
  for (int s0 = 0; s0 < a.B; s0 += chunk, ++group) {
    const int ns = std::min(chunk, a.B - s0);
    const int wb = group % n_bufs_ws;
    char* base = (char*)a.workspace + buf_bytes * wb;
    cf* T = (cf*)base;
    float* partials = (float*)(base + g.t_bytes * (size_t)chunk);
    int* counters = (int*)(base + (g.t_bytes + part_bytes) * (size_t)chunk);

    ColPassParams cp{};
    cp.ksp = a.ksp; cp.sb = a.slice_stride; cp.sa = a.avg_stride; cp.A = a.A; cp.C = a.C; cp.W = (a.flags & MRIACL_PACKED_COLUMNS) ? n_act : a.W;
    cp.act_w = (a.flags & MRIACL_PACKED_COLUMNS) ? pl->act_ident : pl->act_w; cp.act_m = pl->act_m; cp.unit_mask = pl->unit_mask ? 1 : 0; cp.n_act = n_act; cp.n_groups = n_groups;
    cp.tw = pl->twH; cp.T = T; cp.oh = a.oh; cp.ohp = ohp; cp.row0 = row0; cp.flip = flip;
    cp.frame0 = s0 * a.A * a.C; cp.n_frames = ns * a.A * a.C; cp.done = nullptr;
    cp.debug_skip = env_int("MRIACL_CP_DEBUG_SKIP", 0);
    const long long col_items = (long long)cp.n_frames * n_groups;

    RowPassParams rp{};
    rp.T = T; rp.n_act = n_act; rp.oh = a.oh; rp.ohp = ohp;
    rp.sched = overlap ? pl->sched_ovl : pl->sched; rp.sched_len = sched_len;
    rp.n_buf = n_buf; rp.tw = pl->twW; rp.sptw = pl->sptw; rp.sptw_len = sptw_len;
    rp.out = a.out + (size_t)s0 * a.oh * a.ow; rp.partials = partials; rp.ow = a.ow; rp.col0 = col0;
    rp.A = a.A; rp.C = a.C; rp.scale = (float)(1.0 / std::sqrt((double)a.H * (double)a.Wp));
    rp.n_slices = ns; rp.n_tiles = g.n_tiles; rp.done = nullptr; rp.done_target = 0; rp.error_flag = nullptr;
    const int row_items = ns * g.n_tiles;

    NormParams np{};
    np.in = rp.out; np.out = rp.out; np.mean_std = a.mean_std ? a.mean_std + 2 * (size_t)s0 : nullptr;
    np.partials = partials; np.n_part = g.n_tiles; np.n = (long long)a.oh * a.ow; np.eps = a.eps;
    np.normalize = want_norm ? 1 : 0;
    np.n_split = want_norm ? std::max(1, std::min(16, (int)(np.n / 8192))) : 1;
    const bool run_norm = (want_norm || a.mean_std) && do_norm;

    static const int kc_pair = env_int("MRIACL_KC_PAIR", 1);      // co-resident schedule: pair row team when the mask allows
#ifndef MRIACL_EMU
    static const int kc_tma = env_int("MRIACL_KC_TMA", 0);        // co-resident schedule with TMA-fed column teams
    static const int kc_split_sm = env_int("MRIACL_KC_SPLIT_SM", 0);   // > 0: SM-role split with this many column SMs
    CUtensorMap kc_map;
#endif
    if (false) {
#ifndef MRIACL_EMU
    } else if (coresident && kc_tma && !(a.flags & MRIACL_PACKED_COLUMNS) && pl->item_band &&
               coltma_smem_bytes(2, 2, 1) + coltma_table_bytes((int)pl->band_of_item.size(), n_act) + rowpass16_smem_bytes(FUSED_P, FUSED_Q, (int)pl->sptw16.size(), (int)pl->pairs12.size(), pl->rp16_slots, 1, a.ow, a.A) <= SMEM_MAX &&
               kspace_tensor_map(&kc_map, a.ksp, a.slice_stride, a.avg_stride, a.B, a.A, a.C, a.H, a.W)) {
      if (rt_memset_async(counters, 0, 256 * (size_t)ns, a.st)) return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
      CoresTmaParams kp{};
      const int n_bands = (int)pl->band_of_item.size();
      kp.ct.cp = cp; kp.ct.cp.n_groups = n_bands; kp.ct.cp.done = counters;
      kp.ct.item_band = pl->item_band; kp.ct.item_j0 = pl->item_j0;
      kp.ct.n_slots = std::max(2, std::min(CT_MAX_SLOTS, env_int("MRIACL_CT_SLOTS", 2)));
      kp.ct.work_bufs = std::max(1, std::min(2, env_int("MRIACL_CT_WORK", 2)));
      RowPass16Params& q = kp.rp;
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.sched = pl->sched_p12; q.sched_len = (int)pl->pairs12.size();
      q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
      q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C; q.scale = rp.scale;
      q.n_slices = ns; q.n_tiles = g.n_tiles16;
      q.done = counters; q.done_target = a.A * a.C * n_bands; q.error_flag = ov->error_flag;
      q.n_buf = std::max(1, std::min(3, env_int("MRIACL_KC_NBUF", 2)));
      const int kc_ring = std::max(0, env_int("MRIACL_KC_RING", 0));
      if (kc_ring > 0 && kc_ring < ns) {
        kp.ct.cp.ring = kc_ring; kp.ct.cp.rows_done = counters + ns; kp.ct.cp.rows_target = g.n_tiles16;
        q.ring = kc_ring; kp.rows_done = counters + ns;
      }
      int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, q.n_buf, a.ow, a.A);
      const int tab_b = coltma_table_bytes(n_bands, n_act);
      int smem = coltma_smem_bytes(kp.ct.n_slots, 2, kp.ct.work_bufs) + tab_b + smem16;
      if (smem > SMEM_MAX) {   // smallest configuration
        kp.ct.n_slots = 2; kp.ct.work_bufs = 1; q.n_buf = 1;
        smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, 1, a.ow, a.A);
        smem = coltma_smem_bytes(2, 2, 1) + tab_b + smem16;
      }
      if (smem > SMEM_MAX) return fail(MRIACL_ERR_UNSUPPORTED, "co-resident TMA kernel: %d B of shared memory do not fit (slots %d work %d n_buf %d)", smem, kp.ct.n_slots, kp.ct.work_bufs, q.n_buf);
      static const int kc_only = env_int("MRIACL_KC_ONLY", 0);
      if (kc_only == 1) { q.n_slices = 0; if (env_int("MRIACL_KC_NOPUB", 0)) kp.ct.cp.done = nullptr; }
      if (kc_only == 2) { kp.ct.cp.n_frames = 0; q.done = nullptr; }
      np.n_part = g.n_tiles16;
      auto kfn = knee_coresident_tma_kernel<FUSED_P, FUSED_Q>;
      static bool once = false;
      if (!once) { if (rt_allow_smem((const void*)kfn, SMEM_MAX)) return fail(MRIACL_ERR_CUDA, "smem attr"); once = true; }
      const long long items = (long long)cp.n_frames * n_bands;
      MRIACL_LAUNCH(kfn, (int)std::min<long long>(std::max<long long>(items, 1), (long long)a.sms), KT_T, smem, a.st, kc_map, kp);
      if (run_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    } else if (coresident && kc_split_sm > 0 && !(a.flags & MRIACL_PACKED_COLUMNS) && pl->item_band &&
               2 * rowpass16_smem_bytes(FUSED_P, FUSED_Q, (int)pl->sptw16.size(), (int)pl->pairs12.size(), pl->rp16_slots, 1, a.ow, a.A) <= SMEM_MAX &&
               kspace_tensor_map(&kc_map, a.ksp, a.slice_stride, a.avg_stride, a.B, a.A, a.C, a.H, a.W)) {
      // SM-role split (coresident640x368.cuh): kc_split_sm column CTAs, the other SMs run two row teams each
      if (rt_memset_async(counters, 0, 256 * (size_t)ns, a.st)) return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
      SplitParams kp{};
      const int n_bands = (int)pl->band_of_item.size();
      kp.ct.cp = cp; kp.ct.cp.n_groups = n_bands; kp.ct.cp.done = counters;
      kp.ct.item_band = pl->item_band; kp.ct.item_j0 = pl->item_j0;
      kp.ct.n_slots = KX_COL_TEAMS;
      kp.ct.work_bufs = std::max(1, std::min(2, env_int("MRIACL_CT_WORK", 1)));
      RowPass16Params& q = kp.rp;
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.sched = pl->sched_p12; q.sched_len = (int)pl->pairs12.size();
      q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
      q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C; q.scale = rp.scale;
      q.n_slices = ns; q.n_tiles = g.n_tiles16;
      q.done = counters; q.done_target = a.A * a.C * n_bands; q.error_flag = ov->error_flag;
      q.n_buf = std::max(1, std::min(3, env_int("MRIACL_KC_NBUF", 3)));
      int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, q.n_buf, a.ow, a.A);
      while (2 * align_up(smem16, 16) > (size_t)SMEM_MAX && q.n_buf > 1) { --q.n_buf; smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, q.n_buf, a.ow, a.A); }
      kp.row_smem = (int)align_up(smem16, 16);
      const int smem = std::max(coltma_smem_bytes(kp.ct.n_slots, KX_COL_TEAMS, kp.ct.work_bufs) + coltma_table_bytes(n_bands, n_act), 2 * kp.row_smem);
      if (smem > SMEM_MAX) return fail(MRIACL_ERR_UNSUPPORTED, "split kernel: %d B of shared memory do not fit", smem);
      const int grid = a.sms;
      kp.n_col = std::max(1, std::min(grid - 1, kc_split_sm));
      const int kc_ring = std::max(0, env_int("MRIACL_KC_RING", 16));
      if (kc_ring > 0 && kc_ring < ns) {
        kp.ct.cp.ring = kc_ring; kp.ct.cp.rows_done = counters + ns; kp.ct.cp.rows_target = g.n_tiles16;
        q.ring = kc_ring; kp.rows_done = counters + ns;
      }
      static const int kc_only = env_int("MRIACL_KC_ONLY", 0);
      if (kc_only == 1) { q.n_slices = 0; kp.ct.cp.ring = 0; }
      if (kc_only == 2) { kp.ct.cp.n_frames = 0; q.done = nullptr; }
      np.n_part = g.n_tiles16;
      auto kfn = knee_split_kernel<FUSED_P, FUSED_Q>;
      static bool once = false;
      if (!once) { if (rt_allow_smem((const void*)kfn, SMEM_MAX)) return fail(MRIACL_ERR_CUDA, "smem attr"); once = true; }
      MRIACL_LAUNCH(kfn, grid, KX_T, smem, a.st, kc_map, kp);
      if (run_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
#endif
    } else if (coresident && kc_pair && a.A == 1 && pl->rpp.ok) {
      using L = RowPairLayout<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE>;
      if (rt_memset_async(counters, 0, 256 * (size_t)ns, a.st)) return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
      CoresPairParams kp{};
      kp.cp = cp; kp.cp.done = counters;
      RowPairParams& q = kp.rp;
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.slot_of_j = pl->rpp_slot; q.zero_slots = pl->rpp_zero; q.n_zero = (int)pl->rpp.zero_slots.size();
      q.tables = pl->rpp_tab; q.n_slots = L::N_SLOTS;
      q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.C = a.C; q.scale = rp.scale;
      q.n_slices = ns; q.n_tiles = g.n_tiles16;
      q.n_buf = 3;
      q.done = counters; q.done_target = a.C * n_groups; q.error_flag = ov->error_flag;
      const int kc_ring = std::max(0, env_int("MRIACL_KC_RING", 0));    // 0 = no ring (measured faster: the row teams set the pace either way)
      if (kc_ring > 0 && kc_ring < ns) {     // bounded lag between the column teams and the row teams: T stays in L2
        q.ring = kc_ring; q.rows_done = counters + ns;
        kp.cp.ring = kc_ring; kp.cp.rows_done = counters + ns; kp.cp.rows_target = g.n_tiles16;
        kp.cp.l2_hints = 3;
      }
      int smem = L::smem_bytes(q.n_buf, n_act, q.n_zero);
      if (CP_SMEM_BYTES_DB + smem > SMEM_MAX) { q.n_buf = 2; smem = L::smem_bytes(2, n_act, q.n_zero); }
      if (CP_SMEM_BYTES_DB + smem > SMEM_MAX || RPP_ROWS * (a.ow + 1) * 4 > L::ACC_BYTES)
        return fail(MRIACL_ERR_UNSUPPORTED, "co-resident pair tile does not fit shared memory (n_act=%d ow=%d)", n_act, a.ow);
      static const int kc_only = env_int("MRIACL_KC_ONLY", 0);
      if (kc_only == 1) q.n_slices = 0;
      if (kc_only == 2) { kp.cp.n_frames = 0; q.done = nullptr; }
      np.n_part = g.n_tiles16;
#ifdef MRIACL_EMU
      const int grid = 1;
#else
      const int grid = (int)std::min<long long>(col_items, (long long)a.sms);
#endif
      auto kfn = knee_coresident_pair_kernel<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE>;
      MRIACL_LAUNCH(kfn, grid, KP_T, CP_SMEM_BYTES_DB + smem, a.st, kp);
      if (run_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    } else if (coresident) {
      // one persistent launch, one CTA per SM: a column team and a row team in every CTA (coresident640x368.cuh)
      if (rt_memset_async(counters, 0, 256 * (size_t)ns, a.st)) return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
      CoresParams kp{};
      kp.cp = cp; kp.cp.done = counters;
      RowPass16Params& q = kp.rp;
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.sched = pl->sched_p12; q.sched_len = (int)pl->pairs12.size();
      q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
      q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C; q.scale = rp.scale;
      q.n_slices = ns; q.n_tiles = g.n_tiles16;
      q.done = counters; q.done_target = a.A * a.C * n_groups; q.error_flag = ov->error_flag;
      q.n_buf = std::max(1, std::min(2, env_int("MRIACL_KC_NBUF", 2)));   // 1: the CTA stays under 164 KB, which leaves the gather 92 KB of L1
      int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, q.n_buf, a.ow, a.A);
      if (CP_SMEM_BYTES_DB + smem16 > SMEM_MAX) { q.n_buf = 1; smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, 1, a.ow, a.A); }
      if (CP_SMEM_BYTES_DB + smem16 > SMEM_MAX) return fail(MRIACL_ERR_UNSUPPORTED, "co-resident tile does not fit shared memory (n_act=%d ow=%d A=%d)", n_act, a.ow, a.A);
      const bool kc_fuse_norm = env_int("MRIACL_KC_FUSE_NORM", 0) != 0;    // the finishing team normalises (measured slower)
      if (kc_fuse_norm && (want_norm || a.mean_std)) {
        q.tiles_done = counters + ns;
        q.mean_std = a.mean_std ? a.mean_std + 2 * (size_t)s0 : nullptr;
        q.eps = a.eps; q.normalize = want_norm ? 1 : 0;
      }
      np.n_part = g.n_tiles16;
#ifdef MRIACL_EMU   // the emulator runs the CTAs one after another: a single CTA does everything
      const int grid = 1;
#else
      const int grid = (int)std::min<long long>(col_items, (long long)a.sms);
#endif
      static const int kc_only = env_int("MRIACL_KC_ONLY", 0);   // profiling: 1 = column teams only, 2 = row teams only (T from an earlier call)
      if (kc_only == 1) kp.rp.n_slices = 0;
      if (kc_only == 2) { kp.cp.n_frames = 0; kp.rp.done = nullptr; }
      static const int kc_split = env_int("MRIACL_KC_SPLIT", 1);   // 1: warp-group register split (setmaxnreg), 0: uniform 96 registers
      static const int kc_teams = env_int("MRIACL_KC_TEAMS", 1);   // 2: two column teams on 4-column items + the row team
      if (kc_teams == 2) {
        const int groups4 = (n_act + K2_G - 1) / K2_G;
        kp.cp.n_groups = groups4;
        kp.rp.done_target = a.A * a.C * groups4;
        if (kc_only == 2) kp.rp.done = nullptr;
        const long long items4 = (long long)kp.cp.n_frames * groups4;
        const int grid2 = (int)std::min<long long>((items4 + 1) / 2, (long long)a.sms);
        auto kfn = knee_coresident2_kernel<FUSED_P, FUSED_Q>;
        static bool once = false;
        if (!once) { if (rt_allow_smem((const void*)kfn, SMEM_MAX)) return fail(MRIACL_ERR_CUDA, "smem attr"); once = true; }
        MRIACL_LAUNCH(kfn, std::max(1, grid2), K2_T, K2_COL_SMEM + smem16, a.st, kp);
      } else if (kc_teams == 4) {       // one column team on 4-column items + the row team: 116-131 KB of shared memory, 124 KB of L1
        const int groups4 = (n_act + 3) / 4;
        kp.cp.n_groups = groups4;
        kp.rp.done_target = a.A * a.C * groups4;
        if (kc_only == 2) kp.rp.done = nullptr;
        const long long items4 = (long long)kp.cp.n_frames * groups4;
        auto kfn = knee_coresident_kernel<FUSED_P, FUSED_Q, 4>;
        static bool once = false;
        if (!once) { if (rt_allow_smem((const void*)kfn, SMEM_MAX)) return fail(MRIACL_ERR_CUDA, "smem attr"); once = true; }
        MRIACL_LAUNCH(kfn, (int)std::min<long long>(items4, (long long)a.sms), KC_T, CP_SMEM_BYTES_WS + smem16, a.st, kp);
      } else if (kc_split) {
        auto kfn = knee_coresident_split_kernel<FUSED_P, FUSED_Q>;
        MRIACL_LAUNCH(kfn, grid, KS_T, CP_SMEM_BYTES_DB + smem16, a.st, kp);
      } else {
        auto kfn = knee_coresident_kernel<FUSED_P, FUSED_Q>;
        MRIACL_LAUNCH(kfn, grid, KC_T, CP_SMEM_BYTES_DB + smem16, a.st, kp);
      }
      if (run_norm && !kc_fuse_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    } else if (fused_mode) {
      // one persistent launch: column items publish per-slice counters, row items are claimed when ready
      if (rt_memset_async(counters, 0, 256 * (size_t)ns, a.st)) return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
      FusedParams fp{};
      fp.cp = cp; fp.cp.done = counters;
      RowPass16Params& q = fp.rp;
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.sched = pl->sched_p8; q.sched_len = (int)pl->pairs8.size();
      q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
      q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C; q.scale = rp.scale;
      q.n_slices = ns; q.n_tiles = g.n_tiles16; q.done = nullptr; q.done_target = 0; q.error_flag = nullptr;
      q.n_buf = 2;
      int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, 2, a.ow, a.A);
      if (smem16 > SMEM_MAX / 2) { q.n_buf = 1; smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, 1, a.ow, a.A); }
      const int fz_smem = std::max(smem16, CP_SMEM_BYTES_DB);
      if (fz_smem > SMEM_MAX / 2) return fail(MRIACL_ERR_UNSUPPORTED, "fused tile does not fit shared memory (n_act=%d ow=%d A=%d)", n_act, a.ow, a.A);
      fp.state = counters + ns;            // two ints after the per-slice counters (the region is 64 ints per slice)
      fp.n_col_items = (int)col_items; fp.n_row_items = ns * g.n_tiles16;
      static const int col_batch = std::max(1, env_int("MRIACL_FZ_COL_BATCH", 5));
      fp.col_batch = col_batch; fp.done_target = a.A * a.C * n_groups;
      fp.row_ctas_first = env_int("MRIACL_FZ_ROW_FIRST", 0);
      np.n_part = g.n_tiles16;
      const int work = (fp.n_col_items + col_batch - 1) / col_batch + fp.n_row_items;
      auto kfn = fused640_kernel<FUSED_P, FUSED_Q>;
      MRIACL_LAUNCH(kfn, std::min(work, 2 * a.sms), FZ_T, fz_smem, a.st, fp);
      if (run_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    } else if (!overlap) {
      rt_stream_t row_st = pipelined ? ov->side : a.st;
      static const int seq_hints = env_int("MRIACL_SEQ_L2_HINTS", 0);
      if (seq_hints && !pipelined) cp.l2_hints = seq_hints;      // 1 = loads, 2 = stores, 3 = both
      if (pipelined) {
        cp.l2_hints = 3;
        // T buffer wb: its previous reader (row pass of group - 2) must be done
        if (group >= n_bufs_ws && rt_stream_wait_event(a.st, ov->ev_row[wb])) return fail(MRIACL_ERR_CUDA, "stream wait failed");
      }
      const int fuse_norm_env = env_int("MRIACL_FUSE_NORM", 0);   // measured slower than the separate launch (the finishing CTA streams the slice alone)
      const bool fuse_norm = fuse_norm_env && !only && !pair_rows && rp16_cfg != 0 && (want_norm || a.mean_std);
      if (n_groups > 0 && do_col) {
        // tuning knobs: MRIACL_CP_DB=0 single-buffer CTAs, MRIACL_CP_PER_SM=k persistent CTAs per SM (0 = one item per CTA)
        static const int cp_db = env_int("MRIACL_CP_DB", 1), cp_per_sm = env_int("MRIACL_CP_PER_SM", 2);
        static const int cp_ws = env_int("MRIACL_CP_WS", 1);    // warp-specialised gather (default)
        static const int cp_g = env_int("MRIACL_CP_G", 8);      // 4: half-size items (A/B against L1 capacity)
#ifndef MRIACL_EMU
        static const int cp_tma = env_int("MRIACL_CP_TMA", 0);  // n: TMA band gather with n transform teams per CTA
        CUtensorMap kmap;
#endif
        if (false) {
#ifndef MRIACL_EMU
        } else if (cp_tma > 0 && !(a.flags & MRIACL_PACKED_COLUMNS) && pl->item_band &&
            kspace_tensor_map(&kmap, a.ksp, a.slice_stride, a.avg_stride, a.B, a.A, a.C, a.H, a.W)) {
          static const int ct_slots = std::max(cp_tma >= 2 ? 2 : 1, std::min(CT_MAX_SLOTS, env_int("MRIACL_CT_SLOTS", 2)));
          static const int ct_per_sm = std::max(1, std::min(2, env_int("MRIACL_CT_PER_SM", cp_tma == 1 ? 2 : 1)));
          ColTmaParams tp{};
          tp.cp = cp; tp.cp.n_groups = (int)pl->band_of_item.size();
          tp.item_band = pl->item_band; tp.item_j0 = pl->item_j0; tp.n_slots = ct_slots; tp.work_bufs = 2;
          const long long items = (long long)tp.cp.n_frames * tp.cp.n_groups;
          const int nt = cp_tma >= 2 ? 2 : 1;
          const int smem = coltma_smem_bytes(ct_slots, nt) + coltma_table_bytes(tp.cp.n_groups, n_act);
          if (smem > SMEM_MAX / ct_per_sm) return fail(MRIACL_ERR_UNSUPPORTED, "TMA column pass: %d slots x %d teams x %d CTAs/SM do not fit", ct_slots, nt, ct_per_sm);
          static bool once_t = false;
          if (!once_t) {
            rt_allow_smem((const void*)colpass640_tma_kernel<1, 1>, SMEM_MAX); rt_allow_smem((const void*)colpass640_tma_kernel<1, 2>, SMEM_MAX / 2);
            rt_allow_smem((const void*)colpass640_tma_kernel<2, 1>, SMEM_MAX); rt_allow_smem((const void*)colpass640_tma_kernel<2, 2>, SMEM_MAX / 2);
            once_t = true;
          }
          const int grid = (int)std::min<long long>(items, (long long)a.sms * ct_per_sm);
          if (nt == 1 && ct_per_sm == 1) MRIACL_LAUNCH((colpass640_tma_kernel<1, 1>), grid, CP_T, smem, a.st, kmap, tp);
          else if (nt == 1) MRIACL_LAUNCH((colpass640_tma_kernel<1, 2>), grid, CP_T, smem, a.st, kmap, tp);
          else if (ct_per_sm == 1) MRIACL_LAUNCH((colpass640_tma_kernel<2, 1>), grid, 2 * CP_T, smem, a.st, kmap, tp);
          else MRIACL_LAUNCH((colpass640_tma_kernel<2, 2>), grid, 2 * CP_T, smem, a.st, kmap, tp);
#endif
        } else if (cp_ws && cp_db && (cp_g == 4 || cp_g == 2)) {
          ColPassParams c4 = cp;
          c4.n_groups = (n_act + cp_g - 1) / cp_g;
          const long long items4 = (long long)c4.n_frames * c4.n_groups;
          const int per = std::max(2, std::min(4, cp_per_sm));
          const int grid = (int)std::min<long long>(items4, (long long)a.sms * per);
          const int sm4 = 2 * cp_g * CP_PITCH * 8;
          static const int carve4 = env_int("MRIACL_CP_CARVEOUT", -1);
          static bool once4 = false;
          if (!once4) {
            rt_allow_smem((const void*)colpass640_ws_g_kernel<4, 2>, sm4, carve4);
            rt_allow_smem((const void*)colpass640_ws_g_kernel<4, 3>, sm4, carve4);
            rt_allow_smem((const void*)colpass640_ws_g_kernel<4, 4>, sm4, carve4);
            rt_allow_smem((const void*)colpass640_ws_g_kernel<2, 2>, sm4, carve4);
            rt_allow_smem((const void*)colpass640_ws_g_kernel<2, 3>, sm4, carve4);
            rt_allow_smem((const void*)colpass640_ws_g_kernel<2, 4>, sm4, carve4);
            once4 = true;
          }
          if (cp_g == 4 && per == 2) MRIACL_LAUNCH((colpass640_ws_g_kernel<4, 2>), grid, CP_WS_T, sm4, a.st, c4);
          else if (cp_g == 4 && per == 3) MRIACL_LAUNCH((colpass640_ws_g_kernel<4, 3>), grid, CP_WS_T, sm4, a.st, c4);
          else if (cp_g == 4) MRIACL_LAUNCH((colpass640_ws_g_kernel<4, 4>), grid, CP_WS_T, sm4, a.st, c4);
          else if (per == 2) MRIACL_LAUNCH((colpass640_ws_g_kernel<2, 2>), grid, CP_WS_T, sm4, a.st, c4);
          else if (per == 3) MRIACL_LAUNCH((colpass640_ws_g_kernel<2, 3>), grid, CP_WS_T, sm4, a.st, c4);
          else MRIACL_LAUNCH((colpass640_ws_g_kernel<2, 4>), grid, CP_WS_T, sm4, a.st, c4);
        } else if (cp_ws && cp_db) {
          const int grid = (int)std::min<long long>(col_items, (long long)a.sms * std::max(1, cp_per_sm));
          MRIACL_LAUNCH((colpass640_ws_g_kernel<8, 2>), grid, CP_WS_T, CP_SMEM_BYTES_DB, a.st, cp);
        } else if (cp_db) {
          const int grid = (int)std::min<long long>(col_items, (long long)a.sms * std::max(1, cp_per_sm));
          MRIACL_LAUNCH(colpass640_kernel<true>, grid, CP_T, CP_SMEM_BYTES_DB, a.st, cp);
        } else {
          cp.persist = cp_per_sm > 0 ? 1 : 0;
          const int grid = cp.persist ? (int)std::min<long long>(col_items, (long long)a.sms * cp_per_sm) : (int)col_items;
          MRIACL_LAUNCH(colpass640_kernel<false>, grid, CP_T, CP_SMEM_BYTES_SB, a.st, cp);
        }
      }
      if (pipelined && (rt_event_record(ov->ev_start[wb], a.st) || rt_stream_wait_event(ov->side, ov->ev_start[wb])))
        return fail(MRIACL_ERR_CUDA, "pipeline hand-over failed: %s", rt_last_error_string());
      if (fuse_norm && rt_memset_async(counters, 0, sizeof(int) * (size_t)ns, row_st))
        return fail(MRIACL_ERR_CUDA, "memset failed: %s", rt_last_error_string());
      if (do_row && pair_rows) {
        using L = RowPairLayout<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE>;
        RowPairParams q{};
        q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
        q.slot_of_j = pl->rpp_slot; q.zero_slots = pl->rpp_zero; q.n_zero = (int)pl->rpp.zero_slots.size();
        q.tables = pl->rpp_tab; q.n_slots = L::N_SLOTS;
        q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.C = a.C; q.scale = rp.scale;
        q.n_slices = ns; q.n_tiles = g.n_tiles16;
        static const int rpp_buf = std::min(3, std::max(2, env_int("MRIACL_RPP_BUF", 3)));
        q.n_buf = rpp_buf;
        const int smem = L::smem_bytes(q.n_buf, n_act, q.n_zero);
        if (smem > SMEM_MAX / 2 || RPP_ROWS * (a.ow + 1) * 4 > L::ACC_BYTES)
          return fail(MRIACL_ERR_UNSUPPORTED, "pair row pass does not fit shared memory (n_act=%d ow=%d)", n_act, a.ow);
        np.n_part = g.n_tiles16;
        static const int rpp_minb = env_int("MRIACL_RPP_MINB", 2);
        if (rpp_minb == 2) {
          auto kfn = rowpair_kernel<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE, 2>;
          MRIACL_LAUNCH(kfn, std::min(ns * g.n_tiles16, 2 * a.sms), RPP_TT, smem, row_st, q);
        } else {
          auto kfn = rowpair_kernel<FUSED_P, FUSED_Q, RPP_STEP, RPP_NE, 1>;
          MRIACL_LAUNCH(kfn, std::min(ns * g.n_tiles16, a.sms), RPP_TT, smem, row_st, q);
        }
      } else if (do_row && rp16_cfg == 0) {
        auto kfn = rowpass_kernel<FUSED_P, FUSED_Q, RP_NW_SEQ>;
        MRIACL_LAUNCH(kfn, std::min(row_items, a.sms), RP_NW_SEQ * 32, rp_smem, row_st, rp);
      } else if (do_row) {
        const bool w16 = rp16_cfg == 3, w8 = rp16_cfg == 4;
        RowPass16Params q{};
        q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
        q.sched = w16 ? pl->sched_p16 : w8 ? pl->sched_p8 : pl->sched_p12;
        q.sched_len = (int)(w16 ? pl->pairs16.size() : w8 ? pl->pairs8.size() : pl->pairs12.size());
        q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
        q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C; q.scale = rp.scale;
        q.n_slices = ns; q.n_tiles = g.n_tiles16; q.done = nullptr; q.done_target = 0; q.error_flag = nullptr;
        static const int rp_nbuf = std::min(3, std::max(1, env_int("MRIACL_RP_NBUF", 3)));
        q.n_buf = rp_nbuf;
        q.l2_hints = (pipelined || seq_hints) ? 1 : 0;
        static const int rp_reverse = env_int("MRIACL_RP_REVERSE", 1);
        q.reverse = (rp_reverse && do_col) ? 1 : 0;
        q.debug_skip = env_int("MRIACL_RP_DEBUG_SKIP", 0);
        if (fuse_norm) {
          q.tiles_done = counters;
          q.mean_std = a.mean_std ? a.mean_std + 2 * (size_t)s0 : nullptr;
          q.eps = a.eps; q.normalize = want_norm ? 1 : 0;
        }
        int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, q.n_buf, a.ow, a.A);
        const int limit = rp16_cfg == 1 ? SMEM_MAX / 2 : rp16_cfg == 4 ? SMEM_MAX / 3 : SMEM_MAX;
        while (smem16 > limit && q.n_buf > 1) { --q.n_buf; smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, q.n_buf, a.ow, a.A); }
        if (smem16 > limit) return fail(MRIACL_ERR_UNSUPPORTED, "row-pass tile does not fit shared memory (n_act=%d ow=%d A=%d)", n_act, a.ow, a.A);
        const int items16 = ns * g.n_tiles16;
        np.n_part = g.n_tiles16;
        if (rp16_cfg == 1) {
          auto kfn = rowpass16_kernel<FUSED_P, FUSED_Q, 12, 2>;
          MRIACL_LAUNCH(kfn, std::min(items16, 2 * a.sms), 12 * 32, smem16, row_st, q);
        } else if (rp16_cfg == 4) {
          auto kfn = rowpass16_kernel<FUSED_P, FUSED_Q, 8, 3>;
          MRIACL_LAUNCH(kfn, std::min(items16, 3 * a.sms), 8 * 32, smem16, row_st, q);
        } else if (rp16_cfg == 2) {
          auto kfn = rowpass16_kernel<FUSED_P, FUSED_Q, 12, 1>;
          MRIACL_LAUNCH(kfn, std::min(items16, a.sms), 12 * 32, smem16, row_st, q);
        } else {
          auto kfn = rowpass16_kernel<FUSED_P, FUSED_Q, 16, 1>;
          MRIACL_LAUNCH(kfn, std::min(items16, a.sms), 16 * 32, smem16, row_st, q);
        }
      }
      if ((rp16_cfg != 0 || pair_rows) && !do_row) np.n_part = g.n_tiles16;
      if (run_norm && !fuse_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, row_st, np);
      if (pipelined && rt_event_record(ov->ev_row[wb], ov->side)) return fail(MRIACL_ERR_CUDA, "event record failed");
    } else {
      // T buffer wb: its previous reader (row pass of group - n_bufs_ws) must be done
      if (group >= n_bufs_ws && rt_stream_wait_event(a.st, ov->ev_row[wb])) return fail(MRIACL_ERR_CUDA, "stream wait failed");
      if (rt_memset_async(counters, 0, sizeof(int) * (size_t)ns, a.st) || rt_event_record(ov->ev_start[wb], a.st) ||
          rt_stream_wait_event(ov->side, ov->ev_start[wb]))
        return fail(MRIACL_ERR_CUDA, "overlap setup failed: %s", rt_last_error_string());
      cp.done = counters;
      // row pass: the 16-row kernel, one 12-warp CTA per SM, spinning on the per-slice counters; column pass: the
      // warp-specialised kernel, one (or MRIACL_OVL_COL_PER_SM) persistent CTA per SM beside it
      RowPass16Params q{};
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.sched = pl->sched_p12; q.sched_len = (int)pl->pairs12.size();
      q.sptw = pl->sptw16_dev; q.sptw_len = (int)pl->sptw16.size(); q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
      q.out = rp.out; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C; q.scale = rp.scale;
      q.n_slices = ns; q.n_tiles = g.n_tiles16;
      q.done = counters; q.done_target = a.A * a.C * n_groups; q.error_flag = ov->error_flag;
      q.n_buf = 2;
      int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, 2, a.ow, a.A);
      if (smem16 > SMEM_MAX / 2) { q.n_buf = 1; smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, q.sptw_len, q.sched_len, pl->rp16_slots, 1, a.ow, a.A); }
      if (smem16 > SMEM_MAX / 2) return fail(MRIACL_ERR_UNSUPPORTED, "row-pass tile does not fit shared memory (n_act=%d ow=%d A=%d)", n_act, a.ow, a.A);
      np.n_part = g.n_tiles16;
      const int items16 = ns * g.n_tiles16;
      static const int reserve = std::max(0, env_int("MRIACL_OVL_RESERVE_SMS", 4));
      static const int row_per_sm = std::max(1, env_int("MRIACL_OVL_ROW_PER_SM", 1));
      static const int col_per_sm = std::max(1, env_int("MRIACL_OVL_COL_PER_SM", 1));
      const int row_grid = std::min(items16, std::max(1, (a.sms - reserve) * row_per_sm));
      const int col_grid = (int)std::min<long long>(col_items, (long long)a.sms * col_per_sm);
      auto kfn = rowpass16_kernel<FUSED_P, FUSED_Q, 12, 2>;
#ifdef MRIACL_EMU   // the emulator runs launches one after another: producer first
      MRIACL_LAUNCH((colpass640_ws_g_kernel<8, 2>), col_grid, CP_WS_T, CP_SMEM_BYTES_DB, a.st, cp);
      MRIACL_LAUNCH(kfn, row_grid, 12 * 32, smem16, ov->side, q);
#else
      // the row pass goes first so that its CTAs are resident when the column-pass CTAs arrive; a few SMs are
      // left without a row-pass CTA, so the column pass can always make progress whatever the block scheduler does
      MRIACL_LAUNCH(kfn, row_grid, 12 * 32, smem16, ov->side, q);
      MRIACL_LAUNCH((colpass640_ws_g_kernel<8, 2>), col_grid, CP_WS_T, CP_SMEM_BYTES_DB, a.st, cp);
#endif
      if (run_norm) MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, ov->side, np);
      if (rt_event_record(ov->ev_row[wb], ov->side)) return fail(MRIACL_ERR_CUDA, "event record failed");
    }
  }
  if (overlap || pipelined) {   // join: everything on the side stream happens-before whatever the caller enqueues next
    const int used = std::min(group, n_bufs_ws);
    for (int i = 0; i < used; ++i)
      if (rt_stream_wait_event(a.st, ov->ev_row[i])) return fail(MRIACL_ERR_CUDA, "stream join failed");
  }
  return 0;
}

#endif  // MRIACL_EXPERIMENTAL

// The fused 640 x 368 knee plan (product schedule): column pass (persistent, warp-specialised gather) -> 16-row row
// pass -> normalise, back to back on the caller's stream, `chunk` slices per group.  The MRIACL_ONLY_* flags run
// single phases of the same plan (bench.py's per-kernel timings).
int run_fused(const FusedArgs& a, const ReconGeom& g) {
  constexpr unsigned experimental = MRIACL_SCHED_FUSED | MRIACL_SCHED_OVERLAP | MRIACL_SCHED_PAIR | MRIACL_SCHED_CORESIDENT |
                                    MRIACL_SCHED_PIPELINED;
#ifdef MRIACL_EXPERIMENTAL
  if ((a.flags & experimental) || getenv("MRIACL_SCHEDULE") || getenv("MRIACL_RP16_CFG") || getenv("MRIACL_CP_DEBUG_SKIP") ||
      getenv("MRIACL_RP_DEBUG_SKIP") || getenv("MRIACL_CP_PER_SM") || getenv("MRIACL_FUSE_NORM") || getenv("MRIACL_CP_TMA"))
    return run_fused_experimental(a, g);
#else
  if ((a.flags & experimental) && !(a.flags & MRIACL_SEQUENTIAL))
    return fail(MRIACL_ERR_UNSUPPORTED, "schedule not built: this library was compiled without MRIACL_EXPERIMENTAL");
#endif
  PlanPtr pl = get_fused_plan(a.dev, a.H, a.W, a.pad_left, a.Wp, a.oh, a.ow, a.mask, true);
  if (!pl) return fail(MRIACL_ERR_CUDA, "plan upload failed: %s", rt_last_error_string());
  const int n_act = (int)pl->host.act_w.size();
  const int n_groups = (n_act + CP_GW - 1) / CP_GW;
  const int ohp = g.n_tiles * RP_ROWS;
  const int row0 = crop_start(a.H, a.oh), col0 = crop_start(a.Wp, a.ow);
  const bool want_norm = (a.flags & MRIACL_NORM_INSTANCE) != 0;
  const int flip = (a.flags & MRIACL_FLIP_ROWS) ? 1 : 0;
  const unsigned only = a.flags & (MRIACL_ONLY_COLPASS | MRIACL_ONLY_ROWPASS | MRIACL_ONLY_NORM);
  const bool do_col = !only || (only & MRIACL_ONLY_COLPASS);
  const bool do_row = !only || (only & MRIACL_ONLY_ROWPASS);
  const bool do_norm = !only || (only & MRIACL_ONLY_NORM);
  const int chunk = (int)std::min<size_t>((size_t)a.B, a.workspace_bytes / g.per_slice);
  const int sptw_len = (int)pl->sptw16.size(), sched_len = (int)pl->pairs12.size();
  int n_buf = 3;
  int smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, sptw_len, sched_len, pl->rp16_slots, n_buf, a.ow, a.A);
  while (smem16 > SMEM_MAX / 2 && n_buf > 1) { --n_buf; smem16 = rowpass16_smem_bytes(FUSED_P, FUSED_Q, sptw_len, sched_len, pl->rp16_slots, n_buf, a.ow, a.A); }
  if (smem16 > SMEM_MAX / 2)
    return fail(MRIACL_ERR_UNSUPPORTED, "row-pass tile does not fit shared memory (n_act=%d ow=%d A=%d)", n_act, a.ow, a.A);
  for (int s0 = 0; s0 < a.B; s0 += chunk) {
    const int ns = std::min(chunk, a.B - s0);
    char* base = (char*)a.workspace;
    cf* T = (cf*)base;
    float* partials = (float*)(base + g.t_bytes * (size_t)chunk);
    float* out_s0 = a.out + (size_t)s0 * a.oh * a.ow;
    if (n_groups > 0 && do_col) {
      ColPassParams cp{};
      cp.ksp = a.ksp; cp.sb = a.slice_stride; cp.sa = a.avg_stride; cp.A = a.A; cp.C = a.C; cp.W = (a.flags & MRIACL_PACKED_COLUMNS) ? n_act : a.W;
      cp.act_w = (a.flags & MRIACL_PACKED_COLUMNS) ? pl->act_ident : pl->act_w; cp.act_m = pl->act_m; cp.unit_mask = pl->unit_mask ? 1 : 0; cp.n_act = n_act; cp.n_groups = n_groups;
      cp.tw = pl->twH; cp.T = T; cp.oh = a.oh; cp.ohp = ohp; cp.row0 = row0; cp.flip = flip;
      cp.frame0 = s0 * a.A * a.C; cp.n_frames = ns * a.A * a.C; cp.done = nullptr;
      const long long col_items = (long long)cp.n_frames * n_groups;
      MRIACL_LAUNCH(colpass640_ws_kernel, (int)std::min<long long>(col_items, 2LL * a.sms), CP_WS_T, CP_SMEM_BYTES_WS, a.st, cp);
    }
    if (do_row) {
      RowPass16Params q{};
      q.T = T; q.n_act = n_act; q.oh = a.oh; q.ohp = ohp;
      q.sched = pl->sched_p12; q.sched_len = sched_len;
      q.sptw = pl->sptw16_dev; q.sptw_len = sptw_len; q.n_slots = pl->rp16_slots; q.slot_of_j = pl->rp16_slot_dev;
      q.out = out_s0; q.partials = partials; q.ow = a.ow; q.col0 = col0; q.A = a.A; q.C = a.C;
      q.scale = (float)(1.0 / std::sqrt((double)a.H * (double)a.Wp));
      q.n_slices = ns; q.n_tiles = g.n_tiles16; q.n_buf = n_buf;
      q.reverse = (n_groups > 0 && do_col) ? 1 : 0;     // the most recently written slices of T are still in L2
      auto kfn = rowpass16_kernel<FUSED_P, FUSED_Q, 12, 2>;
      MRIACL_LAUNCH(kfn, std::min(ns * g.n_tiles16, 2 * a.sms), 12 * 32, smem16, a.st, q);
    }
    if ((want_norm || a.mean_std) && do_norm) {
      NormParams np{};
      np.in = out_s0; np.out = out_s0; np.mean_std = a.mean_std ? a.mean_std + 2 * (size_t)s0 : nullptr;
      np.partials = partials; np.n_part = g.n_tiles16; np.n = (long long)a.oh * a.ow; np.eps = a.eps;
      np.normalize = want_norm ? 1 : 0;
      np.n_split = want_norm ? std::max(1, std::min(16, (int)(np.n / 8192))) : 1;
      MRIACL_LAUNCH(normalize_instance_kernel, ns * np.n_split, 256, 0, a.st, np);
    }
  }
  return 0;
}

int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

}  // namespace

// =========================================================================================
// Only the C ABI is exported: the library is built with -fvisibility=hidden so that none of its
// inline helpers can be interposed by (or onto) another shared object in the same process.
#pragma GCC visibility push(default)
extern "C" {

int mriacl_abi_version(void) { return MRIACL_ABI_VERSION; }
const char* mriacl_last_error(void) { return g_err.c_str(); }
uint64_t mriacl_launch_count(void) { return launch_counter().load(); }

int mriacl_supported(int H, int W_padded) {
  if (H < 1 || W_padded < 1 || H > MRIACL_MAX_LINE || W_padded > MRIACL_MAX_LINE) return MRIACL_PATH_NONE;
  return fused_shape(H, W_padded) ? MRIACL_PATH_FUSED : MRIACL_PATH_GENERIC;
}

size_t mriacl_recon_rss_workspace_bytes(int slices, int A, int C, int H, int W, int pad_left, int W_padded,
                                        int out_h, int out_w, const float* mask_w_host, unsigned flags) {
  if (slices < 1) slices = 1;
  if (validate_recon(slices, A, C, H, W, pad_left, W_padded, out_h, out_w)) return 0;
  ReconGeom g;
  recon_geom(A, C, H, W, pad_left, W_padded, out_h, out_w, mask_w_host, flags, g);
  return g.per_slice * (size_t)slices;
}

int mriacl_recon_rss_f32(const void* kspace_c64, long long slice_stride, long long avg_stride,
                         const float* mask_w_host, float* out, float* mean_std,
                         int B, int A, int C, int H, int W, int pad_left, int W_padded,
                         int out_h, int out_w, unsigned flags, float eps,
                         void* workspace, size_t workspace_bytes, void* cuda_stream) {
  const int Wp = W_padded, oh = out_h, ow = out_w;
  if (int rc = validate_recon(B, A, C, H, W, pad_left, Wp, oh, ow)) return rc;
  if (B == 0) return MRIACL_OK;
  if (!kspace_c64 || !out || !workspace) return fail(MRIACL_ERR_INVALID, "null kspace/out/workspace pointer");
  const int dev = rt_device();
  if (dev < 0) return fail(MRIACL_ERR_CUDA, "no CUDA device: %s", rt_last_error_string());
  if (ensure_smem_attrs(dev)) return fail(MRIACL_ERR_CUDA, "cudaFuncSetAttribute failed: %s", rt_last_error_string());
  rt_stream_t st = (rt_stream_t)cuda_stream;
  const int sms = device_sms(dev);

  ReconGeom g;
  recon_geom(A, C, H, W, pad_left, Wp, oh, ow, mask_w_host, flags, g);
  if (workspace_bytes < g.per_slice)
    return fail(MRIACL_ERR_WORKSPACE, "workspace %zu B < %zu B needed for one slice", workspace_bytes, g.per_slice);
  int chunk = (int)std::min<size_t>((size_t)B, workspace_bytes / g.per_slice);
  const cf* ksp = (const cf*)kspace_c64;
  const int row0 = crop_start(H, oh), col0 = crop_start(Wp, ow);
  const bool want_norm = (flags & MRIACL_NORM_INSTANCE) != 0;
  const int flip = (flags & MRIACL_FLIP_ROWS) ? 1 : 0;

  if (g.fused) {
    FusedArgs fa{ksp, slice_stride, avg_stride, mask_w_host, out, mean_std, B, A, C, H, W, pad_left, Wp, oh, ow,
                 flags, eps, workspace, workspace_bytes, st, dev, sms};
    const int rc = Wp == FUSED_P * FUSED_Q   ? run_fused(fa, g)
                   : Wp == W372_P * W372_Q ? run_fused_pq<W372_P, W372_Q, W372_NW>(fa, g)
                   : Wp == W400_P * W400_Q ? run_fused_pq<W400_P, W400_Q, W400_NW>(fa, g)
                                           : run_fused640(fa, g);
    if (rc) return rc;
  } else {
    if (flags & MRIACL_PACKED_COLUMNS)
      return fail(MRIACL_ERR_UNSUPPORTED, "packed k-space columns need a plan with the 640-row column pass (H=%d Wp=%d has none)", H, Wp);
    const float* mask_dev = nullptr;
    DevPtr mask_keep;
    if (get_device_mask(dev, mask_w_host, W, &mask_dev, mask_keep)) return fail(MRIACL_ERR_CUDA, "mask upload failed: %s", rt_last_error_string());
    for (int s0 = 0; s0 < B; s0 += chunk) {
      const int ns = std::min(chunk, B - s0);
      cf* img = (cf*)workspace;
      if (int rc = generic_fft2c(ksp + (long long)s0 * slice_stride, slice_stride, avg_stride, A, C, img, ns * A * C,
                                 H, W, pad_left, Wp, mask_dev, 1, dev, st)) return rc;
      RssCropParams rc{};
      rc.img = img; rc.out = out + (size_t)s0 * oh * ow; rc.n_slices = ns; rc.A = A; rc.C = C; rc.H = H; rc.Wp = Wp;
      rc.oh = oh; rc.ow = ow; rc.row0 = row0; rc.col0 = col0; rc.flip = flip;
      MRIACL_LAUNCH(rss_crop_kernel, grid_for((long long)ns * oh * ow, 256), 256, 0, st, rc);
      if (want_norm || mean_std) {
        NormParams np{};
        np.in = rc.out; np.out = rc.out; np.mean_std = mean_std ? mean_std + 2 * (size_t)s0 : nullptr;
        np.partials = nullptr; np.n_part = 0; np.n_split = 1; np.n = (long long)oh * ow; np.eps = eps; np.normalize = want_norm ? 1 : 0;
        MRIACL_LAUNCH(normalize_instance_kernel, ns, 256, 0, st, np);
      }
    }
  }
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
}

size_t mriacl_ifft2c_abs_workspace_bytes(int B, int H, int W) {
  if (B < 1) B = 1;
  if (H < 1 || W < 1) return 0;
  if (fused_shape(H, W) || (pruned_shape(H, W) && pruned_fits(W, W))) return mriacl_recon_rss_workspace_bytes(B, 1, 1, H, W, 0, W, H, W, nullptr, 0);
  return align_up((size_t)B * H * W * sizeof(cf), 256);
}

int mriacl_ifft2c_abs_f32(const void* kspace_c64, float* out, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (B < 0 || H < 1 || W < 1) return fail(MRIACL_ERR_INVALID, "bad dims B=%d H=%d W=%d", B, H, W);
  if (B == 0) return MRIACL_OK;
  // RSS over a single coil is the magnitude: reuse the fused stage without crop or normalisation
  return mriacl_recon_rss_f32(kspace_c64, (long long)H * W, 0, nullptr, out, nullptr, B, 1, 1, H, W, 0, W, H, W,
                              0u, 0.f, workspace, workspace_bytes, cuda_stream);
}

int mriacl_pack_columns_host(const void* kspace_c64_host, void* packed_c64_host, long long n_rows, int W,
                             const float* mask_w_host, int n_threads) {
  if (n_rows < 0 || W < 1) return fail(MRIACL_ERR_INVALID, "bad dims n_rows=%lld W=%d", n_rows, W);
  std::vector<int> idx;
  for (int w = 0; w < W; ++w) if (!mask_w_host || mask_w_host[w] != 0.0f) idx.push_back(w);
  if (n_rows == 0 || idx.empty()) return (int)idx.size();
  if (!kspace_c64_host || !packed_c64_host) return fail(MRIACL_ERR_INVALID, "null pointer");
  pack_columns(kspace_c64_host, packed_c64_host, n_rows, W, idx, n_threads);
  return (int)idx.size();
}

int mriacl_fft2c_c64(const void* in_c64, void* out_c64, int B, int H, int W, int inverse, void* cuda_stream) {
  if (B < 0 || H < 1 || W < 1) return fail(MRIACL_ERR_INVALID, "bad dims B=%d H=%d W=%d", B, H, W);
  if (H > MRIACL_MAX_LINE || W > MRIACL_MAX_LINE) return fail(MRIACL_ERR_UNSUPPORTED, "line length above %d", MRIACL_MAX_LINE);
  if (B == 0) return MRIACL_OK;
  if (!in_c64 || !out_c64) return fail(MRIACL_ERR_INVALID, "null pointer");
  const int dev = rt_device();
  if (dev < 0) return fail(MRIACL_ERR_CUDA, "no CUDA device: %s", rt_last_error_string());
  if (ensure_smem_attrs(dev)) return fail(MRIACL_ERR_CUDA, "cudaFuncSetAttribute failed: %s", rt_last_error_string());
  if (int rc = generic_fft2c((const cf*)in_c64, (long long)H * W, 0, 1, 1, (cf*)out_c64, B, H, W, 0, W, nullptr,
                             inverse ? 1 : 0, dev, (rt_stream_t)cuda_stream)) return rc;
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
}

int mriacl_complex_abs_f32(const void* in_c64, float* out, size_t n, int squared, void* cuda_stream) {
  if (n == 0) return MRIACL_OK;
  if (!in_c64 || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  MRIACL_LAUNCH(complex_abs_kernel, grid_for((long long)n, 256), 256, 0, (rt_stream_t)cuda_stream,
                (const cf*)in_c64, out, (long long)n, squared);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
}

int mriacl_rss_f32(const void* in, float* out, size_t outer, int C, size_t inner, int is_complex, void* cuda_stream) {
  if (C < 1) return fail(MRIACL_ERR_INVALID, "bad coil count %d", C);
  if (outer * inner == 0) return MRIACL_OK;
  if (!in || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  MRIACL_LAUNCH(rss_kernel, grid_for((long long)(outer * inner), 256), 256, 0, (rt_stream_t)cuda_stream,
                (const float*)in, out, (long long)outer, C, (long long)inner, is_complex);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
}

int mriacl_center_crop_or_pad(const void* in, void* out, int B, int H, int W, int out_h, int out_w,
                              int elem_bytes, void* cuda_stream) {
  if (B < 0 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (elem_bytes != 4 && elem_bytes != 8) return fail(MRIACL_ERR_INVALID, "elem_bytes must be 4 or 8");
  if (B == 0) return MRIACL_OK;
  if (!in || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  MRIACL_LAUNCH(crop_or_pad_kernel, grid_for((long long)B * out_h * out_w, 256), 256, 0, (rt_stream_t)cuda_stream,
                (const float*)in, (float*)out, B, H, W, out_h, out_w, elem_bytes / 4);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
}

int mriacl_normalize_instance_f32(const float* in, float* out, float* mean_std, int B, size_t n, float eps,
                                  void* cuda_stream) {
  if (B < 0 || n < 1) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (B == 0) return MRIACL_OK;
  if (!in || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  NormParams np{};
  np.in = in; np.out = out; np.mean_std = mean_std; np.partials = nullptr; np.n_part = 0; np.n_split = 1;
  np.n = (long long)n; np.eps = eps; np.normalize = 1;
  MRIACL_LAUNCH(normalize_instance_kernel, B, 256, 0, (rt_stream_t)cuda_stream, np);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
}

// ---- the per-slice steps after the reconstruction (REF/src/preprocess/mri_preprocess.py:182-191,216-233) ----
#ifdef MRIACL_EMU
#define MRIACL_POST_GUARD return fail(MRIACL_ERR_UNSUPPORTED, "not part of the emulation build")
#else
#define MRIACL_POST_GUARD do {} while (0)
#endif

#ifndef MRIACL_EMU
static int post_allow_smem() {          // once per device: the percentile kernel's 128 KB of lane-private histograms
  static std::mutex mu;
  static std::map<int, bool> done;
  const int dev = rt_device();
  std::lock_guard<std::mutex> lk(mu);
  if (done[dev]) return 0;
  if (rt_allow_smem((const void*)percentile_clip_kernel<1>, POST_HIST_BYTES) || rt_allow_smem((const void*)percentile_clip_kernel<2>, POST_HIST_BYTES) ||
      rt_allow_smem((const void*)percentile_clip_kernel<4>, POST_HIST_BYTES)) return 1;
  done[dev] = true;
  return 0;
}
// CTAs per image of the one-image kernels (thread-block cluster size): as many as keep B clusters within one wave
static int post_cluster_size(int B) {
  const int sms = device_sms(rt_device());
  return 4 * B <= sms ? 4 : 2 * B <= sms ? 2 : 1;
}
static void launch_percentile(const PercentileParams& p, int B, rt_stream_t st) {
  const int cl = post_cluster_size(B);
  if (cl == 4) MRIACL_LAUNCH(percentile_clip_kernel<4>, 4 * B, POST_T, POST_HIST_BYTES, st, p);
  else if (cl == 2) MRIACL_LAUNCH(percentile_clip_kernel<2>, 2 * B, POST_T, POST_HIST_BYTES, st, p);
  else MRIACL_LAUNCH(percentile_clip_kernel<1>, B, POST_T, POST_HIST_BYTES, st, p);
}
static void launch_zscore(const ZscoreParams& p, int B, rt_stream_t st) {
  const int cl = post_cluster_size(B);
  if (cl == 4) MRIACL_LAUNCH(zscore_preview_kernel<4>, 4 * B, POST_T, 0, st, p);
  else if (cl == 2) MRIACL_LAUNCH(zscore_preview_kernel<2>, 2 * B, POST_T, 0, st, p);
  else MRIACL_LAUNCH(zscore_preview_kernel<1>, B, POST_T, 0, st, p);
}
#endif

int mriacl_percentile_clip_f32(const float* in, float* out, float* lo_hi, int B, size_t n, float pmin, float pmax,
                               void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (B < 0 || n < 1) return fail(MRIACL_ERR_INVALID, "bad dims B=%d n=%zu", B, n);
  if (!(pmin >= 0.f && pmin < pmax && pmax <= 100.f)) return fail(MRIACL_ERR_INVALID, "percentiles must satisfy 0 <= pmin < pmax <= 100");
  if (B == 0) return MRIACL_OK;
  if (!in || (!out && !lo_hi)) return fail(MRIACL_ERR_INVALID, "null pointer");
  PercentileParams p{in, out, lo_hi, (long long)n, pmin, pmax};
  if (post_allow_smem()) return fail(MRIACL_ERR_CUDA, "cudaFuncSetAttribute failed: %s", rt_last_error_string());
  launch_percentile(p, B, (rt_stream_t)cuda_stream);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

int mriacl_resize_bilinear_f32(const float* in, float* out, int B, int H, int W, int out_h, int out_w, void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (B < 0 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (B == 0) return MRIACL_OK;
  if (!in || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  ResizeParams p{in, nullptr, out, nullptr, nullptr, B, H, W, out_h, out_w};
  MRIACL_LAUNCH(resize_bilinear_kernel, grid_for((long long)B * out_h * out_w, 256), 256, 0, (rt_stream_t)cuda_stream, p);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

int mriacl_resize_mask_u8(const uint8_t* in, uint8_t* out, int B, int H, int W, int out_h, int out_w, void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (B < 0 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (B == 0) return MRIACL_OK;
  if (!in || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  ResizeParams p{nullptr, in, nullptr, out, nullptr, B, H, W, out_h, out_w};
  MRIACL_LAUNCH(resize_bilinear_kernel, grid_for((long long)B * out_h * out_w, 256), 256, 0, (rt_stream_t)cuda_stream, p);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

int mriacl_zscore_preview_f32(const float* in, const uint8_t* mask, float* out_z, float* out_01, float* stats, int B,
                              size_t n, void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (B < 0 || n < 1) return fail(MRIACL_ERR_INVALID, "bad dims B=%d n=%zu", B, n);
  if (B == 0) return MRIACL_OK;
  if (!in || (!out_z && !out_01 && !stats)) return fail(MRIACL_ERR_INVALID, "null pointer");
  ZscoreParams p{in, mask, out_z, out_01, stats, (long long)n};
  launch_zscore(p, B, (rt_stream_t)cuda_stream);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

int mriacl_clip_resize_zscore_f32(const float* img, const uint8_t* body_mask, float* out_z, float* out_01,
                                  uint8_t* out_mask, float* clip_lo_hi, float* stats, int B, int H, int W, int out_h,
                                  int out_w, float pmin, float pmax, void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (B < 0 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (!(pmin >= 0.f && pmin < pmax && pmax <= 100.f)) return fail(MRIACL_ERR_INVALID, "percentiles must satisfy 0 <= pmin < pmax <= 100");
  if (B == 0) return MRIACL_OK;
  if (!img || !out_z || !clip_lo_hi) return fail(MRIACL_ERR_INVALID, "null pointer (img, out_z and clip_lo_hi are required)");
  if (body_mask && !out_mask) return fail(MRIACL_ERR_INVALID, "out_mask is required when a body mask is given");
  rt_stream_t st = (rt_stream_t)cuda_stream;
  // 1. exact percentiles of every full-resolution image (the clipped image itself is never materialised)
  PercentileParams pp{img, nullptr, clip_lo_hi, (long long)H * W, pmin, pmax};
  if (post_allow_smem()) return fail(MRIACL_ERR_CUDA, "cudaFuncSetAttribute failed: %s", rt_last_error_string());
  launch_percentile(pp, B, st);
  // 2. clip every tap, interpolate to (out_h, out_w); the mask goes through the same interpolation and a 0.5 threshold
  // (one launch: a thread computes its source coordinates and weights once and applies them to both planes)
  ResizeParams ri{img, nullptr, out_z, nullptr, clip_lo_hi, B, H, W, out_h, out_w, body_mask, body_mask ? out_mask : nullptr};
  MRIACL_LAUNCH(resize_bilinear_kernel, grid_for((long long)B * out_h * out_w, 256), 256, 0, st, ri);
  // 3. statistics inside the resized mask, z-score in place, preview
  ZscoreParams zp{out_z, body_mask ? out_mask : nullptr, out_z, out_01, stats, (long long)out_h * out_w};
  launch_zscore(zp, B, st);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

int mriacl_stack25d_f32(const float* in, float* out, int S, size_t n, int k, int repeat, const float* mean_dev,
                        const float* std_dev, void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (S < 0 || n < 1 || k < 1 || (k % 2 == 0 && !repeat)) return fail(MRIACL_ERR_INVALID, "bad dims S=%d n=%zu k=%d (k must be odd)", S, n, k);
  if ((mean_dev == nullptr) != (std_dev == nullptr)) return fail(MRIACL_ERR_INVALID, "mean and std go together");
  if (S == 0) return MRIACL_OK;
  if (!in || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  StackParams p{in, out, mean_dev, std_dev, (long long)n, S, k, repeat ? 1 : 0};
  MRIACL_LAUNCH(stack25d_kernel, grid_for((long long)S * k * (long long)n, 256), 256, 0, (rt_stream_t)cuda_stream, p);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

// ---- GRAPPA weight application and the SENSE-style combine ----
int mriacl_grappa_apply_c64(void* kspace_inout, long long slice_stride, long long sx, long long sy, long long sc,
                            int n_slices, int X, int Y, int C, int kx, int ky,
                            const int* hole_xy, int n_items, const int* item_geom, const int* item_first,
                            const int* item_count, const int* geom_src_start, const int* src_off, int max_sources,
                            const long long* geom_w_start, const void* weights_c64, long long weights_per_slice,
                            void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (n_slices < 0 || X < 1 || Y < 1 || C < 1 || n_items < 0) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (kx < 1 || ky < 1 || kx > 7 || ky > 7 || !(kx & 1) || !(ky & 1)) return fail(MRIACL_ERR_UNSUPPORTED, "kernel size must be odd and at most 7 x 7, got %d x %d", kx, ky);
  if (max_sources < 0 || max_sources > kx * ky) return fail(MRIACL_ERR_INVALID, "max_sources %d outside 0..%d", max_sources, kx * ky);
  if (n_slices == 0 || n_items == 0) return MRIACL_OK;
  if (!kspace_inout || !hole_xy || !item_geom || !item_first || !item_count || !geom_src_start || !src_off || !geom_w_start || !weights_c64)
    return fail(MRIACL_ERR_INVALID, "null pointer");
  const size_t smem = (size_t)max_sources * C * GR_OUT * sizeof(cf);
  if (smem > (size_t)SMEM_MAX) return fail(MRIACL_ERR_UNSUPPORTED, "weight tile of %zu B does not fit shared memory (%d sources x %d coils)", smem, max_sources, C);
  static std::mutex mu;
  static std::map<int, size_t> allowed;             // per device: largest opt-in so far
  const int dev = rt_device();
  {
    std::lock_guard<std::mutex> lk(mu);
    if (allowed[dev] < smem) {
      if (rt_allow_smem((const void*)grappa_apply_kernel, (int)smem)) return fail(MRIACL_ERR_CUDA, "cudaFuncSetAttribute failed: %s", rt_last_error_string());
      allowed[dev] = smem;
    }
  }
  GrappaParams p{};
  p.ksp = (const cf*)kspace_inout; p.out = (cf*)kspace_inout; p.ss = slice_stride; p.sx = sx; p.sy = sy; p.sc = sc; p.nc = C;
  p.hole_xy = hole_xy; p.Y = Y; p.item_geom = item_geom; p.item_first = item_first; p.item_count = item_count;
  p.geom_src_start = geom_src_start; p.src_off = src_off; p.geom_w_start = geom_w_start;
  p.weights = (const cf*)weights_c64; p.w_per_slice = weights_per_slice; p.kx2 = kx / 2; p.ky2 = ky / 2;
  const int groups = (C + GR_OUT - 1) / GR_OUT;
  if (groups > 1) return fail(MRIACL_ERR_UNSUPPORTED, "more than %d coils: in-place application needs all outputs of a hole in one pass", GR_OUT);
  for (int s0 = 0; s0 < n_slices; s0 += 65535) {     // grid.y limit
    const int ns = std::min(65535, n_slices - s0);
    GrappaParams q = p;
    q.ksp = p.ksp + (long long)s0 * slice_stride; q.out = p.out + (long long)s0 * slice_stride;
    q.weights = p.weights + (long long)s0 * weights_per_slice;
    dim3 grid((unsigned)n_items, (unsigned)ns, 1);
    grappa_apply_kernel<<<grid, GR_T, smem, (rt_stream_t)cuda_stream>>>(q);
    launch_counter()++;
  }
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

int mriacl_sense_combine(const void* img_c64, const void* sens_c64, void* out, int B, int C, size_t n, int shared_sens,
                         int magnitude, void* cuda_stream) {
  MRIACL_POST_GUARD;
#ifndef MRIACL_EMU
  if (B < 0 || C < 1 || n < 1) return fail(MRIACL_ERR_INVALID, "bad dims");
  if (B == 0) return MRIACL_OK;
  if (!img_c64 || !sens_c64 || !out) return fail(MRIACL_ERR_INVALID, "null pointer");
  SenseParams p{(const cf*)img_c64, (const cf*)sens_c64, out, (long long)n, B, C, shared_sens ? 1 : 0, magnitude ? 1 : 0};
  MRIACL_LAUNCH(sense_combine_kernel, grid_for((long long)B * (long long)n, 256), 256, 0, (rt_stream_t)cuda_stream, p);
  if (rt_check()) return fail(MRIACL_ERR_CUDA, "kernel launch failed: %s", rt_last_error_string());
  return MRIACL_OK;
#endif
}

}  // extern "C"
#pragma GCC visibility pop
