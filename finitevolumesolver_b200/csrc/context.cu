// context.cu -- host side of libfvs_b200.so: the C ABI of include/fvs_b200.h.
//
// A context plays the role of the reference's fub::amrex::IntegratorContext
// (src/AMReX/IntegratorContext.cpp) for the levels it is given: it owns the
// scratch state of every box in HBM (patch-local SoA, FAB layout), derives
// the iteration domains of the adapters (boxes.hpp), and sequences sweeps the
// way DimensionalSplitLevelIntegrator / SubcycleFineFirstSolver /
// RunSimulation do.  There is no CPU fallback: every numerical entry point
// launches CUDA kernels and fails if no device is present.
#include "../../include/fvs_b200.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <sys/stat.h>
#include <cerrno>
#include <algorithm>
#include <limits>
#include <map>
#include <set>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "boxes.hpp"
#include "launch.h"

namespace {

thread_local std::string g_error;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define FVS_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      throw Error(std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
    }                                                                                         \
  } while (0)

// errors of kernels launched earlier surface at the next check: say which one
inline void check_last(const char* where) {
  const cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) throw Error(std::string(where) + ": " + cudaGetErrorString(e_));
}

template <typename F> int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_error = e.what();
    return FVS_ERROR;
  }
}

// ---- NCCL, loaded lazily (the library must load on a box without it) -------
struct Nccl {
  typedef struct ncclComm* comm_t;
  struct unique_id {
    char internal[128];
  };
  void* handle{nullptr};
  int (*GetUniqueId)(unique_id*){nullptr};
  int (*CommInitRank)(comm_t*, int, unique_id, int){nullptr};
  int (*CommDestroy)(comm_t){nullptr};
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t){nullptr};
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t){nullptr};
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t){nullptr};
  int (*GroupStart)(){nullptr};
  int (*GroupEnd)(){nullptr};
  const char* (*GetErrorString)(int){nullptr};
  static constexpr int kDouble = 8; // ncclFloat64
  static constexpr int kMax = 2;    // ncclMax
  static constexpr int kMin = 3;    // ncclMin

  static Nccl& get() {
    static Nccl n;
    if (!n.handle) {
      n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!n.handle) n.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!n.handle) throw Error(std::string("cannot load NCCL: ") + dlerror());
      auto sym = [&](const char* name) {
        void* p = dlsym(n.handle, name);
        if (!p) throw Error(std::string("NCCL symbol missing: ") + name);
        return p;
      };
      n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
      n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
      n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
      n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
      n.Send = (decltype(n.Send))sym("ncclSend");
      n.Recv = (decltype(n.Recv))sym("ncclRecv");
      n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
      n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
      n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    }
    return n;
  }
  void check(int rc, const char* what) {
    if (rc != 0) throw Error(std::string(what) + ": " + GetErrorString(rc));
  }
};

using fvs::Box;
using fvs::BoxDesc;

constexpr int kNumBuffers = 3;
constexpr unsigned kRepairCap = 1u << 18; // repair entries per sweep before "every cell" (6 MB)
constexpr int kDtSlots = 64;

// Embedded-boundary data of one box (fvs_level_set_eb); CutCellData of
// include/fub/CutCellData.hpp:48-75 plus the cut-cell arrays of
// cutcell::IntegratorContext (src/AMReX/cutcell/IntegratorContext.cpp:470-533).
struct EbPatch {
  int fab_type{FVS_FAB_REGULAR};
  double* alpha{nullptr};
  double* beta[3]{nullptr, nullptr, nullptr};
  double* bcent{nullptr};
  double* bnorm{nullptr};
  double* us[3]{nullptr, nullptr, nullptr};
  double* sl[3]{nullptr, nullptr, nullptr};
  double* sr[3]{nullptr, nullptr, nullptr};
  double* ds[3]{nullptr, nullptr, nullptr};
  double* reference{nullptr};
  double* boundary_flux{nullptr};
  double* stab[3]{nullptr, nullptr, nullptr};
  double* shl[3]{nullptr, nullptr, nullptr};
  double* shr[3]{nullptr, nullptr, nullptr};
  long long* cut_cells{nullptr};
  int n_cut{0};
  unsigned char* complete_mode{nullptr};
  unsigned char* flags{nullptr}; // IsCutCell map
};

struct Patch {
  int global_index{0};
  Box valid, sbox;
  int n[3]{1, 1, 1};
  long long cells{0};
  double* buf[kNumBuffers]{nullptr, nullptr, nullptr};
  double* flux[3]{nullptr, nullptr, nullptr};
  Box fbox[3];
  unsigned char* dtmask[2][3]{{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  // cut-cell adapter: the stable dt runs over every face of the flux box
  // (include/fub/AMReX/cutcell/FluxMethod.hpp:262-289), not over the tiled stencil set
  unsigned char* dtmask_eb[3]{nullptr, nullptr, nullptr};
  EbPatch eb;
};

struct BoundaryCondition {
  int dir, side, kind;
};

// cudaMemcpy from pageable host memory returns once the data sit in the driver's
// staging buffer; the DMA to the device is ordered on the legacy default stream, which
// the contexts' non-blocking streams do not wait for -- a kernel launched right after
// could read the destination before the data arrive.  Every host-to-device upload of a
// descriptor / mask array therefore waits for the legacy stream.
inline cudaError_t upload_sync(void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(cudaStreamLegacy);
}

template <typename T> struct DeviceArray {
  T* ptr{nullptr};
  size_t count{0};
  DeviceArray() = default;
  DeviceArray(const DeviceArray&) = delete;
  DeviceArray& operator=(const DeviceArray&) = delete;
  ~DeviceArray() {
    if (ptr) cudaFree(ptr);
  }
  void upload(const std::vector<T>& host) {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = host.size();
    if (count == 0) return;
    FVS_CUDA(cudaMalloc(&ptr, count * sizeof(T)));
    FVS_CUDA(upload_sync(ptr, host.data(), count * sizeof(T)));
  }
};

struct PeerExchange {
  int peer{0};
  std::vector<fvs::HaloCopy> sends, recvs; // canonical (global plan) order
  long long send_elems{0}, recv_elems{0};
  double* sendbuf{nullptr};
  double* recvbuf{nullptr};
};

// One coarse-fine operation between two levels in executable form: descriptor
// arrays and staging buffers live on the device until the hierarchy changes.
struct CfXfer {
  int src, dst;
  size_t count;
  double* ptr; // the local end (send buffer on src, receive buffer on dst)
};
struct CfPlan {
  std::vector<CfXfer> xfers;
  std::vector<void*> owned; // cudaMalloc'ed: staging buffers, descriptor arrays, masks
  const void* packs{nullptr};
  int n_packs{0};
  long long packs_max{1};
  const void* unpacks{nullptr};
  int n_unpacks{0};
  long long unpacks_max{1};
  const void* main{nullptr};
  int n_main{0};
  long long main_max{1};
  CfPlan() = default;
  CfPlan(const CfPlan&) = delete;
  ~CfPlan() {
    for (void* q : owned) cudaFree(q);
  }
};

struct Level {
  bool defined{false};
  int rank{1};
  Box domain;
  double dx[3]{1, 1, 1};
  bool periodic[3]{false, false, false};
  int scratch_gcw{0}, flux_gcw{0};
  std::vector<Box> boxes; // all boxes of the level
  std::vector<int> owner;
  std::vector<int> local_of_global; // -1 if not owned
  std::vector<Patch> patches;       // owned boxes only
  std::vector<BoundaryCondition> bcs;
  int cur{0};   // buffer holding the current state
  int keep{-1}; // buffer that must not be overwritten (backup), -1 none
  bool have_buf[kNumBuffers]{false, false, false};
  int max_n[3]{1, 1, 1};
  std::vector<fvs::HaloCopy> local_copies;
  std::vector<PeerExchange> peers;
  // descriptor caches, keyed by (kind, dir, src buffer, dst buffer, width)
  std::map<std::tuple<int, int, int, int, int>, std::unique_ptr<DeviceArray<BoxDesc>>> box_cache;
  // register planes (fused sweeps on AMR levels), keyed by (dir, src buffer): descriptors + count
  std::map<std::tuple<int, int>, std::pair<std::unique_ptr<DeviceArray<BoxDesc>>, int>> plane_cache;
  std::map<int, std::unique_ptr<DeviceArray<fvs::CopyDesc>>> copy_cache;
  std::map<std::tuple<int, int>, std::unique_ptr<DeviceArray<fvs::BcDesc>>> bc_cache;
  std::map<std::tuple<int, int, int>, std::unique_ptr<DeviceArray<fvs::PackDesc>>> pack_cache;
  long long max_copy_cells{0};
  // halo exchange over peer memory (p2p_*, below): the peers' state arrays of buffer b, opened
  // through CUDA IPC (indexed by global box, nullptr where not needed), and the region copies
  // from this rank's boxes straight into them
  std::vector<double*> remote_buf[kNumBuffers];
  bool have_remote[kNumBuffers]{false, false, false};
  std::map<int, std::unique_ptr<DeviceArray<fvs::CopyDesc>>> p2p_cache;
  int p2p_count{0};
  long long p2p_max_cells{0};
  bool halo_crosses_ranks{false}; // some halo copy of the level has its ends on different ranks
  // coarse-fine: the flux register this level holds towards its coarser level
  struct CfRegister {
    int patch, dir, side;
    Box plane; // coarse faces (index of the cell to the right of the face)
    double* reg{nullptr};
    unsigned char* open{nullptr};
  };
  std::vector<CfRegister> cf_regs;
  int cf_coarse{-1};
  // coarse-fine operations towards the coarser level, built once per hierarchy
  // configuration: key (operation, dir, buffer of this level, buffer of the coarser level)
  std::map<std::tuple<int, int, int, int>, std::shared_ptr<CfPlan>> cf_plans;
  // embedded boundary
  bool has_eb{false};
  std::map<std::tuple<int, int>, std::unique_ptr<DeviceArray<fvs::EbDesc>>> eb_cache;
  // lazy stable dt: the last max signal speed evaluated per direction (all ranks
  // hold the same value, it comes out of the all-reduce); 0 = none yet
  double last_sig[3]{0.0, 0.0, 0.0};
};

} // namespace

struct fvs_ctx {
  fvs_config cfg{};
  fvs::EqConst eq{};
  fvs::MethodKey method{};
  int device{0};
  cudaStream_t stream{nullptr};
  std::vector<Level> levels;
  long launches{0};
  long cycles{0};
  double time_point{0.0};
  int regrid_frequency{1};
  int cutcell_riemann{FVS_RIEMANN_HLL_EINFELDT};
  int arithmetic{FVS_ARITH_IEEE};
  int ratio[3]{2, 2, 2};
  // stable-dt accumulators
  unsigned long long* d_slots{nullptr};
  unsigned long long* h_slots{nullptr}; // pinned, mapped
  unsigned long long* h_slots_dev{nullptr}; // device address of h_slots
  int* h_fail_dev{nullptr};                 // device address of h_fail
  int* d_fail{nullptr};
  int* h_fail{nullptr}; // pinned
  int next_slot{0};
  // lazy stable dt (numerics.cuh): hints are derived from Level::last_sig and
  // verified after every evaluation; a failed verification repeats the work
  // without a hint
  bool dt_hints{true};
  bool keep_hints_on_upload{false}; // fvs_set_dt_hints(ctx, 2): lets the tests plant a stale hint
  long hint_redos{0};
  // timing
  bool timing{false};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_events;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> free_events;
  double sweep_ms{0.0};
  long sweep_count{0};
  // comm
  Nccl::comm_t comm{nullptr};
  // 1: halo exchange by direct stores into the peers' ghost cells (CUDA IPC + NVLink,
  // FVS_HALO_P2P=1), 0 (default): pack -> grouped ncclSend / ncclRecv -> unpack
  int halo_p2p{-1};
  int* d_barrier{nullptr};
  // mark and repair (SweepArgs::repair_*)
  fvs::RepairEntry* d_repair{nullptr};
  unsigned* d_repair_count{nullptr};
  unsigned repair_cap{kRepairCap}; // FVS_REPAIR_CAP overrides (tests of the overflow path)
  // descriptor staging: a mapped pinned ring the kernels read their descriptor
  // arrays from directly, so building one costs a memcpy on the host and no
  // transfer or synchronisation; rewound whenever the stream is known to be idle
  unsigned char* ring{nullptr};
  unsigned char* ring_dev{nullptr};
  size_t ring_size{0}, ring_used{0};
  std::vector<unsigned char*> ring_retired; // full rings still referenced by queued kernels

  int stencil_width() const { return method.recon == fvs::kReconFirstOrder ? 1 : 2; }
  int ncomp() const { return cfg.rank + 4; }
  int ncons() const { return cfg.rank + 2; }

  Level& level(int l) {
    if (l < 0 || l >= (int)levels.size() || !levels[l].defined) {
      throw Error("level " + std::to_string(l) + " is not defined");
    }
    return levels[l];
  }
  Patch& patch(Level& L, int box) {
    if (box < 0 || box >= (int)L.boxes.size()) throw Error("box index out of range");
    const int loc = L.local_of_global[box];
    if (loc < 0) {
      throw Error("box " + std::to_string(box) + " is owned by rank " +
                  std::to_string(L.owner[box]) + ", not by this rank");
    }
    return L.patches[loc];
  }
  void use_device() { FVS_CUDA(cudaSetDevice(device)); }
};

namespace {

void free_level(Level& L) {
  for (int b = 0; b < kNumBuffers; ++b) {
    for (double* q : L.remote_buf[b])
      if (q) cudaIpcCloseMemHandle(q);
    L.remote_buf[b].clear();
    L.have_remote[b] = false;
  }
  for (Patch& p : L.patches) {
    for (int b = 0; b < kNumBuffers; ++b)
      if (p.buf[b]) cudaFree(p.buf[b]);
    for (int d = 0; d < 3; ++d) {
      if (p.flux[d]) cudaFree(p.flux[d]);
      for (int w = 0; w < 2; ++w)
        if (p.dtmask[w][d]) cudaFree(p.dtmask[w][d]);
      if (p.dtmask_eb[d]) cudaFree(p.dtmask_eb[d]);
    }
    EbPatch& e = p.eb;
    void* single[] = {e.alpha, e.bcent, e.bnorm, e.reference, e.boundary_flux, e.cut_cells,
                      e.complete_mode, e.flags};
    for (void* q : single)
      if (q) cudaFree(q);
    for (int d = 0; d < 3; ++d) {
      void* per_dir[] = {e.beta[d], e.us[d], e.sl[d], e.sr[d], e.ds[d], e.stab[d], e.shl[d],
                         e.shr[d]};
      for (void* q : per_dir)
        if (q) cudaFree(q);
    }
  }
  for (PeerExchange& pe : L.peers) {
    if (pe.sendbuf) cudaFree(pe.sendbuf);
    if (pe.recvbuf) cudaFree(pe.recvbuf);
  }
  for (Level::CfRegister& r : L.cf_regs) {
    if (r.reg) cudaFree(r.reg);
    if (r.open) cudaFree(r.open);
  }
  L = Level();
}

void ensure_buffer(fvs_ctx* ctx, Level& L, int b) {
  if (L.have_buf[b]) return;
  for (Patch& p : L.patches) {
    const size_t bytes = (size_t)p.cells * ctx->ncomp() * sizeof(double);
    FVS_CUDA(cudaMalloc(&p.buf[b], bytes));
    FVS_CUDA(cudaMemsetAsync(p.buf[b], 0, bytes, ctx->stream));
  }
  L.have_buf[b] = true;
  // descriptor arrays are freed only while the stream is idle (kernels in flight read them)
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  L.box_cache.clear();
  L.plane_cache.clear();
}

void ensure_fluxes(fvs_ctx* ctx, Level& L, int dir) {
  bool changed = false;
  for (Patch& p : L.patches) {
    if (!p.flux[dir]) {
      const size_t bytes = (size_t)p.fbox[dir].size() * ctx->ncons() * sizeof(double);
      FVS_CUDA(cudaMalloc(&p.flux[dir], bytes));
      FVS_CUDA(cudaMemsetAsync(p.flux[dir], 0, bytes, ctx->stream));
      changed = true;
    }
  }
  if (changed) {
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    L.box_cache.clear();
    L.plane_cache.clear();
  }
}

// a buffer index different from `a` and `b` (and from the kept backup)
int free_buffer(const Level& L, int a) {
  for (int i = 0; i < kNumBuffers; ++i) {
    if (i != a && i != L.keep) return i;
  }
  throw Error("no free state buffer");
}

enum DescKind { kDescSweep = 0 };

// Patches of a cut-cell level by the flag type of their fab: regular fabs take the
// fused sweep, single-valued (cut) fabs the cut-cell kernels, covered fabs nothing.
enum PatchGroup { kGroupAll = 0, kGroupRegular = 1, kGroupCut = 2 };
bool in_group(const Level& L, const Patch& p, int group) {
  if (group == kGroupAll) return true;
  if (!L.has_eb) return group == kGroupRegular;
  return group == kGroupRegular ? p.eb.fab_type == FVS_FAB_REGULAR
                                : p.eb.fab_type == FVS_FAB_SINGLEVALUED;
}
int group_size(const Level& L, int group) {
  int n = 0;
  for (const Patch& p : L.patches) n += in_group(L, p, group) ? 1 : 0;
  return n;
}

const BoxDesc* box_descs(fvs_ctx* ctx, Level& L, int dir, int src, int dst, bool with_flux,
                         int group = kGroupAll) {
  const int w = ctx->stencil_width();
  auto key = std::make_tuple((with_flux ? 1 : 0) + 2 * group, dir, src, dst, w);
  auto it = L.box_cache.find(key);
  if (it != L.box_cache.end()) return it->second->ptr;
  std::vector<BoxDesc> host;
  for (Patch& p : L.patches) {
    if (!in_group(L, p, group)) continue;
    BoxDesc d{};
    d.src = p.buf[src];
    d.dst = p.buf[dst];
    d.flux = with_flux ? p.flux[dir] : nullptr;
    for (int e = 0; e < 3; ++e) {
      d.n[e] = p.n[e];
      d.fn[e] = p.fbox[dir].len(e);
      d.foff[e] = p.fbox[dir].lo[e] - p.sbox.lo[e];
    }
    d.sc = p.cells;
    d.fsc = p.fbox[dir].size();
    int ulo, uhi, flo, fhi;
    fvs::update_cell_range(p.valid, dir, L.flux_gcw, ulo, uhi);
    fvs::flux_face_range(p.valid, dir, L.scratch_gcw, L.flux_gcw, w, flo, fhi);
    d.ulo = ulo - p.sbox.lo[dir];
    d.uhi = uhi - p.sbox.lo[dir];
    d.flo = flo - p.sbox.lo[dir];
    d.fhi = fhi - p.sbox.lo[dir];
    d.dtmask = p.dtmask[w - 1][dir];
    if (L.has_eb && group == kGroupRegular) {
      // the fused sweep of a regular fab of a cut-cell level: same formulas, but the
      // stable dt covers the faces flo .. fhi (mask index = left cell of the face)
      if (!p.dtmask_eb[dir]) {
        std::vector<unsigned char> mask((size_t)p.n[dir], 0);
        for (int f = d.flo; f <= d.fhi; ++f) mask[(size_t)(f - 1)] = 1;
        FVS_CUDA(cudaMalloc(&p.dtmask_eb[dir], mask.size()));
        FVS_CUDA(upload_sync(p.dtmask_eb[dir], mask.data(), mask.size()));
      }
      d.dtmask = p.dtmask_eb[dir];
    }
    d.alpha = p.eb.alpha;
    d.beta = p.eb.beta[dir];
    d.fab_type = p.eb.fab_type;
    host.push_back(d);
  }
  auto arr = std::make_unique<DeviceArray<BoxDesc>>();
  arr->upload(host);
  const BoxDesc* ptr = arr->ptr;
  L.box_cache[key] = std::move(arr);
  return ptr;
}

int pick_seg_len(int rank, int dir, const int max_n[3], int nboxes) {
  // enough independent pipelines to fill 148 SMs; segments overlap by the
  // stencil, so keep them long when the box already provides the parallelism
  const long long target = 148LL * 2048;
  long long lanes;
  int n = max_n[dir];
  if (dir == 0) {
    lanes = (long long)max_n[1] * max_n[2] * nboxes * 32; // one warp per row
  } else {
    lanes = (long long)max_n[0] * (dir == 1 ? max_n[2] : max_n[1]) * nboxes;
  }
  (void)rank;
  int nseg = (int)std::min<long long>((target + lanes - 1) / lanes, std::max(n / 32, 1));
  nseg = std::max(nseg, 1);
  int seg = (n + nseg - 1) / nseg;
  if (dir == 0) seg = ((seg + 31) / 32) * 32;
  return std::max(seg, 1);
}

// The hint of a lazy stable-dt evaluation: kHintFactor times the last known max
// signal speed of the direction.  Valid iff the evaluation returns at least the
// hint (verify_hint); the max signal speed of a level changes by a fraction of a
// percent from sweep to sweep, so a violated hint is the exception (a redo); the
// start-up phase of a Riemann problem, where the peak overshoots and settles,
// is the place where it happens (bench.py reports the count).
constexpr double kHintFactor = 0.9;
double dt_hint(const fvs_ctx* ctx, const Level& L, int dir) {
  return ctx->dt_hints ? kHintFactor * L.last_sig[dir] : 0.0;
}
// true iff an evaluation with `hint` that returned `sig` is exact: every face
// that was skipped has a signal speed below max(hint, sig) (numerics.cuh)
bool verify_hint(double hint, double sig) { return sig >= hint; }
// new data from outside: what is known about the old state says nothing about it
void drop_hints(const fvs_ctx* ctx, Level& L) {
  if (ctx->keep_hints_on_upload) return;
  for (double& s : L.last_sig) s = 0.0;
}

fvs::SweepArgs sweep_args(fvs_ctx* ctx, Level& L, int dir, double dt, const BoxDesc* boxes,
                          int slot, double hint = 0.0) {
  fvs::SweepArgs A{};
  A.eq = ctx->eq;
  A.eq.dx = L.dx[dir];
  A.eq.limiter2 = ctx->method.limiter;
  A.lambda = dt / L.dx[dir];
  A.lambda_half = 0.5 * dt / L.dx[dir];
  A.boxes = boxes;
  A.nboxes = (int)L.patches.size();
  for (int d = 0; d < 3; ++d) A.max_n[d] = L.max_n[d];
  A.seg_len = pick_seg_len(ctx->cfg.rank, dir, L.max_n, A.nboxes);
  // slot < 0: a kernel without a stable-dt accumulator (general failure flag)
  A.max_signal = ctx->d_slots + std::max(slot, 0);
  A.fail_flag = ctx->d_fail + (slot >= 0 ? slot : kDtSlots);
  A.sig_hint = hint;
  if (!ctx->d_repair) {
    if (const char* e = std::getenv("FVS_REPAIR_CAP")) {
      ctx->repair_cap = (unsigned)std::max(1L, std::atol(e));
    }
    FVS_CUDA(cudaMalloc(&ctx->d_repair, sizeof(fvs::RepairEntry) * ctx->repair_cap));
    FVS_CUDA(cudaMalloc(&ctx->d_repair_count, 2 * sizeof(unsigned)));
    FVS_CUDA(cudaMemsetAsync(ctx->d_repair_count, 0, 2 * sizeof(unsigned), ctx->stream));
  }
  A.repair_list = ctx->d_repair;
  A.repair_count = ctx->d_repair_count;
  A.repair_cap = ctx->repair_cap;
  return A;
}

int take_slot(fvs_ctx* ctx, int count = 1) {
  if (ctx->next_slot + count > kDtSlots) ctx->next_slot = 0;
  const int s = ctx->next_slot;
  ctx->next_slot += count;
  FVS_CUDA(cudaMemsetAsync(ctx->d_slots + s, 0, sizeof(unsigned long long) * count, ctx->stream));
  FVS_CUDA(cudaMemsetAsync(ctx->d_fail + s, 0, sizeof(int) * count, ctx->stream));
  return s;
}

void check_dir(const fvs_ctx* ctx, int dir) {
  if (dir < 0 || dir >= ctx->cfg.rank) throw Error("direction out of range");
}

void check_launch(int rc) {
  if (rc < 0) throw Error("method / dimension combination is not compiled in");
  check_last("launch check #1");
}

// the stream is idle: nothing refers to staged descriptors any more
void rewind_ring(fvs_ctx* ctx) {
  ctx->ring_used = 0;
  for (unsigned char* r : ctx->ring_retired) cudaFreeHost(r);
  ctx->ring_retired.clear();
}
// Copies `bytes` into the descriptor ring and returns their device address.
const void* stage_bytes(fvs_ctx* ctx, const void* data, size_t bytes) {
  const size_t aligned = (std::max<size_t>(bytes, 1) + 255) & ~size_t(255);
  if (ctx->ring_used + aligned > ctx->ring_size) {
    // descriptors staged so far may not have been consumed yet (or not even been
    // handed to a kernel): the full ring is retired, not overwritten
    ctx->ring_retired.push_back(ctx->ring);
    ctx->ring_size = std::max(ctx->ring_size, 2 * aligned);
    FVS_CUDA(cudaHostAlloc(&ctx->ring, ctx->ring_size, cudaHostAllocMapped));
    FVS_CUDA(cudaHostGetDevicePointer(&ctx->ring_dev, ctx->ring, 0));
    ctx->ring_used = 0;
  }
  std::memcpy(ctx->ring + ctx->ring_used, data, bytes);
  const void* dev = ctx->ring_dev + ctx->ring_used;
  ctx->ring_used += aligned;
  return dev;
}
template <typename Desc> const Desc* stage_descs(fvs_ctx* ctx, const std::vector<Desc>& host) {
  return static_cast<const Desc*>(stage_bytes(ctx, host.data(), host.size() * sizeof(Desc)));
}

// Copies the accumulators and the failure flag into mapped pinned host memory
// with plain stores from an SM.  A cudaMemcpyAsync would queue behind whatever
// the device-to-host copy engine is busy with -- a 7.5 GB download of another
// context in the end-to-end pipeline of bench.py delayed each of these 8-byte
// read-backs by up to 160 ms.
// d_fail / h_fail: one Newton-failure flag per accumulator slot plus a general
// one (index kDtSlots) for the kernels that have no slot (cut-cell states).
__global__ void publish_results_kernel(const unsigned long long* d_slots,
                                       unsigned long long* h_slots, int count, const int* d_fail,
                                       int* h_fail, const int* d_general, int* h_general) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    h_slots[i] = d_slots[i];
    h_fail[i] = d_fail[i];
  }
  if (threadIdx.x == 0) *h_general = *d_general;
  __threadfence_system();
}

[[noreturn]] void throw_newton() {
  // include/fub/NewtonIteration.hpp:44-50
  throw Error("Newton Iteration did not converge.");
}

// Reads `count` accumulators back (max over ranks), one synchronisation.
// fails == nullptr: throws if a Newton iteration failed in any of the kernels;
// otherwise the per-slot flags are returned and the caller decides (a sweep
// behind a CFL violation never ran in the reference).
void fetch_slots(fvs_ctx* ctx, int slot, int count, double* out, int* fails = nullptr) {
  if (ctx->comm && ctx->cfg.comm_size > 1) {
    Nccl& n = Nccl::get();
    n.check(n.AllReduce(ctx->d_slots + slot, ctx->d_slots + slot, (size_t)count, Nccl::kDouble,
                        Nccl::kMax, ctx->comm, ctx->stream),
            "ncclAllReduce(max signal)");
    n.check(n.AllReduce(ctx->d_fail, ctx->d_fail, (size_t)kDtSlots + 1, 2 /* ncclInt32 */,
                        Nccl::kMax, ctx->comm, ctx->stream),
            "ncclAllReduce(fail flags)");
  }
  publish_results_kernel<<<1, 32, 0, ctx->stream>>>(
      ctx->d_slots + slot, ctx->h_slots_dev + slot, count, ctx->d_fail + slot,
      ctx->h_fail_dev + slot, ctx->d_fail + kDtSlots, ctx->h_fail_dev + kDtSlots);
  check_last("launch check #2");
  ctx->launches += 1;
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  rewind_ring(ctx);
  const bool general = ctx->h_fail[kDtSlots] != 0;
  if (general) {
    FVS_CUDA(cudaMemsetAsync(ctx->d_fail + kDtSlots, 0, sizeof(int), ctx->stream));
  }
  bool any = general;
  for (int i = 0; i < count; ++i) {
    double v;
    std::memcpy(&v, ctx->h_slots + slot + i, sizeof(double));
    out[i] = v;
    const int f = ctx->h_fail[slot + i] != 0 || general;
    if (fails) fails[i] = f;
    any = any || f;
  }
  if (!fails && any) throw_newton();
}

// FluxMethodAdapter.hpp:78,96: min over faces of dx / max|s|, starting from DBL_MAX
double dt_from_signal(double dx, double max_signal) {
  double dt = std::numeric_limits<double>::max();
  if (max_signal > 0.0) dt = std::min(dt, dx / max_signal);
  return dt;
}

struct EventTimer {
  fvs_ctx* ctx;
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  explicit EventTimer(fvs_ctx* c) : ctx(c) {
    if (!ctx->timing) return;
    if (!ctx->free_events.empty()) {
      ev = ctx->free_events.back();
      ctx->free_events.pop_back();
    } else {
      FVS_CUDA(cudaEventCreate(&ev.first));
      FVS_CUDA(cudaEventCreate(&ev.second));
    }
    FVS_CUDA(cudaEventRecord(ev.first, ctx->stream));
  }
  void stop() {
    if (!ctx->timing) return;
    FVS_CUDA(cudaEventRecord(ev.second, ctx->stream));
    ctx->pending_events.push_back(ev);
  }
};

void drain_events(fvs_ctx* ctx) {
  for (auto& ev : ctx->pending_events) {
    FVS_CUDA(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    FVS_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    ctx->sweep_ms += ms;
    ctx->sweep_count += 1;
    ctx->free_events.push_back(ev);
  }
  ctx->pending_events.clear();
}

// launches the fused sweep cur -> free buffer; does NOT flip.  `hint`: see
// SweepArgs::sig_hint -- the caller verifies it against the fetched result.
int launch_sweep(fvs_ctx* ctx, Level& L, double dt, int dir, int* dst_out, double hint = 0.0,
                 bool force_ieee = false, int group = kGroupAll) {
  const int src = L.cur;
  const int dst = free_buffer(L, src);
  ensure_buffer(ctx, L, dst);
  const BoxDesc* boxes = box_descs(ctx, L, dir, src, dst, false, group);
  const int slot = take_slot(ctx);
  fvs::SweepArgs A = sweep_args(ctx, L, dir, dt, boxes, slot, hint);
  A.nboxes = group_size(L, group);
  if (A.nboxes > 0) {
    EventTimer t(ctx);
    const int rc = (ctx->arithmetic == FVS_ARITH_TOLERANCE && !force_ieee)
                       ? fvs::launch_fused_sweep_tolerance(ctx->method, dir, A, ctx->stream)
                       : fvs::launch_fused_sweep(ctx->method, dir, A, ctx->stream);
    check_launch(rc);
    ctx->launches += rc;
    t.stop();
  }
  *dst_out = dst;
  return slot;
}

// One directional sweep of a cut-cell level, L.cur -> a free buffer (returned in
// *dst_out, not flipped): the regular fabs take the fused sweep (same formulas as on a
// regular level; the stable dt over the cut-cell adapter's face set), the cut fabs are
// copied and then run the cut-cell policies in place on the copy, covered fabs are
// copied.  The stable dt of the level's input state accumulates in the returned slot.
void cc_stable_dt(fvs_ctx*, Level&, int, int, int);
void cc_numeric_fluxes(fvs_ctx*, Level&, double, int, int);
void cc_update(fvs_ctx*, Level&, double, int, int);
void cc_complete(fvs_ctx*, Level&, int);
int eb_sweep(fvs_ctx* ctx, Level& L, double dt, int dir, int* dst_out) {
  const int src = L.cur;
  int dst;
  // cut-cell levels stay in IEEE arithmetic
  const int slot = launch_sweep(ctx, L, dt, dir, &dst, 0.0, /*force_ieee=*/true, kGroupRegular);
  for (Patch& p : L.patches) {
    if (in_group(L, p, kGroupRegular)) continue;
    FVS_CUDA(cudaMemcpyAsync(p.buf[dst], p.buf[src], (size_t)p.cells * ctx->ncomp() * sizeof(double),
                             cudaMemcpyDeviceToDevice, ctx->stream));
  }
  L.cur = dst; // the cut-cell kernels work in place on the copy
  cc_stable_dt(ctx, L, dir, slot, kGroupCut);
  cc_numeric_fluxes(ctx, L, dt, dir, kGroupCut);
  cc_update(ctx, L, dt, dir, kGroupCut);
  cc_complete(ctx, L, kGroupCut);
  L.cur = src;
  *dst_out = dst;
  return slot;
}

// ---- ghost cells ------------------------------------------------------------
const fvs::CopyDesc* copy_descs(fvs_ctx* ctx, Level& L, int buf, int* count) {
  *count = (int)L.local_copies.size();
  auto it = L.copy_cache.find(buf);
  if (it != L.copy_cache.end()) return it->second->ptr;
  std::vector<fvs::CopyDesc> host;
  for (const fvs::HaloCopy& c : L.local_copies) {
    Patch& dp = ctx->patch(L, c.dst_box);
    Patch& sp = ctx->patch(L, c.src_box);
    fvs::CopyDesc d{};
    d.src = sp.buf[buf];
    d.dst = dp.buf[buf];
    for (int e = 0; e < 3; ++e) {
      d.sn[e] = sp.n[e];
      d.dn[e] = dp.n[e];
      d.dlo[e] = c.region.lo[e] - dp.sbox.lo[e];
      d.slo[e] = c.region.lo[e] - c.shift[e] - sp.sbox.lo[e];
      d.len[e] = c.region.len(e);
    }
    d.ssc = sp.cells;
    d.dsc = dp.cells;
    host.push_back(d);
  }
  auto arr = std::make_unique<DeviceArray<fvs::CopyDesc>>();
  arr->upload(host);
  const fvs::CopyDesc* ptr = arr->ptr;
  L.copy_cache[buf] = std::move(arr);
  return ptr;
}

const fvs::PackDesc* pack_descs(fvs_ctx* ctx, Level& L, int peer_index, int buf, bool recv,
                                int* count, long long* max_cells) {
  PeerExchange& pe = L.peers[peer_index];
  const std::vector<fvs::HaloCopy>& list = recv ? pe.recvs : pe.sends;
  *count = (int)list.size();
  *max_cells = 0;
  for (const fvs::HaloCopy& c : list) *max_cells = std::max(*max_cells, c.region.size());
  auto key = std::make_tuple(peer_index, buf, recv ? 1 : 0);
  auto it = L.pack_cache.find(key);
  if (it != L.pack_cache.end()) return it->second->ptr;
  std::vector<fvs::PackDesc> host;
  long long offset = 0;
  for (const fvs::HaloCopy& c : list) {
    Patch& p = ctx->patch(L, recv ? c.dst_box : c.src_box);
    fvs::PackDesc d{};
    d.array = p.buf[buf];
    for (int e = 0; e < 3; ++e) {
      d.n[e] = p.n[e];
      d.lo[e] = c.region.lo[e] - (recv ? 0 : c.shift[e]) - p.sbox.lo[e];
      d.len[e] = c.region.len(e);
    }
    d.sc = p.cells;
    d.offset = offset;
    offset += c.region.size() * ctx->ncomp();
    host.push_back(d);
  }
  auto arr = std::make_unique<DeviceArray<fvs::PackDesc>>();
  arr->upload(host);
  const fvs::PackDesc* ptr = arr->ptr;
  L.pack_cache[key] = std::move(arr);
  return ptr;
}

// FVS_HALO_TRACE=1: per-phase device times of the halo exchange on stderr (events and a
// synchronisation per phase: a diagnosis switch, never on in a timed run)
struct HaloTrace {
  bool on;
  cudaEvent_t ev[4];
  double ms[3] = {0, 0, 0};
  long calls = 0;
  HaloTrace() : on(std::getenv("FVS_HALO_TRACE") != nullptr) {}
  void mark(int i, cudaStream_t s) {
    if (!on) return;
    if (calls == 0 && i == 0) {
      for (cudaEvent_t& e : ev) cudaEventCreate(&e);
    }
    cudaEventRecord(ev[i], s);
  }
  void done(cudaStream_t s, int rank) {
    if (!on) return;
    cudaStreamSynchronize(s);
    for (int i = 0; i < 3; ++i) {
      float t = 0;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      ms[i] += t;
    }
    calls += 1;
    if (calls % 8 == 0) {
      std::fprintf(stderr,
                   "[fvs halo trace] rank %d, %ld exchanges: pack %.3f ms, send/recv %.3f ms, "
                   "local copies + unpack %.3f ms (means)\n",
                   rank, calls, ms[0] / calls, ms[1] / calls, ms[2] / calls);
    }
  }
};

// ---- halo exchange over peer memory -------------------------------------------------------
// One process per GPU, every GPU of the node reaches every other through NVLink / NVSwitch: a
// rank can store its boundary cells straight into the ghost cells of its neighbours' state
// arrays.  The arrays are plain cudaMalloc allocations, exported with cudaIpcGetMemHandle; the
// handles of buffer b travel once per level configuration in one all-reduce (every rank fills
// the slots of its own boxes, the rest is zero).  An exchange is then
//   barrier (every rank has finished the kernels that write its state / read its old ghosts)
//   ONE copy_regions launch: this rank's send regions -> the peers' arrays
//   barrier (the neighbours' stores into this rank's ghost cells have landed)
// where a barrier is a one-word ncclAllReduce on the stream: no pack / unpack kernels, no
// staging buffers, no second pass over the halo through NCCL's FIFOs.
bool p2p_enabled(fvs_ctx* ctx) {
  if (ctx->halo_p2p < 0) {
    // opt-in: on the measured boxes SM stores through the cudaIpc peer mapping run at about
    // 60 GB/s (a 125 MB face in 2 ms, whatever the copy kernel does) where NCCL's own peer
    // mappings reach 312 GB/s (0.4 ms), profiles/r02_tuning.md -- the path is kept
    // bit-identical and tested (tests/test_gpu_multirank.py) for nodes where peer stores run
    // at NVLink speed
    const char* e = std::getenv("FVS_HALO_P2P");
    ctx->halo_p2p = (e && e[0] == '1') ? 1 : 0;
  }
  return ctx->halo_p2p == 1;
}

void stream_barrier(fvs_ctx* ctx) {
  Nccl& n = Nccl::get();
  if (!ctx->d_barrier) {
    FVS_CUDA(cudaMalloc(&ctx->d_barrier, sizeof(int)));
    FVS_CUDA(cudaMemset(ctx->d_barrier, 0, sizeof(int)));
  }
  n.check(n.AllReduce(ctx->d_barrier, ctx->d_barrier, 1, 2 /* ncclInt32 */, Nccl::kMax, ctx->comm,
                      ctx->stream),
          "ncclAllReduce(barrier)");
}

// Collective over the communicator: opens the peers' arrays of buffer `buf` that this rank
// sends to.  Returns false (on every rank alike) if any rank could not open a handle; the
// caller then falls back to the NCCL exchange for good.
bool p2p_open(fvs_ctx* ctx, Level& L, int buf) {
  if (L.have_remote[buf]) return true;
  Nccl& n = Nccl::get();
  const int nb = (int)L.boxes.size();
  constexpr int kWords = (int)(sizeof(cudaIpcMemHandle_t) / sizeof(int));
  // last words: failures, and the buffer index (the ranks march in lockstep: all of them are
  // about to exchange the same buffer)
  std::vector<int> table((size_t)nb * kWords + 2, 0);
  table[(size_t)nb * kWords + 1] = buf;
  for (Patch& p : L.patches) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p.buf[buf]) != cudaSuccess) {
      cudaGetLastError();
      table[(size_t)nb * kWords] = 1;
      continue;
    }
    std::memcpy(&table[(size_t)p.global_index * kWords], &h, sizeof(h));
  }
  int* dev = nullptr;
  FVS_CUDA(cudaMalloc(&dev, table.size() * sizeof(int)));
  FVS_CUDA(cudaMemcpyAsync(dev, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice,
                           ctx->stream));
  n.check(n.AllReduce(dev, dev, table.size(), 2 /* ncclInt32 */, 0 /* ncclSum */, ctx->comm,
                      ctx->stream),
          "ncclAllReduce(ipc handles)");
  FVS_CUDA(cudaMemcpyAsync(table.data(), dev, table.size() * sizeof(int), cudaMemcpyDeviceToHost,
                           ctx->stream));
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  int failures = table[(size_t)nb * kWords];
  if (table[(size_t)nb * kWords + 1] != buf * ctx->cfg.comm_size) failures = 1;
  L.remote_buf[buf].assign(nb, nullptr);
  if (failures == 0) {
    for (PeerExchange& pe : L.peers) {
      for (const fvs::HaloCopy& c : pe.sends) {
        if (L.remote_buf[buf][c.dst_box]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, &table[(size_t)c.dst_box * kWords], sizeof(h));
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          failures = 1;
          break;
        }
        L.remote_buf[buf][c.dst_box] = (double*)q;
      }
      if (failures) break;
    }
  }
  // the decision must be the same everywhere
  FVS_CUDA(cudaMemcpyAsync(dev, &failures, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  n.check(n.AllReduce(dev, dev, 1, 2 /* ncclInt32 */, Nccl::kMax, ctx->comm, ctx->stream),
          "ncclAllReduce(ipc status)");
  FVS_CUDA(cudaMemcpyAsync(&failures, dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(dev);
  if (failures) {
    for (double*& q : L.remote_buf[buf]) {
      if (q) cudaIpcCloseMemHandle(q);
      q = nullptr;
    }
    return false;
  }
  L.have_remote[buf] = true;
  return true;
}

const fvs::CopyDesc* p2p_descs(fvs_ctx* ctx, Level& L, int buf) {
  auto it = L.p2p_cache.find(buf);
  if (it != L.p2p_cache.end()) return it->second->ptr;
  std::vector<fvs::CopyDesc> host;
  L.p2p_max_cells = 0;
  for (PeerExchange& pe : L.peers) {
    for (const fvs::HaloCopy& c : pe.sends) {
      Patch& sp = ctx->patch(L, c.src_box);
      const Box dbox = fvs::state_box(L.boxes[c.dst_box], L.rank, L.scratch_gcw);
      fvs::CopyDesc d{};
      d.src = sp.buf[buf];
      d.dst = L.remote_buf[buf][c.dst_box];
      for (int e = 0; e < 3; ++e) {
        d.sn[e] = sp.n[e];
        d.dn[e] = dbox.len(e);
        d.dlo[e] = c.region.lo[e] - dbox.lo[e];
        d.slo[e] = c.region.lo[e] - c.shift[e] - sp.sbox.lo[e];
        d.len[e] = c.region.len(e);
      }
      d.ssc = sp.cells;
      d.dsc = dbox.size();
      L.p2p_max_cells = std::max(L.p2p_max_cells, c.region.size());
      host.push_back(d);
    }
  }
  L.p2p_count = (int)host.size();
  auto arr = std::make_unique<DeviceArray<fvs::CopyDesc>>();
  arr->upload(host);
  const fvs::CopyDesc* ptr = arr->ptr;
  L.p2p_cache[buf] = std::move(arr);
  return ptr;
}

// Collective: closes the peers' arrays and waits until every rank has done so -- an exported
// allocation must not be freed while a peer still has it open.
void p2p_release(fvs_ctx* ctx, Level& L) {
  bool any = false;
  for (int b = 0; b < kNumBuffers; ++b) {
    any = any || L.have_remote[b];
    for (double* q : L.remote_buf[b])
      if (q) cudaIpcCloseMemHandle(q);
    L.remote_buf[b].clear();
    L.have_remote[b] = false;
  }
  L.p2p_cache.clear();
  if (any && ctx->comm) {
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    stream_barrier(ctx);
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
}

void fill_boundary(fvs_ctx* ctx, Level& L) {
  static thread_local HaloTrace trace;
  const int buf = L.cur;
  trace.mark(0, ctx->stream);
  // every rank of the communicator takes part (the barriers are collectives), also one whose
  // own boxes have no remote neighbour
  if (L.halo_crosses_ranks && ctx->comm && p2p_enabled(ctx)) {
    if (!p2p_open(ctx, L, buf)) {
      ctx->halo_p2p = 0; // on every rank alike
    } else {
      const fvs::CopyDesc* rd = p2p_descs(ctx, L, buf);
      stream_barrier(ctx);
      trace.mark(1, ctx->stream);
      ctx->launches += fvs::launch_copy_regions_wide(rd, L.p2p_count, ctx->ncomp(),
                                                     L.p2p_max_cells, ctx->stream);
      stream_barrier(ctx);
      trace.mark(2, ctx->stream);
      int count;
      const fvs::CopyDesc* descs = copy_descs(ctx, L, buf, &count);
      ctx->launches +=
          fvs::launch_copy_regions(descs, count, ctx->ncomp(), L.max_copy_cells, ctx->stream);
      check_last("launch check #3 (peer-memory halo)");
      trace.mark(3, ctx->stream);
      trace.done(ctx->stream, ctx->cfg.comm_rank);
      return;
    }
  }
  // remote part first so that the transfers overlap the local copies
  if (!L.peers.empty()) {
    if (!ctx->comm) throw Error("level has boxes on other ranks but fvs_comm_init was not called");
    Nccl& n = Nccl::get();
    for (size_t i = 0; i < L.peers.size(); ++i) {
      int count;
      long long max_cells;
      const fvs::PackDesc* descs = pack_descs(ctx, L, (int)i, buf, false, &count, &max_cells);
      ctx->launches += fvs::launch_pack(descs, count, ctx->ncomp(), L.peers[i].sendbuf,
                                        max_cells, 0, ctx->stream);
    }
    trace.mark(1, ctx->stream);
    n.check(n.GroupStart(), "ncclGroupStart");
    for (PeerExchange& pe : L.peers) {
      if (pe.send_elems)
        n.check(n.Send(pe.sendbuf, (size_t)pe.send_elems, Nccl::kDouble, pe.peer, ctx->comm,
                       ctx->stream),
                "ncclSend");
      if (pe.recv_elems)
        n.check(n.Recv(pe.recvbuf, (size_t)pe.recv_elems, Nccl::kDouble, pe.peer, ctx->comm,
                       ctx->stream),
                "ncclRecv");
    }
    n.check(n.GroupEnd(), "ncclGroupEnd");
  } else {
    trace.mark(1, ctx->stream);
  }
  trace.mark(2, ctx->stream);
  int count;
  const fvs::CopyDesc* descs = copy_descs(ctx, L, buf, &count);
  ctx->launches +=
      fvs::launch_copy_regions(descs, count, ctx->ncomp(), L.max_copy_cells, ctx->stream);
  for (size_t i = 0; i < L.peers.size(); ++i) {
    int n_regions;
    long long max_cells;
    const fvs::PackDesc* pd = pack_descs(ctx, L, (int)i, buf, true, &n_regions, &max_cells);
    ctx->launches += fvs::launch_pack(pd, n_regions, ctx->ncomp(), L.peers[i].recvbuf, max_cells,
                                      1, ctx->stream);
  }
  check_last("launch check #3");
  trace.mark(3, ctx->stream);
  trace.done(ctx->stream, ctx->cfg.comm_rank);
}

void apply_bc(fvs_ctx* ctx, Level& L, int bc_index) {
  const BoundaryCondition& bc = L.bcs[bc_index];
  if (L.periodic[bc.dir]) return;
  const int buf = L.cur;
  auto key = std::make_tuple(bc_index, buf);
  auto it = L.bc_cache.find(key);
  if (it == L.bc_cache.end()) {
    std::vector<fvs::BcDesc> host;
    const Box slab = fvs::boundary_slab(ctx->cfg.rank, L.domain, L.periodic, L.scratch_gcw,
                                        bc.dir, bc.side, bc.kind == FVS_BC_REFLECTIVE);
    for (Patch& p : L.patches) {
      const Box fill = fvs::intersect(p.sbox, slab);
      if (!fill.ok()) continue;
      fvs::BcDesc d{};
      d.array = p.buf[buf];
      for (int e = 0; e < 3; ++e) {
        d.n[e] = p.n[e];
        d.lo[e] = fill.lo[e] - p.sbox.lo[e];
        d.len[e] = fill.len(e);
      }
      d.sc = p.cells;
      const int boundary = bc.side == 0 ? L.domain.lo[bc.dir] : L.domain.hi[bc.dir];
      d.src_plane = boundary - p.sbox.lo[bc.dir];
      // reflective: src = 2*lo - 1 - idx (lower side), 2*hi + 1 - idx (upper side)
      const int mirror = bc.side == 0 ? 2 * L.domain.lo[bc.dir] - 1 : 2 * L.domain.hi[bc.dir] + 1;
      d.mirror_sum = mirror - 2 * p.sbox.lo[bc.dir];
      if (d.src_plane < 0 || d.src_plane >= p.n[bc.dir]) continue;
      host.push_back(d);
    }
    auto arr = std::make_unique<DeviceArray<fvs::BcDesc>>();
    arr->upload(host);
    it = L.bc_cache.emplace(key, std::move(arr)).first;
  }
  DeviceArray<fvs::BcDesc>& arr = *it->second;
  if (arr.count == 0) return;
  long long max_cells = 0;
  for (Patch& p : L.patches) max_cells = std::max(max_cells, p.cells);
  ctx->launches += fvs::launch_boundary(arr.ptr, (int)arr.count, ctx->ncomp(), bc.dir, bc.kind,
                                        max_cells, ctx->stream);
  check_last("launch check #4");
}

void apply_boundary_condition(fvs_ctx* ctx, Level& L, int dir) {
  for (size_t i = 0; i < L.bcs.size(); ++i) {
    if (L.bcs[i].dir == dir) apply_bc(ctx, L, (int)i);
  }
}

void fill_ghost_layer(fvs_ctx* ctx, Level& L) {
  fill_boundary(ctx, L);
  for (size_t i = 0; i < L.bcs.size(); ++i) apply_bc(ctx, L, (int)i);
}

// sweeps of one split step: (slot of the direction list, fraction of dt)
struct SweepPlanEntry {
  int slot;
  double dt;
};
void split2(int kind, double dt, int s1, int s2, std::vector<SweepPlanEntry>& out) {
  // src/split_method/{Godunov,Strang,StrangLumped}Splitting.cpp:25-39
  if (kind == FVS_SPLIT_GODUNOV) {
    out.push_back({s1, dt});
    out.push_back({s2, dt});
  } else if (kind == FVS_SPLIT_STRANG) {
    out.push_back({s1, 0.5 * dt});
    out.push_back({s2, 0.5 * dt});
    out.push_back({s2, 0.5 * dt});
    out.push_back({s1, 0.5 * dt});
  } else {
    out.push_back({s1, 0.5 * dt});
    out.push_back({s2, dt});
    out.push_back({s1, 0.5 * dt});
  }
}
std::vector<SweepPlanEntry> split_plan(int rank, int kind, double dt) {
  std::vector<SweepPlanEntry> out;
  if (rank == 1) {
    out.push_back({0, dt});
  } else if (rank == 2) {
    split2(kind, dt, 0, 1, out);
  } else if (kind == FVS_SPLIT_STRANG) {
    // StrangSplitting.cpp:44-69
    const double h = 0.5 * dt;
    for (int s : {0, 1, 2, 2, 1, 0}) out.push_back({s, h});
  } else {
    // SplittingMethod.hpp:53-60: Advance(dt, a1, [a2, a3 recursively])
    std::vector<SweepPlanEntry> outer;
    split2(kind, dt, 0, -1, outer);
    for (const SweepPlanEntry& e : outer) {
      if (e.slot == 0) {
        out.push_back(e);
      } else {
        split2(kind, e.dt, 1, 2, out);
      }
    }
  }
  return out;
}


// ---- embedded boundary -------------------------------------------------------
Box beta_box(const Patch& p, int d) {
  Box b = p.sbox;
  b.hi[d] += 1;
  return b;
}

template <typename T> T* device_alloc_zero(fvs_ctx* ctx, size_t count) {
  T* ptr = nullptr;
  if (count == 0) return ptr;
  FVS_CUDA(cudaMalloc(&ptr, count * sizeof(T)));
  FVS_CUDA(cudaMemsetAsync(ptr, 0, count * sizeof(T), ctx->stream));
  return ptr;
}

// allocates the cut-cell arrays of every owned box; geometry defaults to "all
// regular" (alpha = beta = 1)
void enable_eb(fvs_ctx* ctx, Level& L) {
  if (L.has_eb) return;
  const int rank = ctx->cfg.rank;
  if (rank < 2) throw Error("cut cells need rank 2 or 3");
  if (L.scratch_gcw != L.flux_gcw + ctx->stencil_width()) {
    throw Error("EB levels need scratch_gcw == flux_gcw + stencil width");
  }
  for (int d = 0; d < rank; ++d) ensure_fluxes(ctx, L, d);
  for (Patch& p : L.patches) {
    EbPatch& e = p.eb;
    const size_t cells = (size_t)p.cells;
    e.alpha = device_alloc_zero<double>(ctx, cells);
    e.bcent = device_alloc_zero<double>(ctx, cells * rank);
    e.bnorm = device_alloc_zero<double>(ctx, cells * rank);
    e.reference = device_alloc_zero<double>(ctx, cells * ctx->ncomp());
    e.boundary_flux = device_alloc_zero<double>(ctx, cells * ctx->ncons());
    e.complete_mode = device_alloc_zero<unsigned char>(ctx, cells);
    e.flags = device_alloc_zero<unsigned char>(ctx, cells);
    std::vector<double> ones(cells, 1.0);
    FVS_CUDA(cudaMemcpyAsync(e.alpha, ones.data(), cells * sizeof(double), cudaMemcpyHostToDevice,
                             ctx->stream));
    std::vector<unsigned char> mode(cells, 1);
    FVS_CUDA(cudaMemcpyAsync(e.complete_mode, mode.data(), cells, cudaMemcpyHostToDevice,
                             ctx->stream));
    for (int d = 0; d < rank; ++d) {
      const size_t faces = (size_t)beta_box(p, d).size();
      e.beta[d] = device_alloc_zero<double>(ctx, faces);
      std::vector<double> fones(faces, 1.0);
      FVS_CUDA(cudaMemcpyAsync(e.beta[d], fones.data(), faces * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
      e.us[d] = device_alloc_zero<double>(ctx, 2 * faces);
      e.sl[d] = device_alloc_zero<double>(ctx, 2 * faces);
      e.sr[d] = device_alloc_zero<double>(ctx, 2 * faces);
      e.ds[d] = device_alloc_zero<double>(ctx, 2 * faces);
      const size_t fl = (size_t)p.fbox[d].size() * ctx->ncons();
      e.stab[d] = device_alloc_zero<double>(ctx, fl);
      e.shl[d] = device_alloc_zero<double>(ctx, fl);
      e.shr[d] = device_alloc_zero<double>(ctx, fl);
      FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  L.has_eb = true;
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
  L.box_cache.clear();
  L.eb_cache.clear();
}

// host twin of fub::IsCutCell (src/CutCellData.cpp:225-256), array-relative
bool host_is_cut_cell(int rank, const Patch& p, const double* alpha, const double* const beta[3],
                      int i, int j, int k) {
  const double a = alpha[i + (long long)p.n[0] * (j + (long long)p.n[1] * k)];
  if (a <= 0.0) return false;
  if (a < 1.0) return true;
  for (int d = 0; d < rank; ++d) {
    const long long bn0 = p.n[0] + (d == 0 ? 1 : 0);
    const long long bn1 = p.n[1] + (d == 1 ? 1 : 0);
    const long long left = i + bn0 * (j + bn1 * k);
    const long long step = d == 0 ? 1 : (d == 1 ? bn0 : bn0 * bn1);
    if (beta[d][left] < 1.0 || beta[d][left + step] < 1.0) return true;
  }
  return false;
}

// cutcell::Reconstruction (include/fub/AMReX/cutcell/Reconstruction.hpp:169-193)
// takes the EBCellFlagFab type of every GROWN TILE BOX (tiles 1024000 x 16 x 32,
// ForEachFab.hpp:48-57): all cells regular -> plain CompleteFromCons, all
// covered -> nothing, otherwise masked by alpha > 0.
std::vector<unsigned char> complete_modes(int rank, const Level& L, const Patch& p,
                                          const double* alpha, const double* const beta[3]) {
  std::vector<unsigned char> mode((size_t)p.cells, 1);
  const int tile[3] = {1024000, 16, 32};
  std::vector<std::array<int, 2>> ranges[3];
  for (int d = 0; d < 3; ++d) {
    if (d < rank) {
      ranges[d] = fvs::tile_ranges(p.valid.lo[d], p.valid.hi[d], tile[d]);
      for (auto& r : ranges[d]) {
        if (r[0] == p.valid.lo[d]) r[0] -= L.scratch_gcw;
        if (r[1] == p.valid.hi[d]) r[1] += L.scratch_gcw;
      }
    } else {
      ranges[d].push_back({p.valid.lo[d], p.valid.hi[d]});
    }
  }
  for (const auto& rz : ranges[2])
    for (const auto& ry : ranges[1])
      for (const auto& rx : ranges[0]) {
        long long n = 0, ncov = 0, nreg = 0;
        for (int k = rz[0]; k <= rz[1]; ++k)
          for (int j = ry[0]; j <= ry[1]; ++j)
            for (int i = rx[0]; i <= rx[1]; ++i) {
              const int ai = i - p.sbox.lo[0], aj = j - p.sbox.lo[1], ak = k - p.sbox.lo[2];
              ++n;
              if (alpha[ai + (long long)p.n[0] * (aj + (long long)p.n[1] * ak)] <= 0.0) {
                ++ncov;
              } else if (!host_is_cut_cell(rank, p, alpha, beta, ai, aj, ak)) {
                ++nreg;
              }
            }
        const unsigned char m = nreg == n ? 1 : (ncov == n ? 0 : 2);
        for (int k = rz[0]; k <= rz[1]; ++k)
          for (int j = ry[0]; j <= ry[1]; ++j)
            for (int i = rx[0]; i <= rx[1]; ++i) {
              const int ai = i - p.sbox.lo[0], aj = j - p.sbox.lo[1], ak = k - p.sbox.lo[2];
              mode[(size_t)(ai + (long long)p.n[0] * (aj + (long long)p.n[1] * ak))] = m;
            }
      }
  return mode;
}

const fvs::EbDesc* eb_descs(fvs_ctx* ctx, Level& L, int dir, int buf, int group = kGroupAll) {
  auto key = std::make_tuple(dir, buf + 8 * group);
  auto it = L.eb_cache.find(key);
  if (it != L.eb_cache.end()) return it->second->ptr;
  const int w = ctx->stencil_width();
  std::vector<fvs::EbDesc> host;
  for (Patch& p : L.patches) {
    if (!in_group(L, p, group)) continue;
    fvs::EbDesc d{};
    const EbPatch& e = p.eb;
    d.fab_type = e.fab_type;
    const Box bb = beta_box(p, dir);
    for (int k = 0; k < 3; ++k) {
      d.n[k] = p.n[k];
      d.fn[k] = p.fbox[dir].len(k);
      d.foff[k] = p.fbox[dir].lo[k] - p.sbox.lo[k];
      d.bn[k] = bb.len(k);
    }
    d.sc = p.cells;
    d.fsc = p.fbox[dir].size();
    d.bsc = bb.size();
    int ulo, uhi, flo, fhi;
    fvs::update_cell_range(p.valid, dir, L.flux_gcw, ulo, uhi);
    fvs::flux_face_range(p.valid, dir, L.scratch_gcw, L.flux_gcw, w, flo, fhi);
    d.ulo = ulo - p.sbox.lo[dir];
    d.uhi = uhi - p.sbox.lo[dir];
    d.flo = flo - p.sbox.lo[dir];
    d.fhi = fhi - p.sbox.lo[dir];
    d.state = p.buf[buf];
    d.alpha = e.alpha;
    d.beta = e.beta[dir];
    d.bcent = e.bcent;
    d.bnorm = e.bnorm;
    d.us = e.us[dir];
    d.sl = e.sl[dir];
    d.sr = e.sr[dir];
    d.reference = e.reference;
    d.boundary_flux = e.boundary_flux;
    d.flux = p.flux[dir];
    d.stab = e.stab[dir];
    d.shl = e.shl[dir];
    d.shr = e.shr[dir];
    d.cut_cells = e.cut_cells;
    d.n_cut = e.n_cut;
    d.complete_mode = e.complete_mode;
    host.push_back(d);
  }
  auto arr = std::make_unique<DeviceArray<fvs::EbDesc>>();
  arr->upload(host);
  const fvs::EbDesc* ptr = arr->ptr;
  L.eb_cache[key] = std::move(arr);
  return ptr;
}

fvs::EbArgs eb_args(fvs_ctx* ctx, Level& L, int dir, double dt, int group = kGroupAll) {
  fvs::EbArgs A{};
  A.eq = ctx->eq;
  A.boxes = eb_descs(ctx, L, dir, L.cur, group);
  A.nboxes = group_size(L, group);
  A.rank = ctx->cfg.rank;
  A.dir = dir;
  A.riemann = ctx->cutcell_riemann;
  A.lambda = dt / L.dx[dir];
  A.max_cells = 0;
  A.max_cut = 0;
  for (const Patch& p : L.patches) {
    if (!in_group(L, p, group)) continue;
    // one more plane along dir for the face-centred kernels
    A.max_cells = std::max(A.max_cells, p.cells + p.cells / p.n[dir]);
    A.max_cut = std::max(A.max_cut, p.eb.n_cut);
  }
  A.fail_flag = ctx->d_fail + kDtSlots;
  return A;
}

// cutcell::FluxMethod::ComputeStableDt into accumulator `slot`
void cc_stable_dt(fvs_ctx* ctx, Level& L, int dir, int slot, int group = kGroupAll) {
  const int n = group_size(L, group);
  if (n == 0) return;
  const BoxDesc* boxes = box_descs(ctx, L, dir, L.cur, L.cur, true, group);
  fvs::SweepArgs A = sweep_args(ctx, L, dir, 0.0, boxes, slot);
  A.nboxes = n;
  const int rc = fvs::launch_cc_stable_dt(ctx->method, dir, A, ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
}
// cutcell::FluxMethod::ComputeNumericFluxes: boundary fluxes, (masked) regular
// fluxes, stabilised fluxes
void cc_numeric_fluxes(fvs_ctx* ctx, Level& L, double dt, int dir, int group = kGroupAll) {
  const int n = group_size(L, group);
  if (n == 0) return;
  for (Patch& p : L.patches) { // boundary_fluxes.setVal(0.0), FluxMethod.hpp:173
    if (!in_group(L, p, group)) continue;
    FVS_CUDA(cudaMemsetAsync(p.eb.boundary_flux, 0,
                             (size_t)p.cells * ctx->ncons() * sizeof(double), ctx->stream));
  }
  const fvs::EbArgs E = eb_args(ctx, L, dir, dt, group);
  int rc = fvs::launch_cc_boundary_fluxes(E, ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
  const BoxDesc* boxes = box_descs(ctx, L, dir, L.cur, L.cur, true, group);
  fvs::SweepArgs A = sweep_args(ctx, L, dir, dt, boxes, -1);
  A.nboxes = n;
  rc = fvs::launch_cc_numeric_fluxes(ctx->method, dir, A, ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
  rc = fvs::launch_cc_stabilised_fluxes(E, ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
}
void cc_update(fvs_ctx* ctx, Level& L, double dt, int dir, int group = kGroupAll) {
  if (group_size(L, group) == 0) return;
  const int rc = fvs::launch_cc_update(eb_args(ctx, L, dir, dt, group), ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
}
void cc_complete(fvs_ctx* ctx, Level& L, int group = kGroupAll) {
  if (group_size(L, group) == 0) return;
  const int rc = fvs::launch_cc_complete(eb_args(ctx, L, 0, 0.0, group), ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
}
void cc_pre_advance(fvs_ctx* ctx, Level& L) {
  if (!L.has_eb || group_size(L, kGroupCut) == 0) return;
  const int rc =
      fvs::launch_cc_reference_states(eb_args(ctx, L, 0, 0.0, kGroupCut), ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
}


// ---- coarse-fine ----------------------------------------------------------------
// All five operations enumerate the GLOBAL (fine box, coarse box) pairs in the
// same order on every rank.  A pair whose two boxes live on this rank is
// handled by one kernel on the arrays themselves; otherwise the producing
// rank packs (or computes into) a contiguous staging buffer, the buffers of a
// call travel in one grouped ncclSend / ncclRecv (the halo's pattern), and the
// consuming rank runs the SAME kernel with the staging buffer standing in for
// the remote array -- so the arithmetic, and the result, does not depend on
// the distribution of the boxes.
int coarsen_index(int i, int r) { return i >= 0 ? i / r : -((-i + r - 1) / r); }
Box coarsen_box(const Box& b, const int ratio[3]) {
  Box c;
  for (int d = 0; d < 3; ++d) {
    c.lo[d] = coarsen_index(b.lo[d], ratio[d]);
    c.hi[d] = coarsen_index(b.hi[d], ratio[d]);
  }
  return c;
}
Box refine_box(const Box& b, const int ratio[3]) {
  Box f;
  for (int d = 0; d < 3; ++d) {
    f.lo[d] = b.lo[d] * ratio[d];
    f.hi[d] = (b.hi[d] + 1) * ratio[d] - 1;
  }
  return f;
}
bool box_contains(const Box& b, const int p[3]) {
  return p[0] >= b.lo[0] && p[0] <= b.hi[0] && p[1] >= b.lo[1] && p[1] <= b.hi[1] &&
         p[2] >= b.lo[2] && p[2] <= b.hi[2];
}
fvs::ArrayRef array_ref(double* p, const Box& b, long long sc) {
  fvs::ArrayRef a{};
  a.p = p;
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = b.lo[d];
    a.n[d] = b.len(d);
  }
  a.sc = sc;
  return a;
}
void effective_ratio(const fvs_ctx* ctx, int ratio[3]) {
  for (int d = 0; d < 3; ++d) ratio[d] = d < ctx->cfg.rank ? ctx->ratio[d] : 1;
}
double cell_volume(const fvs_ctx* ctx, const Level& L) {
  const int rank = ctx->cfg.rank;
  return L.dx[0] * (rank > 1 ? L.dx[1] : 1.0) * (rank > 2 ? L.dx[2] : 1.0);
}
Patch* local_patch(Level& L, int global_box) {
  const int loc = L.local_of_global[(size_t)global_box];
  return loc < 0 ? nullptr : &L.patches[(size_t)loc];
}

// staging buffers and the transfers of one coarse-fine call.  With a plan the
// buffers are persistent (owned by the plan) and the transfer list is recorded in it;
// without one they are stream-ordered temporaries of this call.
struct Transfers {
  fvs_ctx* ctx;
  CfPlan* plan;
  std::vector<CfXfer> list;
  std::vector<double*> owned;
  explicit Transfers(fvs_ctx* c, CfPlan* p = nullptr) : ctx(c), plan(p) {}
  Transfers(const Transfers&) = delete;
  ~Transfers() {
    // stream ordered: the memory is recycled once the work queued so far is done
    for (double* p : owned) cudaFreeAsync(p, ctx->stream);
  }
  double* buffer(size_t count) {
    double* p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(double);
    if (plan) {
      FVS_CUDA(cudaMalloc(&p, bytes));
      plan->owned.push_back(p);
    } else {
      FVS_CUDA(cudaMallocAsync(&p, bytes, ctx->stream));
      owned.push_back(p);
    }
    return p;
  }
  // registers a transfer this rank takes part in; returns its local buffer
  double* add(int src, int dst, size_t count) {
    double* p = buffer(count);
    list.push_back({src, dst, count, p});
    if (plan) plan->xfers.push_back({src, dst, count, p});
    return p;
  }
  void run() { run_transfers(ctx, list); }
  static void run_transfers(fvs_ctx* ctx, const std::vector<CfXfer>& list) {
    if (list.empty()) return;
    if (!ctx->comm) throw Error("levels span several ranks but fvs_comm_init was not called");
    Nccl& n = Nccl::get();
    const int me = ctx->cfg.comm_rank;
    n.check(n.GroupStart(), "ncclGroupStart");
    for (const CfXfer& x : list) {
      if (x.src == me) {
        n.check(n.Send(x.ptr, x.count, Nccl::kDouble, x.dst, ctx->comm, ctx->stream), "ncclSend");
      } else {
        n.check(n.Recv(x.ptr, x.count, Nccl::kDouble, x.src, ctx->comm, ctx->stream), "ncclRecv");
      }
    }
    n.check(n.GroupEnd(), "ncclGroupEnd");
  }
};

// descriptor array of a plan: persistent device copy
template <typename Desc> const Desc* persist_descs(CfPlan& plan, const std::vector<Desc>& host) {
  if (host.empty()) return nullptr;
  Desc* dev = nullptr;
  FVS_CUDA(cudaMalloc(&dev, host.size() * sizeof(Desc)));
  plan.owned.push_back(dev);
  FVS_CUDA(upload_sync(dev, host.data(), host.size() * sizeof(Desc)));
  return dev;
}

// copies `region` (ncomp components) between two arrays; used to pack a region
// into a staging buffer (whose box is the region itself) and to unpack it
struct RegionCopies {
  std::vector<fvs::CopyDesc> descs;
  long long max_cells{1};
  void add(const fvs::ArrayRef& dst, const fvs::ArrayRef& src, const Box& region) {
    fvs::CopyDesc d{};
    d.src = src.p;
    d.dst = dst.p;
    for (int e = 0; e < 3; ++e) {
      d.sn[e] = src.n[e];
      d.dn[e] = dst.n[e];
      d.slo[e] = region.lo[e] - src.lo[e];
      d.dlo[e] = region.lo[e] - dst.lo[e];
      d.len[e] = region.len(e);
    }
    d.ssc = src.sc;
    d.dsc = dst.sc;
    max_cells = std::max(max_cells, region.size());
    descs.push_back(d);
  }
  void run(fvs_ctx* ctx, int ncomp, Transfers&) {
    if (descs.empty()) return;
    const fvs::CopyDesc* dev = stage_descs(ctx, descs);
    ctx->launches += fvs::launch_copy_regions(dev, (int)descs.size(), ncomp, max_cells, ctx->stream);
  }
  void persist_packs(CfPlan& plan) {
    plan.packs = persist_descs(plan, descs);
    plan.n_packs = (int)descs.size();
    plan.packs_max = max_cells;
  }
  void persist_unpacks(CfPlan& plan) {
    plan.unpacks = persist_descs(plan, descs);
    plan.n_unpacks = (int)descs.size();
    plan.unpacks_max = max_cells;
  }
};
void run_copies(fvs_ctx* ctx, const void* descs, int n, long long max_cells, int ncomp) {
  if (n == 0) return;
  ctx->launches += fvs::launch_copy_regions(static_cast<const fvs::CopyDesc*>(descs), n, ncomp,
                                            max_cells, ctx->stream);
}
template <typename Desc>
const Desc* upload_descs(fvs_ctx* ctx, const std::vector<Desc>& host, Transfers&) {
  return stage_descs(ctx, host);
}
CfPlan* find_plan(Level& fine, int op, int dir, int fine_buf, int coarse_buf) {
  auto it = fine.cf_plans.find(std::make_tuple(op, dir, fine_buf, coarse_buf));
  return it == fine.cf_plans.end() ? nullptr : it->second.get();
}
CfPlan* new_plan(Level& fine, int op, int dir, int fine_buf, int coarse_buf) {
  auto plan = std::make_shared<CfPlan>();
  CfPlan* raw = plan.get();
  fine.cf_plans[std::make_tuple(op, dir, fine_buf, coarse_buf)] = std::move(plan);
  return raw;
}

// plane of coarse faces of one side of a fine box, and its Reflux mask
// (ClearInternalBorders: closed where the outside cell lies under a fine box
// or outside the domain)
Box register_plane(const Box& fine_valid, const int ratio[3], int dir, int side) {
  Box plane = coarsen_box(fine_valid, ratio);
  plane.lo[dir] = plane.hi[dir] = side == 0 ? plane.lo[dir] : plane.hi[dir] + 1;
  return plane;
}
// Periodic images of a level's boxes: the index shifts (multiples of the domain
// length in the periodic directions, the zero shift first) under which a box of
// the level is seen from across a periodic boundary.  amrex::FluxRegister::
// CrseInit / Reflux and FillPatchTwoLevels copy with geom.periodicity().
std::vector<std::array<int, 3>> period_shifts(const Level& L, int rank) {
  std::vector<std::array<int, 3>> out;
  out.push_back({0, 0, 0});
  int r[3];
  for (int d = 0; d < 3; ++d) r[d] = (d < rank && L.periodic[d]) ? 1 : 0;
  for (int sz = -r[2]; sz <= r[2]; ++sz)
    for (int sy = -r[1]; sy <= r[1]; ++sy)
      for (int sx = -r[0]; sx <= r[0]; ++sx) {
        if (sx == 0 && sy == 0 && sz == 0) continue;
        out.push_back({sx * L.domain.len(0), sy * L.domain.len(1), sz * L.domain.len(2)});
      }
  return out;
}
fvs::ArrayRef shifted_ref(fvs::ArrayRef a, const std::array<int, 3>& s) {
  for (int d = 0; d < 3; ++d) a.lo[d] += s[(size_t)d];
  return a;
}
std::vector<unsigned char> register_open(const Level& fine, const Level& coarse, const int ratio[3],
                                         const Box& faces, int dir, int side) {
  std::vector<unsigned char> open((size_t)faces.size(), 1);
  size_t n = 0;
  const int len = coarse.domain.len(dir);
  for (int k = faces.lo[2]; k <= faces.hi[2]; ++k)
    for (int j = faces.lo[1]; j <= faces.hi[1]; ++j)
      for (int i = faces.lo[0]; i <= faces.hi[0]; ++i, ++n) {
        int out[3] = {i, j, k};
        if (side == 0) out[dir] -= 1;
        if (coarse.periodic[dir]) { // the outside cell across a periodic boundary
          if (out[dir] < coarse.domain.lo[dir]) out[dir] += len;
          if (out[dir] > coarse.domain.hi[dir]) out[dir] -= len;
        }
        for (const Box& g : fine.boxes) {
          if (box_contains(coarsen_box(g, ratio), out)) open[n] = 0;
        }
        if (!box_contains(coarse.domain, out)) open[n] = 0;
      }
  return open;
}
// FluxRegister of `fine` towards `coarse`: one plane per side of every OWNED
// fine box
void define_flux_register(fvs_ctx* ctx, Level& fine, Level& coarse, int coarse_index) {
  FVS_CUDA(cudaStreamSynchronize(ctx->stream)); // registers and plans may still be in use
  for (Level::CfRegister& r : fine.cf_regs) {
    if (r.reg) cudaFree(r.reg);
    if (r.open) cudaFree(r.open);
  }
  fine.cf_regs.clear();
  fine.cf_plans.clear();
  int ratio[3];
  effective_ratio(ctx, ratio);
  const int rank = ctx->cfg.rank;
  for (size_t f = 0; f < fine.patches.size(); ++f) {
    for (int d = 0; d < rank; ++d)
      for (int side = 0; side < 2; ++side) {
        Level::CfRegister r;
        r.patch = (int)f;
        r.dir = d;
        r.side = side;
        r.plane = register_plane(fine.patches[f].valid, ratio, d, side);
        const std::vector<unsigned char> open = register_open(fine, coarse, ratio, r.plane, d, side);
        r.reg = device_alloc_zero<double>(ctx, open.size() * ctx->ncons());
        FVS_CUDA(cudaMalloc(&r.open, open.size()));
        FVS_CUDA(upload_sync(r.open, open.data(), open.size()));
        fine.cf_regs.push_back(r);
      }
  }
  fine.cf_coarse = coarse_index;
}
Level::CfRegister* find_register(Level& fine, int local_patch_index, int dir, int side) {
  for (Level::CfRegister& r : fine.cf_regs) {
    if (r.patch == local_patch_index && r.dir == dir && r.side == side) return &r;
  }
  return nullptr;
}

void run_registers(fvs_ctx* ctx, const CfPlan& plan, int op, int dir, double mult, double vol) {
  if (plan.n_main == 0) return;
  int ratio[3];
  effective_ratio(ctx, ratio);
  ctx->launches += fvs::launch_flux_register(static_cast<const fvs::RegisterDesc*>(plan.main),
                                             plan.n_main, op, ctx->ncons(), dir, ratio, mult, vol,
                                             plan.main_max, ctx->stream);
  check_last("launch check #5");
}
void set_main(CfPlan& plan, const std::vector<fvs::RegisterDesc>& host, long long max_faces) {
  plan.main = persist_descs(plan, host);
  plan.n_main = (int)host.size();
  plan.main_max = max_faces;
}
enum CfOp { kOpFineAdd = 0, kOpCrseInit = 1, kOpReflux = 2, kOpAverageDown = 3, kOpInterp = 4 };

// FineAdd: registers and fine fluxes are both local to the owner of the fine box
void fine_add(fvs_ctx* ctx, Level& fine, int dir, double mult) {
  CfPlan* plan = find_plan(fine, kOpFineAdd, dir, 0, 0);
  if (!plan) {
    plan = new_plan(fine, kOpFineAdd, dir, 0, 0);
    std::vector<fvs::RegisterDesc> host;
    long long max_faces = 1;
    for (const Level::CfRegister& r : fine.cf_regs) {
      if (r.dir != dir) continue;
      Patch& p = fine.patches[(size_t)r.patch];
      fvs::RegisterDesc d{};
      d.reg = array_ref(r.reg, r.plane, r.plane.size());
      d.open = r.open;
      d.side = r.side;
      d.other = array_ref(p.flux[dir], p.fbox[dir], p.fbox[dir].size());
      for (int e = 0; e < 3; ++e) {
        d.lo[e] = r.plane.lo[e];
        d.len[e] = r.plane.len(e);
      }
      max_faces = std::max(max_faces, r.plane.size());
      host.push_back(d);
    }
    set_main(*plan, host, max_faces);
  }
  run_registers(ctx, *plan, 0, dir, mult, 1.0);
}

// CrseInit(ADD): the coarse fluxes on the register planes travel to the owner
// of the fine box
void crse_init(fvs_ctx* ctx, Level& fine, Level& coarse, int dir, double mult) {
  const int ncons = ctx->ncons();
  CfPlan* plan = find_plan(fine, kOpCrseInit, dir, 0, 0);
  if (!plan) {
    plan = new_plan(fine, kOpCrseInit, dir, 0, 0);
    int ratio[3];
    effective_ratio(ctx, ratio);
    const int me = ctx->cfg.comm_rank;
    Transfers xf(ctx, plan);
    RegionCopies packs;
    std::vector<fvs::RegisterDesc> host;
    long long max_faces = 1;
    const std::vector<std::array<int, 3>> shifts = period_shifts(coarse, ctx->cfg.rank);
    for (size_t f = 0; f < fine.boxes.size(); ++f)
      for (int side = 0; side < 2; ++side) {
        const Box plane = register_plane(fine.boxes[f], ratio, dir, side);
        for (size_t c = 0; c < coarse.boxes.size(); ++c)
          for (const std::array<int, 3>& shift : shifts) {
            // the coarse box, or its image across a periodic boundary
            const Box image = fvs::shifted(coarse.boxes[c], shift.data());
            Box nodes = image;
            nodes.hi[dir] += 1;
            const Box faces = fvs::intersect(plane, nodes);
            if (!faces.ok()) continue;
            const int rf = fine.owner[f], rc = coarse.owner[c];
            if (rf != me && rc != me) continue;
            Patch* cp = local_patch(coarse, (int)c);
            fvs::ArrayRef flux_ref;
            if (rf == me && rc == me) {
              flux_ref =
                  shifted_ref(array_ref(cp->flux[dir], cp->fbox[dir], cp->fbox[dir].size()), shift);
            } else {
              double* buf = xf.add(rc, rf, (size_t)faces.size() * ncons);
              flux_ref = array_ref(buf, faces, faces.size());
              if (rc == me) {
                packs.add(flux_ref,
                          shifted_ref(array_ref(cp->flux[dir], cp->fbox[dir], cp->fbox[dir].size()),
                                      shift),
                          faces);
                continue;
              }
            }
            Level::CfRegister* r = find_register(fine, fine.local_of_global[f], dir, side);
            if (!r) throw Error("flux register is not defined (fvs_reset_coarse_fine_fluxes)");
            fvs::RegisterDesc d{};
            d.reg = array_ref(r->reg, r->plane, r->plane.size());
            d.open = r->open;
            d.side = side;
            d.other = flux_ref;
            for (int e = 0; e < 3; ++e) {
              d.box_lo[e] = image.lo[e];
              d.box_hi[e] = image.hi[e];
              d.lo[e] = faces.lo[e];
              d.len[e] = faces.len(e);
            }
            max_faces = std::max(max_faces, faces.size());
            host.push_back(d);
          }
      }
    packs.persist_packs(*plan);
    set_main(*plan, host, max_faces);
  }
  run_copies(ctx, plan->packs, plan->n_packs, plan->packs_max, ncons);
  Transfers::run_transfers(ctx, plan->xfers);
  run_registers(ctx, *plan, 1, dir, mult, 1.0);
}

// Reflux: the registers travel to the owners of the coarse boxes
void reflux(fvs_ctx* ctx, Level& fine, Level& coarse, int dir, double scale) {
  const int ncons = ctx->ncons();
  const double vol = cell_volume(ctx, coarse);
  CfPlan* plan = find_plan(fine, kOpReflux, dir, 0, coarse.cur);
  if (!plan) {
    plan = new_plan(fine, kOpReflux, dir, 0, coarse.cur);
    int ratio[3];
    effective_ratio(ctx, ratio);
    const int me = ctx->cfg.comm_rank;
    Transfers xf(ctx, plan);
    RegionCopies packs;
    std::vector<fvs::RegisterDesc> host;
    long long max_faces = 1;
    const std::vector<std::array<int, 3>> shifts = period_shifts(coarse, ctx->cfg.rank);
    for (size_t f = 0; f < fine.boxes.size(); ++f)
      for (int side = 0; side < 2; ++side) {
        const Box plane = register_plane(fine.boxes[f], ratio, dir, side);
        for (size_t c = 0; c < coarse.boxes.size(); ++c)
          for (const std::array<int, 3>& shift : shifts) {
            // faces of the plane whose OUTSIDE cell belongs to coarse box c (or to
            // its image across a periodic boundary)
            const Box image = fvs::shifted(coarse.boxes[c], shift.data());
            Box cells = image;
            if (side == 0) {
              cells.lo[dir] += 1;
              cells.hi[dir] += 1;
            }
            const Box faces = fvs::intersect(plane, cells);
            if (!faces.ok()) continue;
            const int rf = fine.owner[f], rc = coarse.owner[c];
            if (rf != me && rc != me) continue;
            Patch* cp = local_patch(coarse, (int)c);
            fvs::ArrayRef reg_ref;
            const unsigned char* open = nullptr;
            if (rf == me && rc == me) {
              Level::CfRegister* r = find_register(fine, fine.local_of_global[f], dir, side);
              if (!r) throw Error("flux register is not defined (fvs_reset_coarse_fine_fluxes)");
              reg_ref = array_ref(r->reg, r->plane, r->plane.size());
              open = r->open;
            } else {
              double* buf = xf.add(rf, rc, (size_t)faces.size() * ncons);
              reg_ref = array_ref(buf, faces, faces.size());
              if (rf == me) {
                Level::CfRegister* r = find_register(fine, fine.local_of_global[f], dir, side);
                if (!r) throw Error("flux register is not defined (fvs_reset_coarse_fine_fluxes)");
                packs.add(reg_ref, array_ref(r->reg, r->plane, r->plane.size()), faces);
                continue;
              }
              const std::vector<unsigned char> mask =
                  register_open(fine, coarse, ratio, faces, dir, side);
              unsigned char* dmask = nullptr;
              FVS_CUDA(cudaMalloc(&dmask, std::max<size_t>(mask.size(), 1)));
              plan->owned.push_back(dmask);
              FVS_CUDA(upload_sync(dmask, mask.data(), mask.size()));
              open = dmask;
            }
            fvs::RegisterDesc d{};
            d.reg = reg_ref;
            d.open = open;
            d.side = side;
            d.other = shifted_ref(array_ref(cp->buf[coarse.cur], cp->sbox, cp->cells), shift);
            for (int e = 0; e < 3; ++e) {
              d.box_lo[e] = image.lo[e];
              d.box_hi[e] = image.hi[e];
              d.lo[e] = faces.lo[e];
              d.len[e] = faces.len(e);
            }
            max_faces = std::max(max_faces, faces.size());
            host.push_back(d);
          }
      }
    packs.persist_packs(*plan);
    set_main(*plan, host, max_faces);
  }
  run_copies(ctx, plan->packs, plan->n_packs, plan->packs_max, ncons);
  Transfers::run_transfers(ctx, plan->xfers);
  run_registers(ctx, *plan, 2, dir, scale, vol);
}

// amrex::average_down (cut-cell levels: amrex::EB_average_down, weighted with the
// fine volume fractions): the owner of the fine box averages, the result travels
void average_down(fvs_ctx* ctx, Level& fine, Level& coarse) {
  const int ncons = ctx->ncons();
  int ratio[3];
  effective_ratio(ctx, ratio);
  CfPlan* plan = find_plan(fine, kOpAverageDown, 0, fine.cur, coarse.cur);
  if (!plan) {
    plan = new_plan(fine, kOpAverageDown, 0, fine.cur, coarse.cur);
    const int me = ctx->cfg.comm_rank;
    Transfers xf(ctx, plan);
    RegionCopies unpacks;
    std::vector<fvs::AvgDownDesc> host;
    long long max_cells = 1;
    for (size_t f = 0; f < fine.boxes.size(); ++f) {
      const Box cf = coarsen_box(fine.boxes[f], ratio);
      for (size_t c = 0; c < coarse.boxes.size(); ++c) {
        const Box region = fvs::intersect(cf, coarse.boxes[c]);
        if (!region.ok()) continue;
        const int rf = fine.owner[f], rc = coarse.owner[c];
        if (rf != me && rc != me) continue;
        Patch* fp = local_patch(fine, (int)f);
        Patch* cp = local_patch(coarse, (int)c);
        fvs::ArrayRef target;
        if (rf == me && rc == me) {
          target = array_ref(cp->buf[coarse.cur], cp->sbox, cp->cells);
        } else {
          double* buf = xf.add(rf, rc, (size_t)region.size() * ncons);
          target = array_ref(buf, region, region.size());
          if (rc == me) {
            unpacks.add(array_ref(cp->buf[coarse.cur], cp->sbox, cp->cells), target, region);
            continue;
          }
        }
        fvs::AvgDownDesc d{};
        d.fine = array_ref(fp->buf[fine.cur], fp->sbox, fp->cells);
        d.crse = target;
        d.vfrac = (fine.has_eb && fp->eb.fab_type == FVS_FAB_SINGLEVALUED) ? fp->eb.alpha : nullptr;
        for (int e = 0; e < 3; ++e) {
          d.lo[e] = region.lo[e];
          d.len[e] = region.len(e);
        }
        max_cells = std::max(max_cells, region.size());
        host.push_back(d);
      }
    }
    plan->main = persist_descs(*plan, host);
    plan->n_main = (int)host.size();
    plan->main_max = max_cells;
    unpacks.persist_unpacks(*plan);
  }
  if (plan->n_main > 0) {
    ctx->launches += fvs::launch_average_down(static_cast<const fvs::AvgDownDesc*>(plan->main),
                                              plan->n_main, ncons, ratio, plan->main_max, ctx->stream);
    check_last("launch check #6");
  }
  Transfers::run_transfers(ctx, plan->xfers);
  run_copies(ctx, plan->unpacks, plan->n_unpacks, plan->unpacks_max, ncons);
}

// FillPatchTwoLevels with pc_interp: the coarse parents travel to the owner of
// the fine box
void pc_interp(fvs_ctx* ctx, Level& fine, Level& coarse, bool skip_valid = true) {
  const int ncomp = ctx->ncomp();
  int ratio[3];
  effective_ratio(ctx, ratio);
  CfPlan* plan = find_plan(fine, kOpInterp, skip_valid ? 1 : 0, fine.cur, coarse.cur);
  if (!plan) {
    plan = new_plan(fine, kOpInterp, skip_valid ? 1 : 0, fine.cur, coarse.cur);
    const int me = ctx->cfg.comm_rank;
    Transfers xf(ctx, plan);
    RegionCopies packs;
    std::vector<fvs::InterpDesc> host;
    long long max_cells = 1;
    // ghost cells beyond a periodic boundary take their parent from the coarse
    // level's periodic image (the fill domain is grown in the periodic directions)
    const std::vector<std::array<int, 3>> shifts = period_shifts(coarse, ctx->cfg.rank);
    Box fill_domain = fine.domain;
    for (int d = 0; d < ctx->cfg.rank; ++d) {
      if (fine.periodic[d]) {
        fill_domain.lo[d] -= fine.scratch_gcw;
        fill_domain.hi[d] += fine.scratch_gcw;
      }
    }
    for (size_t f = 0; f < fine.boxes.size(); ++f) {
      const Box fsbox = fvs::state_box(fine.boxes[f], ctx->cfg.rank, fine.scratch_gcw);
      for (size_t c = 0; c < coarse.boxes.size(); ++c)
        for (const std::array<int, 3>& shift : shifts) {
          const Box image = fvs::shifted(coarse.boxes[c], shift.data());
          const Box region =
              fvs::intersect(fvs::intersect(fsbox, refine_box(image, ratio)), fill_domain);
          if (!region.ok()) continue;
          const int rf = fine.owner[f], rc = coarse.owner[c];
          if (rf != me && rc != me) continue;
          Patch* fp = local_patch(fine, (int)f);
          Patch* cp = local_patch(coarse, (int)c);
          fvs::ArrayRef source;
          if (rf == me && rc == me) {
            source = shifted_ref(array_ref(cp->buf[coarse.cur], cp->sbox, cp->cells), shift);
          } else {
            const Box parents = coarsen_box(region, ratio);
            double* buf = xf.add(rc, rf, (size_t)parents.size() * ncomp);
            source = array_ref(buf, parents, parents.size());
            if (rc == me) {
              packs.add(source,
                        shifted_ref(array_ref(cp->buf[coarse.cur], cp->sbox, cp->cells), shift),
                        parents);
              continue;
            }
          }
          fvs::InterpDesc d{};
          d.fine = array_ref(fp->buf[fine.cur], fp->sbox, fp->cells);
          d.crse = source;
          for (int e = 0; e < 3; ++e) {
            d.lo[e] = region.lo[e];
            d.len[e] = region.len(e);
            // skip_valid == false (regrid): an empty "own box" -- every cell is filled
            d.vlo[e] = skip_valid ? fine.boxes[f].lo[e] : 1;
            d.vhi[e] = skip_valid ? fine.boxes[f].hi[e] : 0;
          }
          max_cells = std::max(max_cells, region.size());
          host.push_back(d);
        }
    }
    packs.persist_packs(*plan);
    plan->main = persist_descs(*plan, host);
    plan->n_main = (int)host.size();
    plan->main_max = max_cells;
  }
  run_copies(ctx, plan->packs, plan->n_packs, plan->packs_max, ncomp);
  Transfers::run_transfers(ctx, plan->xfers);
  if (plan->n_main > 0) {
    ctx->launches += fvs::launch_pc_interp(static_cast<const fvs::InterpDesc*>(plan->main),
                                           plan->n_main, ncomp, ratio, plan->main_max, ctx->stream);
    check_last("launch check #7");
  }
}

// ParallelCopy of the old level's valid cells into the new level's scratch arrays
// (valid and ghost cells, periodic images included) -- the fine part of the
// FillPatchTwoLevels behind GriddingAlgorithm::RemakeLevel.
void copy_from_old_level(fvs_ctx* ctx, Level& L, Level& old) {
  const int me = ctx->cfg.comm_rank;
  const int ncomp = ctx->ncomp();
  Transfers xf(ctx);
  RegionCopies packs, unpacks, local;
  const std::vector<std::array<int, 3>> shifts = period_shifts(old, ctx->cfg.rank);
  for (size_t f = 0; f < L.boxes.size(); ++f) {
    const Box sbox = fvs::state_box(L.boxes[f], ctx->cfg.rank, L.scratch_gcw);
    for (size_t o = 0; o < old.boxes.size(); ++o)
      for (const std::array<int, 3>& shift : shifts) {
        const Box image = fvs::shifted(old.boxes[o], shift.data());
        const Box region = fvs::intersect(sbox, image);
        if (!region.ok()) continue;
        const int rn = L.owner[f], ro = old.owner[o];
        if (rn != me && ro != me) continue;
        Patch* np = local_patch(L, (int)f);
        Patch* op = local_patch(old, (int)o);
        if (rn == me && ro == me) {
          local.add(array_ref(np->buf[L.cur], np->sbox, np->cells),
                    shifted_ref(array_ref(op->buf[old.cur], op->sbox, op->cells), shift), region);
          continue;
        }
        double* buf = xf.add(ro, rn, (size_t)region.size() * ncomp);
        const fvs::ArrayRef stage = array_ref(buf, region, region.size());
        if (ro == me) {
          packs.add(stage, shifted_ref(array_ref(op->buf[old.cur], op->sbox, op->cells), shift), region);
        } else {
          unpacks.add(array_ref(np->buf[L.cur], np->sbox, np->cells), stage, region);
        }
      }
  }
  local.run(ctx, ncomp, xf);
  packs.run(ctx, ncomp, xf);
  xf.run();
  unpacks.run(ctx, ncomp, xf);
}

} // namespace

// =============================================================================
extern "C" {

const char* fvs_last_error(void) { return g_error.c_str(); }
int fvs_version(void) { return FVS_VERSION; }

int fvs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fvs_create(const fvs_config* config, fvs_ctx** out) {
  return guarded([&] {
    if (!config || !out) throw Error("null argument");
    if (config->rank < 1 || config->rank > 3) throw Error("rank must be 1, 2 or 3");
    if (!(config->gamma > 1.0)) throw Error("gamma must be > 1");
    // 1 / (gamma - 1) is a windowed divisor of the branch-free divisions (fastmath.cuh, divw)
    if (!(config->gamma - 1.0 >= 0x1p-60 && config->gamma <= 0x1p60)) {
      throw Error("gamma - 1 must lie in [2^-60, 2^60]");
    }
    if (fvs_device_count() == 0) {
      throw Error("no CUDA device: libfvs_b200 has no CPU fallback");
    }
    std::unique_ptr<fvs_ctx> ctx(new fvs_ctx);
    ctx->cfg = *config;
    if (ctx->cfg.comm_size < 1) ctx->cfg.comm_size = 1;
    if (ctx->cfg.device < 0) {
      FVS_CUDA(cudaGetDevice(&ctx->device));
    } else {
      ctx->device = ctx->cfg.device;
    }
    ctx->use_device();
    ctx->eq.gamma = config->gamma;
    ctx->eq.gamma_minus_1_inv = 1.0 / (config->gamma - 1.0);
    ctx->eq.gamma_minus_1 = config->gamma - 1.0;
    FVS_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    FVS_CUDA(cudaMalloc(&ctx->d_slots, sizeof(unsigned long long) * kDtSlots));
    FVS_CUDA(cudaMemset(ctx->d_slots, 0, sizeof(unsigned long long) * kDtSlots));
    FVS_CUDA(cudaHostAlloc(&ctx->h_slots, sizeof(unsigned long long) * kDtSlots,
                           cudaHostAllocMapped));
    FVS_CUDA(cudaHostGetDevicePointer(&ctx->h_slots_dev, ctx->h_slots, 0));
    FVS_CUDA(cudaMalloc(&ctx->d_fail, sizeof(int) * (kDtSlots + 1)));
    FVS_CUDA(cudaMemset(ctx->d_fail, 0, sizeof(int) * (kDtSlots + 1)));
    FVS_CUDA(cudaHostAlloc(&ctx->h_fail, sizeof(int) * (kDtSlots + 1), cudaHostAllocMapped));
    FVS_CUDA(cudaHostGetDevicePointer(&ctx->h_fail_dev, ctx->h_fail, 0));
    std::memset(ctx->h_fail, 0, sizeof(int) * (kDtSlots + 1));
    ctx->ring_size = size_t(8) << 20;
    FVS_CUDA(cudaHostAlloc(&ctx->ring, ctx->ring_size, cudaHostAllocMapped));
    FVS_CUDA(cudaHostGetDevicePointer(&ctx->ring_dev, ctx->ring, 0));
    { // temporaries come from the stream-ordered allocator: keep its memory across syncs
      cudaMemPool_t pool;
      FVS_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
      unsigned long long keep = ~0ull;
      FVS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    fvs_ctx* raw = ctx.release();
    *out = raw;
    const int rc = fvs_set_method(raw, config->base_method, config->reconstruction, config->limiter);
    if (rc != FVS_OK) {
      fvs_destroy(raw);
      *out = nullptr;
      return rc;
    }
    return (int)FVS_OK;
  });
}

void fvs_destroy(fvs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (Level& L : ctx->levels) {
    try {
      p2p_release(ctx, L);
    } catch (...) {
    }
    free_level(L);
  }
  for (auto& ev : ctx->pending_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  for (auto& ev : ctx->free_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  if (ctx->comm) Nccl::get().CommDestroy(ctx->comm);
  if (ctx->d_slots) cudaFree(ctx->d_slots);
  if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
  if (ctx->d_fail) cudaFree(ctx->d_fail);
  if (ctx->d_barrier) cudaFree(ctx->d_barrier);
  if (ctx->d_repair) cudaFree(ctx->d_repair);
  if (ctx->d_repair_count) cudaFree(ctx->d_repair_count);
  if (ctx->h_fail) cudaFreeHost(ctx->h_fail);
  if (ctx->ring) cudaFreeHost(ctx->ring);
  for (unsigned char* r : ctx->ring_retired) cudaFreeHost(r);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int fvs_set_method(fvs_ctx* ctx, int base_method, int reconstruction, int limiter) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (base_method < 0 || base_method > 3) throw Error("unknown base method");
    if (reconstruction < 0 || reconstruction > FVS_RECON_MH2_CHARACTERISTICS) {
      throw Error("unknown reconstruction");
    }
    const bool mh2 = reconstruction >= FVS_RECON_MH2_CONSERVATIVE;
    if (!mh2 && reconstruction >= FVS_RECON_MUSCL_HANCOCK_PRIM &&
        base_method != FVS_BASE_HLLEM_LARROUTUROU) {
      // src/equations/perfect_gas/MusclHancockPrimMethod.hpp:175,
      // include/fub/equations/perfect_gas/MusclHancockCharactersticMethod.hpp:174
      throw Error("MusclHancockPrim / MusclHancockChar are defined over HLLEM (Larrouturou) only");
    }
    // MusclHancock2 takes the limiters of include/fub/flux_method/Gradient.hpp:31-59
    if (limiter < 0 || limiter > (mh2 ? FVS_LIMITER2_VANLEER : 2)) throw Error("unknown limiter");
    if (ctx->arithmetic == FVS_ARITH_TOLERANCE && base_method == FVS_BASE_GODUNOV_EXACT) {
      throw Error("the exact Riemann solver has no tolerance-arithmetic build");
    }
    if (ctx->arithmetic == FVS_ARITH_TOLERANCE && mh2) {
      throw Error("MusclHancock2 has no tolerance-arithmetic build");
    }
    const int w = reconstruction == FVS_RECON_FIRST_ORDER ? 1 : 2;
    for (Level& L : ctx->levels) {
      if (L.defined && L.scratch_gcw < L.flux_gcw + w) {
        throw Error("scratch_gcw must be >= flux_gcw + stencil width");
      }
    }
    ctx->method = fvs::MethodKey{ctx->cfg.rank, base_method, reconstruction, limiter};
    return (int)FVS_OK;
  });
}

int fvs_set_arithmetic(fvs_ctx* ctx, int arithmetic) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (arithmetic != FVS_ARITH_IEEE && arithmetic != FVS_ARITH_TOLERANCE) {
      throw Error("unknown arithmetic");
    }
    if (arithmetic == FVS_ARITH_TOLERANCE && ctx->method.base == FVS_BASE_GODUNOV_EXACT) {
      throw Error("the exact Riemann solver has no tolerance-arithmetic build");
    }
    if (arithmetic == FVS_ARITH_TOLERANCE && ctx->method.recon >= FVS_RECON_MH2_CONSERVATIVE) {
      throw Error("MusclHancock2 has no tolerance-arithmetic build");
    }
    ctx->arithmetic = arithmetic;
    return (int)FVS_OK;
  });
}

int fvs_stencil_width(const fvs_ctx* ctx) { return ctx ? ctx->stencil_width() : FVS_ERROR; }

namespace {
// builds the patches, buffers and halo plan of `L` (an empty Level) on this rank
void define_level(fvs_ctx* ctx, Level& L, const Box& domain, const double dx[3],
                  const bool periodic[3], const std::vector<Box>& boxes, const int* owner,
                  int scratch_gcw, int flux_gcw) {
  const int rank = ctx->cfg.rank;
  const int nboxes = (int)boxes.size();
  L.rank = rank;
  L.domain = domain;
  for (int d = 0; d < 3; ++d) {
    L.dx[d] = dx[d];
    L.periodic[d] = periodic[d];
    if (d >= rank && (L.domain.lo[d] != 0 || L.domain.hi[d] != 0)) {
      throw Error("unused dimensions of the domain must be [0,0]");
    }
  }
  L.scratch_gcw = scratch_gcw;
  L.flux_gcw = flux_gcw;
  L.boxes = boxes;
  L.owner.assign(nboxes, 0);
  L.local_of_global.assign(nboxes, -1);
  for (int b = 0; b < nboxes; ++b) {
    if (!L.boxes[b].ok()) throw Error("empty box");
    for (int d = rank; d < 3; ++d) {
      if (L.boxes[b].lo[d] != 0 || L.boxes[b].hi[d] != 0) {
        throw Error("unused dimensions of a box must be [0,0]");
      }
    }
    if (owner) L.owner[b] = owner[b];
    if (L.owner[b] < 0 || L.owner[b] >= ctx->cfg.comm_size) throw Error("owner out of range");
  }
  for (int d = 0; d < 3; ++d) L.max_n[d] = 1;
  for (int b = 0; b < nboxes; ++b) {
    if (L.owner[b] != ctx->cfg.comm_rank) continue;
    Patch p;
    p.global_index = b;
    p.valid = L.boxes[b];
    p.sbox = fvs::state_box(p.valid, rank, scratch_gcw);
    for (int d = 0; d < 3; ++d) {
      p.n[d] = p.sbox.len(d);
      L.max_n[d] = std::max(L.max_n[d], p.n[d]);
    }
    p.cells = p.sbox.size();
    for (int d = 0; d < rank; ++d) {
      p.fbox[d] = fvs::flux_box(p.valid, rank, d, scratch_gcw, flux_gcw);
      for (int w = 1; w <= 2; ++w) {
        const std::vector<unsigned char> mask =
            fvs::stable_dt_face_mask(p.valid, scratch_gcw, w, d);
        FVS_CUDA(cudaMalloc(&p.dtmask[w - 1][d], mask.size()));
        FVS_CUDA(upload_sync(p.dtmask[w - 1][d], mask.data(), mask.size()));
      }
    }
    L.local_of_global[b] = (int)L.patches.size();
    L.patches.push_back(p);
  }
  L.defined = true;
  L.cur = 0;
  L.keep = -1;
  ensure_buffer(ctx, L, 0);
  // halo plan
  const std::vector<fvs::HaloCopy> plan =
      fvs::halo_plan(rank, L.domain, L.periodic, L.boxes, scratch_gcw);
  const int me = ctx->cfg.comm_rank;
  std::map<int, PeerExchange> peers;
  for (const fvs::HaloCopy& c : plan) {
    const int so = L.owner[c.src_box], dow = L.owner[c.dst_box];
    if (so == me && dow == me) {
      L.local_copies.push_back(c);
      L.max_copy_cells = std::max(L.max_copy_cells, c.region.size());
    }
    if (so != dow) L.halo_crosses_ranks = true; // the same on every rank (global plan)
    if (so == me && dow == me) {
    } else if (so == me) {
      PeerExchange& pe = peers[dow];
      pe.peer = dow;
      pe.sends.push_back(c);
      pe.send_elems += c.region.size() * ctx->ncomp();
    } else if (dow == me) {
      PeerExchange& pe = peers[so];
      pe.peer = so;
      pe.recvs.push_back(c);
      pe.recv_elems += c.region.size() * ctx->ncomp();
    }
  }
  for (auto& kv : peers) {
    PeerExchange& pe = kv.second;
    if (pe.send_elems) FVS_CUDA(cudaMalloc(&pe.sendbuf, pe.send_elems * sizeof(double)));
    if (pe.recv_elems) FVS_CUDA(cudaMalloc(&pe.recvbuf, pe.recv_elems * sizeof(double)));
    L.peers.push_back(std::move(pe));
  }
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
}
} // namespace

int fvs_level_define(fvs_ctx* ctx, int level, const int domain_lo[3], const int domain_hi[3],
                     const double dx[3], const int periodic[3], int nboxes,
                     const int (*box_lo)[3], const int (*box_hi)[3], const int* owner,
                     int scratch_gcw, int flux_gcw) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (level < 0 || level > 64) throw Error("level out of range");
    if (nboxes < 0) throw Error("negative box count");
    if (scratch_gcw < 0 || flux_gcw < 0) throw Error("negative ghost width");
    if (scratch_gcw < flux_gcw + ctx->stencil_width()) {
      throw Error("scratch_gcw must be >= flux_gcw + stencil width");
    }
    ctx->use_device();
    if ((int)ctx->levels.size() <= level) ctx->levels.resize(level + 1);
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    p2p_release(ctx, ctx->levels[level]);
    free_level(ctx->levels[level]);
    std::vector<Box> boxes(nboxes);
    for (int b = 0; b < nboxes; ++b) boxes[b] = fvs::make_box(box_lo[b], box_hi[b]);
    bool per[3];
    for (int d = 0; d < 3; ++d) per[d] = periodic && periodic[d] != 0;
    define_level(ctx, ctx->levels[level], fvs::make_box(domain_lo, domain_hi), dx, per, boxes, owner,
                 scratch_gcw, flux_gcw);
    if (level > 0) ctx->levels[level - 1].plane_cache.clear();
    if (level + 1 < (int)ctx->levels.size()) ctx->levels[level + 1].cf_plans.clear();
    return (int)FVS_OK;
  });
}

// GriddingAlgorithm::RemakeLevel / MakeNewLevelFromCoarse followed by
// IntegratorContext::CopyDataToScratch (src/AMReX/GriddingAlgorithm.cpp:217-260,
// src/AMReX/IntegratorContext.cpp:629-638): the level gets new boxes; every cell of
// the new scratch arrays inside the domain takes the old level's valid data where
// an old box holds it and the coarse parent (pc_interp) elsewhere, then the
// level's own ghost fill (fine-fine copies, periodic images, physical boundaries).
int fvs_regrid_level(fvs_ctx* ctx, int level, int nboxes, const int (*box_lo)[3],
                     const int (*box_hi)[3], const int* owner) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (level < 1) throw Error("only levels above the base level are regridded");
    if (nboxes < 0) throw Error("negative box count");
    ctx->use_device();
    if ((int)ctx->levels.size() <= level) ctx->levels.resize(level + 1);
    Level& coarse = ctx->level(level - 1);
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    coarse.plane_cache.clear(); // its CrseInit planes follow the boxes of this level (stream idle)
    if (nboxes == 0) { // GriddingAlgorithm::ClearLevel: the level and everything finer goes
      for (int l = level; l < (int)ctx->levels.size(); ++l) {
        p2p_release(ctx, ctx->levels[l]);
        free_level(ctx->levels[l]);
      }
      return (int)FVS_OK;
    }
    int ratio[3];
    effective_ratio(ctx, ratio);
    const bool had_old = ctx->levels[level].defined;
    Level old = std::move(ctx->levels[level]);
    ctx->levels[level] = Level();
    Level& L = ctx->levels[level];
    std::vector<Box> boxes(nboxes);
    for (int b = 0; b < nboxes; ++b) boxes[b] = fvs::make_box(box_lo[b], box_hi[b]);
    try {
      if (had_old) {
        define_level(ctx, L, old.domain, old.dx, old.periodic, boxes, owner, old.scratch_gcw,
                     old.flux_gcw);
        L.bcs = old.bcs;
      } else {
        double dx[3];
        for (int d = 0; d < 3; ++d) dx[d] = coarse.dx[d] / ratio[d];
        define_level(ctx, L, refine_box(coarse.domain, ratio), dx, coarse.periodic, boxes, owner,
                     coarse.scratch_gcw, coarse.flux_gcw);
        L.bcs = coarse.bcs;
      }
      pc_interp(ctx, L, coarse, /*skip_valid=*/false);
      if (had_old) copy_from_old_level(ctx, L, old);
      fill_ghost_layer(ctx, L);
      // plans are freed only while the stream is idle: queued kernels read their
      // descriptor arrays, and cudaFree does not wait for a non-blocking stream
      FVS_CUDA(cudaStreamSynchronize(ctx->stream));
      L.cf_plans.clear(); // the fill above was a one-off
      if (level + 1 < (int)ctx->levels.size()) ctx->levels[level + 1].cf_plans.clear();
    } catch (...) {
      free_level(old);
      throw;
    }
    p2p_release(ctx, old);
    free_level(old);
    // the register of a finer level refers to this level's boxes only through the
    // coarse index space: it stays valid; this level's own register is rebuilt by
    // the next fvs_reset_coarse_fine_fluxes
    return (int)FVS_OK;
  });
}

int fvs_level_add_boundary(fvs_ctx* ctx, int level, int dir, int side, int kind) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    if (side < 0 || side > 1) throw Error("side must be 0 or 1");
    if (kind != FVS_BC_TRANSMISSIVE && kind != FVS_BC_REFLECTIVE) throw Error("unknown boundary kind");
    L.bcs.push_back({dir, side, kind});
    return (int)FVS_OK;
  });
}

int fvs_state_box(const fvs_ctx* ctx, int level, int box, int lo[3], int hi[3], int* ncomp) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    const Level& L = const_cast<fvs_ctx*>(ctx)->level(level);
    if (box < 0 || box >= (int)L.boxes.size()) throw Error("box index out of range");
    const Box b = fvs::state_box(L.boxes[box], ctx->cfg.rank, L.scratch_gcw);
    for (int d = 0; d < 3; ++d) {
      lo[d] = b.lo[d];
      hi[d] = b.hi[d];
    }
    if (ncomp) *ncomp = ctx->ncomp();
    return (int)FVS_OK;
  });
}

int fvs_flux_box(const fvs_ctx* ctx, int level, int box, int dir, int lo[3], int hi[3],
                 int* ncomp) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    const Level& L = const_cast<fvs_ctx*>(ctx)->level(level);
    if (box < 0 || box >= (int)L.boxes.size()) throw Error("box index out of range");
    check_dir(ctx, dir);
    const Box b = fvs::flux_box(L.boxes[box], ctx->cfg.rank, dir, L.scratch_gcw, L.flux_gcw);
    for (int d = 0; d < 3; ++d) {
      lo[d] = b.lo[d];
      hi[d] = b.hi[d];
    }
    if (ncomp) *ncomp = ctx->ncons();
    return (int)FVS_OK;
  });
}

int fvs_upload_state(fvs_ctx* ctx, int level, int box, const double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    drop_hints(ctx, L);
    Patch& p = ctx->patch(L, box);
    FVS_CUDA(cudaMemcpyAsync(p.buf[L.cur], host_fab, (size_t)p.cells * ctx->ncomp() * sizeof(double),
                             cudaMemcpyHostToDevice, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_download_state(fvs_ctx* ctx, int level, int box, double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    Patch& p = ctx->patch(L, box);
    FVS_CUDA(cudaMemcpyAsync(host_fab, p.buf[L.cur], (size_t)p.cells * ctx->ncomp() * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_download_state_region(fvs_ctx* ctx, int level, int box, const int lo[3], const int hi[3],
                              double* host) {
  return guarded([&] {
    if (!ctx || !lo || !hi || !host) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    Patch& p = ctx->patch(L, box);
    int off[3], len[3];
    for (int d = 0; d < 3; ++d) {
      off[d] = lo[d] - p.sbox.lo[d];
      len[d] = hi[d] - lo[d] + 1;
      if (off[d] < 0 || len[d] <= 0 || off[d] + len[d] > p.n[d]) {
        throw Error("region is not inside the state array of the box");
      }
    }
    const size_t region = (size_t)len[0] * len[1] * len[2];
    for (int c = 0; c < ctx->ncomp(); ++c) {
      cudaMemcpy3DParms m{};
      m.srcPtr = make_cudaPitchedPtr(p.buf[L.cur] + (size_t)c * p.cells, sizeof(double) * p.n[0],
                                     (size_t)p.n[0], (size_t)p.n[1]);
      m.srcPos = make_cudaPos(sizeof(double) * off[0], (size_t)off[1], (size_t)off[2]);
      m.dstPtr = make_cudaPitchedPtr(host + (size_t)c * region, sizeof(double) * len[0],
                                     (size_t)len[0], (size_t)len[1]);
      m.dstPos = make_cudaPos(0, 0, 0);
      m.extent = make_cudaExtent(sizeof(double) * len[0], (size_t)len[1], (size_t)len[2]);
      m.kind = cudaMemcpyDeviceToHost;
      FVS_CUDA(cudaMemcpy3DAsync(&m, ctx->stream));
    }
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_upload_state_async(fvs_ctx* ctx, int level, int box, const double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    drop_hints(ctx, L);
    Patch& p = ctx->patch(L, box);
    FVS_CUDA(cudaMemcpyAsync(p.buf[L.cur], host_fab, (size_t)p.cells * ctx->ncomp() * sizeof(double),
                             cudaMemcpyHostToDevice, ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_download_state_async(fvs_ctx* ctx, int level, int box, double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    Patch& p = ctx->patch(L, box);
    FVS_CUDA(cudaMemcpyAsync(host_fab, p.buf[L.cur], (size_t)p.cells * ctx->ncomp() * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    return (int)FVS_OK;
  });
}

namespace {
// valid cells of a patch <-> a dense host array, one 3-D copy per component
void copy_valid(fvs_ctx* ctx, Level& L, Patch& p, double* host, bool to_device) {
  const int g[3] = {p.valid.lo[0] - p.sbox.lo[0], p.valid.lo[1] - p.sbox.lo[1],
                    p.valid.lo[2] - p.sbox.lo[2]};
  const size_t vx = (size_t)p.valid.len(0), vy = (size_t)p.valid.len(1), vz = (size_t)p.valid.len(2);
  for (int c = 0; c < ctx->ncomp(); ++c) {
    cudaMemcpy3DParms parms{};
    double* dev = p.buf[L.cur] + (size_t)p.cells * c;
    double* hst = host + vx * vy * vz * c;
    const cudaPitchedPtr dptr =
        make_cudaPitchedPtr(dev, (size_t)p.n[0] * sizeof(double), (size_t)p.n[0], (size_t)p.n[1]);
    const cudaPitchedPtr hptr = make_cudaPitchedPtr(hst, vx * sizeof(double), vx, vy);
    const cudaPos dpos = make_cudaPos((size_t)g[0] * sizeof(double), (size_t)g[1], (size_t)g[2]);
    const cudaPos hpos = make_cudaPos(0, 0, 0);
    parms.srcPtr = to_device ? hptr : dptr;
    parms.srcPos = to_device ? hpos : dpos;
    parms.dstPtr = to_device ? dptr : hptr;
    parms.dstPos = to_device ? dpos : hpos;
    parms.extent = make_cudaExtent(vx * sizeof(double), vy, vz);
    parms.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    FVS_CUDA(cudaMemcpy3DAsync(&parms, ctx->stream));
  }
}
} // namespace

int fvs_copy_data_to_scratch_async(fvs_ctx* ctx, int level, int box, const double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    drop_hints(ctx, L);
    copy_valid(ctx, L, ctx->patch(L, box), const_cast<double*>(host_fab), true);
    return (int)FVS_OK;
  });
}

int fvs_copy_scratch_to_data_async(fvs_ctx* ctx, int level, int box, double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    copy_valid(ctx, L, ctx->patch(L, box), host_fab, false);
    return (int)FVS_OK;
  });
}

int fvs_synchronize(fvs_ctx* ctx) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_upload_fluxes(fvs_ctx* ctx, int level, int box, int dir, const double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    ensure_fluxes(ctx, L, dir);
    Patch& p = ctx->patch(L, box);
    FVS_CUDA(cudaMemcpyAsync(p.flux[dir], host_fab,
                             (size_t)p.fbox[dir].size() * ctx->ncons() * sizeof(double),
                             cudaMemcpyHostToDevice, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_download_fluxes(fvs_ctx* ctx, int level, int box, int dir, double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    ensure_fluxes(ctx, L, dir);
    Patch& p = ctx->patch(L, box);
    FVS_CUDA(cudaMemcpyAsync(host_fab, p.flux[dir],
                             (size_t)p.fbox[dir].size() * ctx->ncons() * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_state_device_ptr(fvs_ctx* ctx, int level, int box, double** dptr) {
  return guarded([&] {
    if (!ctx || !dptr) throw Error("null argument");
    Level& L = ctx->level(level);
    *dptr = ctx->patch(L, box).buf[L.cur];
    return (int)FVS_OK;
  });
}

int fvs_compute_stable_dt(fvs_ctx* ctx, int level, int dir, double* dt_out) {
  return guarded([&] {
    if (!ctx || !dt_out) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    if (L.has_eb) {
      const int slot = take_slot(ctx);
      cc_stable_dt(ctx, L, dir, slot);
      double sig;
      fetch_slots(ctx, slot, 1, &sig);
      *dt_out = dt_from_signal(L.dx[dir], sig);
      return (int)FVS_OK;
    }
    const BoxDesc* boxes = box_descs(ctx, L, dir, L.cur, L.cur, false);
    const int slot = take_slot(ctx);
    fvs::SweepArgs A = sweep_args(ctx, L, dir, 0.0, boxes, slot);
    if (!L.patches.empty()) {
      const int rc = fvs::launch_stable_dt(ctx->method, dir, A, ctx->stream);
      check_launch(rc);
      ctx->launches += rc;
    }
    double sig;
    fetch_slots(ctx, slot, 1, &sig);
    *dt_out = dt_from_signal(L.dx[dir], sig);
    return (int)FVS_OK;
  });
}

int fvs_compute_numeric_fluxes(fvs_ctx* ctx, int level, double dt, int dir) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    if (L.has_eb) {
      cc_numeric_fluxes(ctx, L, dt, dir);
      return (int)FVS_OK;
    }
    ensure_fluxes(ctx, L, dir);
    const BoxDesc* boxes = box_descs(ctx, L, dir, L.cur, L.cur, true);
    fvs::SweepArgs A = sweep_args(ctx, L, dir, dt, boxes, -1);
    if (!L.patches.empty()) {
      const int rc = fvs::launch_numeric_fluxes(ctx->method, dir, A, ctx->stream);
      check_launch(rc);
      ctx->launches += rc;
    }
    return (int)FVS_OK;
  });
}

int fvs_update_conservatively(fvs_ctx* ctx, int level, double dt, int dir) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    if (L.has_eb) {
      cc_update(ctx, L, dt, dir);
      return (int)FVS_OK;
    }
    ensure_fluxes(ctx, L, dir);
    const BoxDesc* boxes = box_descs(ctx, L, dir, L.cur, L.cur, true);
    fvs::SweepArgs A = sweep_args(ctx, L, dir, dt, boxes, -1);
    if (!L.patches.empty()) {
      const int rc = fvs::launch_update_conservatively(ctx->cfg.rank, dir, A, ctx->stream);
      check_launch(rc);
      ctx->launches += rc;
    }
    return (int)FVS_OK;
  });
}

int fvs_complete_from_cons(fvs_ctx* ctx, int level) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    Level& L = ctx->level(level);
    if (L.has_eb) {
      cc_complete(ctx, L);
      return (int)FVS_OK;
    }
    const BoxDesc* boxes = box_descs(ctx, L, 0, L.cur, L.cur, false);
    fvs::SweepArgs A = sweep_args(ctx, L, 0, 0.0, boxes, -1);
    if (!L.patches.empty()) {
      const int rc = fvs::launch_complete_from_cons(ctx->cfg.rank, A, ctx->stream);
      check_launch(rc);
      ctx->launches += rc;
    }
    return (int)FVS_OK;
  });
}

namespace {
void plane_fluxes(fvs_ctx* ctx, Level& L, int level, int dir, double dt, int src);
void accumulate_coarse_fine(fvs_ctx* ctx, int level, double scale, int dir);
// does a sweep on `level` feed a flux register (as the fine or as the coarse side)?
bool feeds_register(fvs_ctx* ctx, int level) {
  const Level& L = ctx->level(level);
  if (level > 0 && L.cf_coarse >= 0) return true;
  return level + 1 < (int)ctx->levels.size() && ctx->levels[(size_t)level + 1].defined &&
         ctx->levels[(size_t)level + 1].cf_coarse == level;
}
} // namespace

int fvs_sweep(fvs_ctx* ctx, int level, double dt, int dir, double* level_dt_out) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    Level& L = ctx->level(level);
    check_dir(ctx, dir);
    // DimensionalSplitLevelIntegrator.hpp:262-264: AccumulateCoarseFineFluxes follows the
    // sweep; the fused kernel keeps no flux array, so the register planes are evaluated
    // from the sweep's input buffer
    const bool feeds = feeds_register(ctx, level);
    if (L.has_eb) {
      // cut-cell levels: the unfused sequence, in place; the dt check comes
      // first so that a rejected sweep leaves the state untouched
      const int slot = take_slot(ctx);
      cc_stable_dt(ctx, L, dir, slot);
      double sig;
      fetch_slots(ctx, slot, 1, &sig);
      const double level_dt = dt_from_signal(L.dx[dir], sig);
      if (level_dt_out) *level_dt_out = level_dt;
      if (level_dt < dt) return (int)FVS_TIME_STEP_TOO_LARGE;
      cc_numeric_fluxes(ctx, L, dt, dir);
      cc_update(ctx, L, dt, dir);
      cc_complete(ctx, L);
      if (feeds) accumulate_coarse_fine(ctx, level, dt, dir);
      return (int)FVS_OK;
    }
    const int src = L.cur;
    int dst;
    double hint = dt_hint(ctx, L, dir);
    int slot = launch_sweep(ctx, L, dt, dir, &dst, hint);
    double sig;
    fetch_slots(ctx, slot, 1, &sig);
    if (!verify_hint(hint, sig)) { // the source buffer is untouched: once more, eagerly
      ctx->hint_redos += 1;
      slot = launch_sweep(ctx, L, dt, dir, &dst, 0.0);
      fetch_slots(ctx, slot, 1, &sig);
    }
    L.last_sig[dir] = sig;
    const double level_dt = dt_from_signal(L.dx[dir], sig);
    if (level_dt_out) *level_dt_out = level_dt;
    // DimensionalSplitLevelIntegrator.hpp:243-245
    if (level_dt < dt) return (int)FVS_TIME_STEP_TOO_LARGE;
    L.cur = dst;
    if (feeds) {
      plane_fluxes(ctx, L, level, dir, dt, src);
      accumulate_coarse_fine(ctx, level, dt, dir);
    }
    return (int)FVS_OK;
  });
}

int fvs_fill_ghost_layer(fvs_ctx* ctx, int level) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    fill_ghost_layer(ctx, ctx->level(level));
    return (int)FVS_OK;
  });
}

int fvs_apply_boundary_condition(fvs_ctx* ctx, int level, int dir) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    check_dir(ctx, dir);
    apply_boundary_condition(ctx, ctx->level(level), dir);
    return (int)FVS_OK;
  });
}

namespace {
// DimensionalSplitLevelIntegrator::ComputeStableDt of one level, enqueued: the max
// signal speeds of the `rank` directions go to three consecutive accumulator slots.
struct DtAllLaunch {
  int slot;
  double hints[3];
};
DtAllLaunch launch_dt_all(fvs_ctx* ctx, Level& L, bool use_hints) {
  const int rank = ctx->cfg.rank;
  DtAllLaunch out{};
  out.slot = take_slot(ctx, 3);
  for (int d = 0; d < 3; ++d) out.hints[d] = 0.0;
  if (L.has_eb) {
    for (int d = 0; d < rank; ++d) cc_stable_dt(ctx, L, d, out.slot + d);
    return out;
  }
  fvs::DtAllArgs A{};
  A.eq = ctx->eq;
  for (int d = 0; d < 3; ++d) {
    A.boxes[d] = box_descs(ctx, L, d < rank ? d : 0, L.cur, L.cur, false);
    A.max_n[d] = L.max_n[d];
  }
  A.nboxes = (int)L.patches.size();
  A.max_signal = ctx->d_slots + out.slot;
  A.fail_flag = ctx->d_fail + out.slot;
  for (int d = 0; d < 3; ++d) {
    out.hints[d] = use_hints && d < rank ? dt_hint(ctx, L, d) : 0.0;
    A.sig_hint[d] = out.hints[d];
  }
  if (!L.patches.empty()) {
    const int rc = ctx->arithmetic == FVS_ARITH_TOLERANCE
                       ? fvs::launch_stable_dt_all_tolerance(ctx->method, A, ctx->stream)
                       : fvs::launch_stable_dt_all(ctx->method, A, ctx->stream);
    check_launch(rc);
    ctx->launches += rc;
  }
  return out;
}
// true if the hints held; then the level's dt (DimensionalSplitLevelIntegrator.hpp:204-219)
bool finish_dt_all(fvs_ctx* ctx, Level& L, const DtAllLaunch& launch, const double sig[3],
                   double* dt_out) {
  const int rank = ctx->cfg.rank;
  for (int d = 0; d < rank; ++d)
    if (!verify_hint(launch.hints[d], sig[d])) return false;
  double min_dt = std::numeric_limits<double>::max();
  for (int d = 0; d < rank; ++d) {
    if (!L.has_eb) L.last_sig[d] = sig[d];
    min_dt = std::min(min_dt, dt_from_signal(L.dx[d], sig[d]));
  }
  *dt_out = min_dt;
  return true;
}
} // namespace

int fvs_compute_stable_dt_all(fvs_ctx* ctx, int level, double* dt_out) {
  return guarded([&] {
    if (!ctx || !dt_out) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    for (int attempt = 0;; ++attempt) {
      const DtAllLaunch launch = launch_dt_all(ctx, L, attempt == 0);
      double sig[3];
      fetch_slots(ctx, launch.slot, 3, sig);
      if (finish_dt_all(ctx, L, launch, sig, dt_out)) break;
      ctx->hint_redos += 1;
      if (attempt > 0) throw Error("stable dt evaluation without hints failed its own check");
    }
    return (int)FVS_OK;
  });
}

namespace {
int num_levels(const fvs_ctx* ctx) {
  int n = 0;
  while (n < (int)ctx->levels.size() && ctx->levels[n].defined) ++n;
  return n;
}
int max_ratio(const fvs_ctx* ctx) {
  int m = 1;
  for (int d = 0; d < ctx->cfg.rank; ++d) m = std::max(m, ctx->ratio[d]);
  return m;
}
int total_ratio(const fvs_ctx* ctx, int level) {
  int r = 1;
  for (int l = 1; l <= level; ++l) r *= max_ratio(ctx);
  return r;
}
void call(int rc) {
  if (rc == FVS_ERROR) throw Error(g_error);
}
struct TooLarge {
  double dt;
};

// SubcycleFineFirstSolver::AdvanceLevel (include/fub/solver/SubcycleFineFirstSolver.hpp:114-175)
// for a hierarchy: finer levels first, `ratio` subcycles each, flux registers, flux
// correction, conservative coarsening -- speculative like the single-level step: every
// sweep also evaluates the stable dt of its input state into an accumulator slot, the
// slots are read back with ONE synchronisation per coarse step (or when they run out)
// and checked in execution order, so the outcome (TimeStepTooLarge with the dt of the
// first failing check, the list of dts up to it, untouched state) is the reference's
// check-before-sweep.  Regular levels run the fused sweep; the fluxes the registers need
// -- the boundary planes of the level's own boxes (FineAdd) and of the finer level's
// boxes (CrseInit) -- are evaluated from the sweep's input buffer on those planes only.
// Cut-cell levels run the cut-cell policies in place.
struct PendingSweep {
  int slot, level, dir;
  double split_dt, hint;
};
struct HintRedo {};
struct AmrStep {
  std::vector<PendingSweep> pending; // execution order; slots are consecutive
  std::vector<double> dts;           // checked per-sweep dts
  bool use_hints{true};
};

// reads the pending accumulators back and applies the reference's checks in order
void flush_pending(fvs_ctx* ctx, AmrStep& st) {
  if (st.pending.empty()) return;
  const int count = (int)st.pending.size();
  std::vector<double> sig((size_t)count);
  std::vector<int> fails((size_t)count);
  fetch_slots(ctx, st.pending.front().slot, count, sig.data(), fails.data());
  std::vector<PendingSweep> pending;
  pending.swap(st.pending);
  for (int i = 0; i < count; ++i) {
    const PendingSweep& ps = pending[(size_t)i];
    Level& L = ctx->levels[(size_t)ps.level];
    if (!verify_hint(ps.hint, sig[(size_t)i])) throw HintRedo{};
    L.last_sig[ps.dir] = sig[(size_t)i];
    const double level_dt = dt_from_signal(L.dx[ps.dir], sig[(size_t)i]);
    st.dts.push_back(level_dt);
    if (level_dt < ps.split_dt) throw TooLarge{level_dt * total_ratio(ctx, ps.level)};
    if (fails[(size_t)i]) throw_newton();
  }
}

// descriptors of the register planes of direction dir of level `level`, reading `src`
const BoxDesc* plane_descs(fvs_ctx* ctx, Level& L, int level, int dir, int src, int* count) {
  auto key = std::make_tuple(dir, src);
  auto it = L.plane_cache.find(key);
  if (it != L.plane_cache.end()) {
    *count = it->second.second;
    return it->second.first->ptr;
  }
  const int w = ctx->stencil_width();
  int ratio[3];
  effective_ratio(ctx, ratio);
  const bool feeds_coarser = level > 0 && L.cf_coarse >= 0;
  Level* finer = (level + 1 < (int)ctx->levels.size() && ctx->levels[(size_t)level + 1].defined &&
                  ctx->levels[(size_t)level + 1].cf_coarse == level)
                     ? &ctx->levels[(size_t)level + 1]
                     : nullptr;
  const std::vector<std::array<int, 3>> shifts = period_shifts(L, ctx->cfg.rank);
  std::vector<BoxDesc> host;
  for (Patch& p : L.patches) {
    if (!in_group(L, p, kGroupRegular)) continue; // cut fabs: cc_numeric_fluxes fills the array
    std::set<int> planes; // face index = index of the cell to its right
    if (feeds_coarser) {
      planes.insert(p.valid.lo[dir]);
      planes.insert(p.valid.hi[dir] + 1);
    }
    if (finer) {
      for (const Box& g : finer->boxes) {
        const Box cg = coarsen_box(g, ratio);
        for (const std::array<int, 3>& sh : shifts) {
          // the coarse box's image holds the register faces <=> in the box's own index
          // space the plane sits at P - shift and the fine box at cg - shift
          Box local = cg;
          for (int e = 0; e < 3; ++e) {
            local.lo[e] -= sh[(size_t)e];
            local.hi[e] -= sh[(size_t)e];
          }
          Box nodes = p.valid;
          nodes.hi[dir] += 1;
          for (int side = 0; side < 2; ++side) {
            Box plane = local;
            plane.lo[dir] = plane.hi[dir] = side == 0 ? local.lo[dir] : local.hi[dir] + 1;
            if (fvs::intersect(plane, nodes).ok()) planes.insert(plane.lo[dir]);
          }
        }
      }
    }
    for (int P : planes) {
      BoxDesc d{};
      d.src = p.buf[src];
      d.dst = p.buf[src];
      d.flux = p.flux[dir];
      for (int e = 0; e < 3; ++e) {
        d.n[e] = p.n[e];
        d.fn[e] = p.fbox[dir].len(e);
        d.foff[e] = p.fbox[dir].lo[e] - p.sbox.lo[e];
      }
      d.sc = p.cells;
      d.fsc = p.fbox[dir].size();
      int ulo, uhi, flo, fhi;
      fvs::update_cell_range(p.valid, dir, L.flux_gcw, ulo, uhi);
      fvs::flux_face_range(p.valid, dir, L.scratch_gcw, L.flux_gcw, w, flo, fhi);
      if (P < flo || P > fhi) continue;
      d.ulo = ulo - p.sbox.lo[dir];
      d.uhi = uhi - p.sbox.lo[dir];
      d.flo = d.fhi = P - p.sbox.lo[dir];
      d.dtmask = p.dtmask[w - 1][dir];
      d.fab_type = FVS_FAB_REGULAR;
      host.push_back(d);
    }
  }
  auto arr = std::make_unique<DeviceArray<BoxDesc>>();
  if (!host.empty()) arr->upload(host);
  const BoxDesc* ptr = host.empty() ? nullptr : arr->ptr;
  *count = (int)host.size();
  L.plane_cache[key] = std::make_pair(std::move(arr), *count);
  return ptr;
}

// numeric fluxes of the sweep (dir, dt) on the register planes, from buffer `src`
void plane_fluxes(fvs_ctx* ctx, Level& L, int level, int dir, double dt, int src) {
  ensure_fluxes(ctx, L, dir);
  int count = 0;
  const BoxDesc* descs = plane_descs(ctx, L, level, dir, src, &count);
  if (count == 0) return;
  fvs::SweepArgs A = sweep_args(ctx, L, dir, dt, descs, -1);
  A.nboxes = count;
  A.max_n[dir] = 1; // one plane of faces per descriptor
  const int rc = fvs::launch_numeric_fluxes(ctx->method, dir, A, ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
}

void accumulate_coarse_fine(fvs_ctx* ctx, int level, double scale, int dir) {
  Level& L = ctx->level(level);
  ensure_fluxes(ctx, L, dir);
  const double vol = cell_volume(ctx, L);
  const double vol_dt_over_dx = vol * scale / L.dx[dir];
  if (level > 0 && L.cf_coarse >= 0) fine_add(ctx, L, dir, vol_dt_over_dx);
  if (level + 1 < (int)ctx->levels.size() && ctx->levels[(size_t)level + 1].defined &&
      ctx->levels[(size_t)level + 1].cf_coarse == level) {
    crse_init(ctx, ctx->levels[(size_t)level + 1], L, dir, -vol_dt_over_dx);
  }
}

void advance_level(fvs_ctx* ctx, int level, double dt, int sub_first, int split_kind, AmrStep& st) {
  Level& L = ctx->level(level);
  const int rank = ctx->cfg.rank;
  cc_pre_advance(ctx, L); // PreAdvanceLevel
  const int next = level + 1;
  const bool has_next = next < num_levels(ctx);
  if (has_next) {
    call(fvs_reset_coarse_fine_fluxes(ctx, next, level));
    const int r = max_ratio(ctx);
    for (int i = 0; i < r; ++i) advance_level(ctx, next, dt / r, i, split_kind, st);
  }
  const bool feeds = has_next || (level > 0 && L.cf_coarse >= 0);
  // AdvanceLevelNonRecursively; MakeSplitDirections uses the cycles of level 0
  const bool is_odd = ((ctx->cycles % 2) + sub_first) % 2 != 0;
  int dirs[3];
  for (int i = 0; i < rank; ++i) dirs[i] = is_odd ? (rank - 1 - i) : i;
  const std::vector<SweepPlanEntry> plan = split_plan(rank, split_kind, dt);
  int counts[3] = {0, 0, 0};
  for (const SweepPlanEntry& e : plan) {
    const int dir = dirs[e.slot];
    if (counts[e.slot] > 0) apply_boundary_condition(ctx, L, dir);
    counts[e.slot] += 1;
    if (ctx->next_slot + 1 > kDtSlots) { // out of accumulators: check what is pending
      flush_pending(ctx, st);
      ctx->next_slot = 0;
    }
    if (L.has_eb) {
      const int src = L.cur;
      int dst;
      const int slot = eb_sweep(ctx, L, e.dt, dir, &dst);
      st.pending.push_back({slot, level, dir, e.dt, 0.0});
      // regular fabs: register-plane fluxes from the input buffer; cut fabs: the
      // cut-cell kernels have written the whole (unshielded) flux array
      if (feeds) plane_fluxes(ctx, L, level, dir, e.dt, src);
      L.cur = dst;
    } else {
      const double hint = st.use_hints ? dt_hint(ctx, L, dir) : 0.0;
      const int src = L.cur;
      int dst;
      const int slot = launch_sweep(ctx, L, e.dt, dir, &dst, hint, /*force_ieee=*/true);
      st.pending.push_back({slot, level, dir, e.dt, hint});
      if (feeds) plane_fluxes(ctx, L, level, dir, e.dt, src);
      L.cur = dst;
    }
    if (feeds) accumulate_coarse_fine(ctx, level, e.dt, dir);
  }
  if (has_next) {
    call(fvs_apply_flux_correction(ctx, next, level, dt));
    call(fvs_coarsen_conservatively(ctx, next, level));
    call(fvs_complete_from_cons(ctx, level));
  }
}

int advance_hierarchy_multilevel(fvs_ctx* ctx, double dt, int split_kind, double* failure_dt,
                                 double* sweep_dts, int* n_sweep_dts) {
  const int nl = num_levels(ctx);
  std::vector<int> entry(nl);
  for (int l = 0; l < nl; ++l) entry[l] = ctx->levels[l].cur;
  auto roll_back = [&] {
    for (int l = 0; l < nl; ++l) {
      ctx->levels[l].cur = entry[l];
      ctx->levels[l].keep = -1;
    }
  };
  AmrStep st;
  int status = FVS_OK;
  for (int attempt = 0;; ++attempt) {
    st = AmrStep();
    st.use_hints = attempt == 0;
    ctx->next_slot = 0;
    for (int l = 0; l < nl; ++l) {
      Level& L = ctx->levels[l];
      L.cur = entry[l];
      L.keep = entry[l]; // the sweeps ping-pong between the other two buffers
    }
    try {
      advance_level(ctx, 0, dt, 0, split_kind, st);
      flush_pending(ctx, st);
    } catch (const TooLarge& e) {
      if (failure_dt) *failure_dt = e.dt;
      status = FVS_TIME_STEP_TOO_LARGE;
    } catch (const HintRedo&) {
      // a stale lazy-dt hint: the sweeps behind it are void, repeat without hints
      FVS_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->hint_redos += 1;
      roll_back();
      continue;
    } catch (...) {
      cudaStreamSynchronize(ctx->stream);
      roll_back();
      throw;
    }
    break;
  }
  const std::vector<double>& dts = st.dts;
  const int cap = (sweep_dts && n_sweep_dts) ? *n_sweep_dts : 0;
  for (int i = 0; i < (int)dts.size() && i < cap; ++i) sweep_dts[i] = dts[(size_t)i];
  if (n_sweep_dts) *n_sweep_dts = (int)dts.size();
  if (status != FVS_OK) {
    FVS_CUDA(cudaStreamSynchronize(ctx->stream)); // work queued behind the failing check
    roll_back(); // every level: the entry buffers were never written
    return status;
  }
  for (int l = 0; l < nl; ++l) ctx->levels[l].keep = -1;
  ctx->cycles += 1;
  ctx->time_point = ctx->time_point + dt;
  if (ctx->cycles % ctx->regrid_frequency == 0) { // SubcycleFineFirstSolver.hpp:177-192
    fill_ghost_layer(ctx, ctx->levels[0]);
    for (int l = 1; l < nl; ++l) call(fvs_fill_ghost_layer_two_levels(ctx, l, l - 1));
  }
  return (int)FVS_OK;
}
} // namespace

int fvs_compute_stable_dt_hierarchy(fvs_ctx* ctx, double* dt_out) {
  return guarded([&] {
    if (!ctx || !dt_out) throw Error("null argument");
    ctx->use_device();
    // SubcycleFineFirstSolver::ComputeStableDt (SubcycleFineFirstSolver.hpp:100-112); the
    // passes of all levels are enqueued back to back and read with one synchronisation
    const int nl = num_levels(ctx);
    if (3 * nl > kDtSlots) throw Error("too many levels");
    ctx->next_slot = 0;
    std::vector<DtAllLaunch> launches((size_t)nl);
    for (int l = 0; l < nl; ++l) launches[(size_t)l] = launch_dt_all(ctx, ctx->levels[(size_t)l], true);
    std::vector<double> sig((size_t)(3 * nl));
    fetch_slots(ctx, 0, 3 * nl, sig.data());
    double min_dt = std::numeric_limits<double>::max();
    for (int l = 0; l < nl; ++l) {
      Level& L = ctx->levels[(size_t)l];
      double d = 0.0;
      if (!finish_dt_all(ctx, L, launches[(size_t)l], sig.data() + 3 * l, &d)) {
        ctx->hint_redos += 1; // a stale hint on this level: once more without
        const DtAllLaunch again = launch_dt_all(ctx, L, false);
        double s3[3];
        fetch_slots(ctx, again.slot, 3, s3);
        if (!finish_dt_all(ctx, L, again, s3, &d)) {
          throw Error("stable dt evaluation without hints failed its own check");
        }
      }
      min_dt = std::min(min_dt, total_ratio(ctx, l) * d);
    }
    *dt_out = min_dt;
    return (int)FVS_OK;
  });
}

int fvs_advance_hierarchy(fvs_ctx* ctx, double dt, int split_kind, double* failure_dt,
                          double* sweep_dts, int* n_sweep_dts) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    if (split_kind < 0 || split_kind > 2) throw Error("unknown splitting method");
    if (num_levels(ctx) > 1) {
      return advance_hierarchy_multilevel(ctx, dt, split_kind, failure_dt, sweep_dts, n_sweep_dts);
    }
    Level& L = ctx->level(0);
    const int rank = ctx->cfg.rank;
    if (split_kind < 0 || split_kind > 2) throw Error("unknown splitting method");
    // MakeSplitDirections, include/fub/split_method/SplittingMethod.hpp:115-126
    const bool is_odd = (ctx->cycles % 2) != 0;
    int dirs[3];
    for (int i = 0; i < rank; ++i) dirs[i] = is_odd ? (rank - 1 - i) : i;
    const std::vector<SweepPlanEntry> plan = split_plan(rank, split_kind, dt);
    if ((int)plan.size() > kDtSlots) throw Error("too many sweeps");
    const int entry = L.cur;
    const int old_keep = L.keep;
    // Speculative step: all sweeps are launched back to back (each also evaluates
    // the stable dt of its input state, which the reference checks before the
    // sweep), read back with ONE synchronisation and checked in order.  The lazy
    // stable dt runs with hints first; a hint that turns out too high (the max
    // signal speed fell by more than 10 % since its last evaluation) invalidates
    // the result of that sweep, and the step is repeated from the entry buffer
    // without hints.
    for (int attempt = 0;; ++attempt) {
      const bool use_hints = attempt == 0;
      L.keep = entry; // the state at entry survives until the step is accepted
      int counts[3] = {0, 0, 0};
      std::vector<int> slots;
      std::vector<double> hints;
      ctx->next_slot = 0;
      try {
        // cutcell::IntegratorContext::PreAdvanceLevel, :802-806
        if (L.has_eb) cc_pre_advance(ctx, L);
        for (const SweepPlanEntry& e : plan) {
          const int dir = dirs[e.slot];
          if (counts[e.slot] > 0) apply_boundary_condition(ctx, L, dir); // :228-234
          counts[e.slot] += 1;
          if (L.has_eb) {
            int dst;
            slots.push_back(eb_sweep(ctx, L, e.dt, dir, &dst));
            hints.push_back(0.0);
            L.cur = dst;
            continue;
          }
          int dst;
          const double hint = use_hints ? dt_hint(ctx, L, dir) : 0.0;
          slots.push_back(launch_sweep(ctx, L, e.dt, dir, &dst, hint));
          hints.push_back(hint);
          L.cur = dst;
        }
      } catch (...) {
        L.cur = entry;
        L.keep = old_keep;
        throw;
      }
      // one synchronisation for all sweeps of the step
      std::vector<double> sig(plan.size());
      std::vector<int> fails(plan.size());
      // slots are consecutive (next_slot was reset above)
      try {
        fetch_slots(ctx, slots.front(), (int)plan.size(), sig.data(), fails.data());
      } catch (...) {
        L.cur = entry;
        L.keep = old_keep;
        throw;
      }
      int n_out = 0;
      const int cap = (sweep_dts && n_sweep_dts) ? *n_sweep_dts : 0;
      int status = FVS_OK;
      bool redo = false;
      bool newton = false;
      for (size_t i = 0; i < plan.size(); ++i) {
        const int dir = dirs[plan[i].slot];
        if (!verify_hint(hints[i], sig[i])) {
          redo = true;
          break;
        }
        L.last_sig[dir] = sig[i];
        const double level_dt = dt_from_signal(L.dx[dir], sig[i]);
        if (n_out < cap) sweep_dts[n_out] = level_dt;
        ++n_out;
        if (level_dt < plan[i].dt) { // :243-245
          // the reference stops here: nothing behind this check ever ran, so a
          // Newton failure of this or a later sweep is not an error
          if (failure_dt) *failure_dt = level_dt;
          status = FVS_TIME_STEP_TOO_LARGE;
          break;
        }
        if (fails[i]) {
          newton = true;
          break;
        }
      }
      L.keep = old_keep;
      if (redo) {
        ctx->hint_redos += 1;
        L.cur = entry;
        continue;
      }
      if (newton) {
        L.cur = entry;
        throw_newton();
      }
      if (n_sweep_dts) *n_sweep_dts = n_out;
      if (status != FVS_OK) {
        L.cur = entry; // roll back: the entry buffer was never written
        return status;
      }
      break;
    }
    // SubcycleFineFirstSolver.hpp:164-174,177-192
    ctx->cycles += 1;
    ctx->time_point = ctx->time_point + dt;
    if (ctx->cycles % ctx->regrid_frequency == 0) fill_ghost_layer(ctx, L);
    return (int)FVS_OK;
  });
}

long fvs_run(fvs_ctx* ctx, double cfl, double final_time, long max_cycles, int split_kind,
             double* dt_sequence, long dt_capacity, int* retries_out) {
  long accepted = 0;
  const int rc = guarded([&] {
    if (!ctx) throw Error("null context");
    const double eps = 1e-12; // RunOptions::smallest_time_step_size, RunSimulation.hpp:61-65
    bool have_failure = false;
    double failure_dt = 0.0;
    int retries = 0;
    while (ctx->time_point + eps < final_time && (max_cycles < 0 || ctx->cycles < max_cycles)) {
      double stable_dt;
      if (have_failure) {
        stable_dt = failure_dt;
      } else if (fvs_compute_stable_dt_hierarchy(ctx, &stable_dt) != FVS_OK) {
        throw Error(g_error);
      }
      const double limited_dt = std::min(final_time - ctx->time_point, cfl * stable_dt);
      double fail = 0.0;
      const int status = fvs_advance_hierarchy(ctx, limited_dt, split_kind, &fail, nullptr, nullptr);
      if (status == FVS_ERROR) throw Error(g_error);
      if (status == FVS_TIME_STEP_TOO_LARGE) {
        have_failure = true;
        failure_dt = fail;
        if (++retries > 100) throw Error("too many CFL retries");
      } else {
        if (dt_sequence && accepted < dt_capacity) dt_sequence[accepted] = limited_dt;
        ++accepted;
        have_failure = false;
      }
    }
    if (retries_out) *retries_out = retries;
    return (int)FVS_OK;
  });
  return rc == FVS_OK ? accepted : (long)FVS_ERROR;
}

int fvs_set_regrid_frequency(fvs_ctx* ctx, int frequency) {
  return guarded([&] {
    if (!ctx || frequency < 1) throw Error("invalid regrid frequency");
    ctx->regrid_frequency = frequency;
    return (int)FVS_OK;
  });
}
int fvs_set_cycles(fvs_ctx* ctx, long cycles, double time_point) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->cycles = cycles;
    ctx->time_point = time_point;
    return (int)FVS_OK;
  });
}
long fvs_cycles(const fvs_ctx* ctx) { return ctx ? ctx->cycles : -1; }
double fvs_time_point(const fvs_ctx* ctx) { return ctx ? ctx->time_point : 0.0; }

int fvs_min_all(fvs_ctx* ctx, double* inout) {
  return guarded([&] {
    if (!ctx || !inout) throw Error("null argument");
    if (!ctx->comm || ctx->cfg.comm_size <= 1) return (int)FVS_OK;
    ctx->use_device();
    Nccl& n = Nccl::get();
    double* d = reinterpret_cast<double*>(ctx->d_slots + (kDtSlots - 1));
    FVS_CUDA(cudaMemcpyAsync(d, inout, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    n.check(n.AllReduce(d, d, 1, Nccl::kDouble, Nccl::kMin, ctx->comm, ctx->stream),
            "ncclAllReduce(min)");
    FVS_CUDA(cudaMemcpyAsync(inout, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_comm_unique_id(void* out_id) {
  return guarded([&] {
    if (!out_id) throw Error("null argument");
    Nccl& n = Nccl::get();
    Nccl::unique_id id;
    n.check(n.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out_id, &id, sizeof(id));
    return (int)FVS_OK;
  });
}

int fvs_comm_init(fvs_ctx* ctx, const void* id_bytes) {
  return guarded([&] {
    if (!ctx || !id_bytes) throw Error("null argument");
    ctx->use_device();
    Nccl& n = Nccl::get();
    Nccl::unique_id id;
    std::memcpy(&id, id_bytes, sizeof(id));
    n.check(n.CommInitRank(&ctx->comm, ctx->cfg.comm_size, id, ctx->cfg.comm_rank),
            "ncclCommInitRank");
    return (int)FVS_OK;
  });
}

long fvs_halo_plan(int rank, const int domain_lo[3], const int domain_hi[3],
                   const int periodic[3], int nboxes, const int (*box_lo)[3],
                   const int (*box_hi)[3], const int* owner, int scratch_gcw,
                   fvs_halo_copy* out, long capacity) {
  long count = 0;
  const int rc = guarded([&] {
    std::vector<Box> boxes(nboxes);
    for (int b = 0; b < nboxes; ++b) boxes[b] = fvs::make_box(box_lo[b], box_hi[b]);
    bool per[3];
    for (int d = 0; d < 3; ++d) per[d] = periodic && periodic[d] != 0;
    const auto plan =
        fvs::halo_plan(rank, fvs::make_box(domain_lo, domain_hi), per, boxes, scratch_gcw);
    for (const fvs::HaloCopy& c : plan) {
      if (out && count < capacity) {
        fvs_halo_copy& o = out[count];
        o.dst_box = c.dst_box;
        o.src_box = c.src_box;
        o.dst_owner = owner ? owner[c.dst_box] : 0;
        o.src_owner = owner ? owner[c.src_box] : 0;
        for (int d = 0; d < 3; ++d) {
          o.lo[d] = c.region.lo[d];
          o.hi[d] = c.region.hi[d];
          o.shift[d] = c.shift[d];
        }
      }
      ++count;
    }
    return (int)FVS_OK;
  });
  return rc == FVS_OK ? count : (long)FVS_ERROR;
}

int fvs_stable_dt_face_mask(int rank, const int box_lo[3], const int box_hi[3], int scratch_gcw,
                            int stencil_width, int dir, unsigned char* mask, int capacity) {
  return guarded([&] {
    if (dir < 0 || dir >= rank) throw Error("direction out of range");
    const auto m = fvs::stable_dt_face_mask(fvs::make_box(box_lo, box_hi), scratch_gcw,
                                            stencil_width, dir);
    if (mask) {
      for (int i = 0; i < (int)m.size() && i < capacity; ++i) mask[i] = m[i];
    }
    return (int)m.size();
  });
}

int fvs_distribute_boxes(int rank, int nboxes, const int (*box_lo)[3], const int (*box_hi)[3],
                         int nranks, int* owner_out) {
  return guarded([&] {
    (void)rank;
    if (nranks < 1 || !owner_out) throw Error("invalid argument");
    std::vector<Box> boxes(nboxes);
    for (int b = 0; b < nboxes; ++b) boxes[b] = fvs::make_box(box_lo[b], box_hi[b]);
    const std::vector<int> owner = fvs::distribute_boxes(boxes, nranks);
    for (int b = 0; b < nboxes; ++b) owner_out[b] = owner[b];
    return (int)FVS_OK;
  });
}

int fvs_distribute_boxes_cutcell(int rank, int nboxes, const int (*box_lo)[3],
                                 const int (*box_hi)[3], const int* fab_type,
                                 const long long* covered_cells, double cutcell_weight, int nranks,
                                 int* owner_out) {
  return guarded([&] {
    (void)rank;
    if (nranks < 1 || !owner_out || !fab_type) throw Error("invalid argument");
    std::vector<Box> boxes(nboxes);
    for (int b = 0; b < nboxes; ++b) boxes[b] = fvs::make_box(box_lo[b], box_hi[b]);
    const std::vector<int> owner = fvs::knapsack(
        fvs::cutcell_box_weights(boxes, fab_type, covered_cells, cutcell_weight), nranks);
    for (int b = 0; b < nboxes; ++b) owner_out[b] = owner[b];
    return (int)FVS_OK;
  });
}

long fvs_cluster_tags(int rank, const int domain_lo[3], const int domain_hi[3],
                      const int periodic[3], int nboxes, const int (*box_lo)[3],
                      const int (*box_hi)[3], const char* tags, const int ratio[3],
                      int blocking_factor, int max_grid_size, int nesting_buffer,
                      int (*out_lo)[3], int (*out_hi)[3], long capacity) {
  long count = 0;
  const int rc = guarded([&] {
    if (rank < 1 || rank > 3 || !tags || !ratio) throw Error("invalid argument");
    if (blocking_factor < 1 || max_grid_size < blocking_factor) {
      throw Error("need 1 <= blocking_factor <= max_grid_size");
    }
    for (int d = 0; d < rank; ++d) {
      if (ratio[d] < 1 || blocking_factor % ratio[d] != 0) {
        throw Error("blocking_factor must be a multiple of the refinement ratio");
      }
    }
    std::vector<Box> boxes(nboxes);
    for (int b = 0; b < nboxes; ++b) boxes[b] = fvs::make_box(box_lo[b], box_hi[b]);
    bool per[3];
    for (int d = 0; d < 3; ++d) per[d] = periodic && periodic[d] != 0;
    const std::vector<Box> out =
        fvs::cluster_tags(rank, fvs::make_box(domain_lo, domain_hi), per, boxes, tags, ratio,
                          blocking_factor, max_grid_size, nesting_buffer);
    count = (long)out.size();
    for (long i = 0; i < count && i < capacity; ++i)
      for (int d = 0; d < 3; ++d) {
        out_lo[i][d] = out[(size_t)i].lo[d];
        out_hi[i][d] = out[(size_t)i].hi[d];
      }
    return (int)FVS_OK;
  });
  return rc == FVS_OK ? count : (long)FVS_ERROR;
}

// ---- embedded boundary ---------------------------------------------------------
int fvs_level_set_eb(fvs_ctx* ctx, int level, int box, const double* alpha,
                     const double* const beta[3], const double* bcent, const double* bnorm,
                     int fab_type) {
  return guarded([&] {
    if (!ctx || !alpha || !beta || !bcent || !bnorm) throw Error("null argument");
    if (fab_type < FVS_FAB_COVERED || fab_type > FVS_FAB_SINGLEVALUED) throw Error("unknown fab type");
    ctx->use_device();
    Level& L = ctx->level(level);
    const int rank = ctx->cfg.rank;
    enable_eb(ctx, L);
    Patch& p = ctx->patch(L, box);
    EbPatch& e = p.eb;
    e.fab_type = fab_type;
    const size_t cells = (size_t)p.cells;
    FVS_CUDA(cudaMemcpyAsync(e.alpha, alpha, cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FVS_CUDA(cudaMemcpyAsync(e.bcent, bcent, cells * rank * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FVS_CUDA(cudaMemcpyAsync(e.bnorm, bnorm, cells * rank * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    for (int d = 0; d < rank; ++d) {
      if (!beta[d]) throw Error("null face-fraction array");
      const Box bb = beta_box(p, d);
      FVS_CUDA(cudaMemcpyAsync(e.beta[d], beta[d], (size_t)bb.size() * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
      const int bn[3] = {bb.len(0), bb.len(1), bb.len(2)};
      FVS_CUDA(cudaMemsetAsync(e.us[d], 0, 2 * (size_t)bb.size() * sizeof(double), ctx->stream));
      FVS_CUDA(cudaMemsetAsync(e.sl[d], 0, 2 * (size_t)bb.size() * sizeof(double), ctx->stream));
      FVS_CUDA(cudaMemsetAsync(e.sr[d], 0, 2 * (size_t)bb.size() * sizeof(double), ctx->stream));
      FVS_CUDA(cudaMemsetAsync(e.ds[d], 0, 2 * (size_t)bb.size() * sizeof(double), ctx->stream));
      ctx->launches += fvs::launch_fill_cutcell_data(e.beta[d], e.us[d], e.sl[d], e.sr[d], e.ds[d],
                                                     bn, d, ctx->stream);
    }
    // IsCutCell map on the device, compacted into the cut-cell list
    ctx->launches += fvs::launch_cut_cell_flags(e.alpha, e.beta, p.n, rank, e.flags, ctx->stream);
    std::vector<unsigned char> flags(cells);
    FVS_CUDA(cudaMemcpyAsync(flags.data(), e.flags, cells, cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    check_last("launch check #8");
    std::vector<long long> list;
    for (size_t c = 0; c < cells; ++c)
      if (flags[c]) list.push_back((long long)c);
    if (e.cut_cells) cudaFree(e.cut_cells);
    e.cut_cells = nullptr;
    e.n_cut = (int)list.size();
    if (!list.empty()) {
      FVS_CUDA(cudaMalloc(&e.cut_cells, list.size() * sizeof(long long)));
      FVS_CUDA(upload_sync(e.cut_cells, list.data(), list.size() * sizeof(long long)));
    }
    const std::vector<unsigned char> mode = complete_modes(rank, L, p, alpha, beta);
    FVS_CUDA(upload_sync(e.complete_mode, mode.data(), cells));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    L.box_cache.clear();
    L.eb_cache.clear();
    return (int)FVS_OK;
  });
}

int fvs_set_cutcell_riemann_solver(fvs_ctx* ctx, int kind) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (kind != FVS_RIEMANN_HLL_EINFELDT && kind != FVS_RIEMANN_EXACT) {
      throw Error("unknown Riemann solver");
    }
    ctx->cutcell_riemann = kind;
    return (int)FVS_OK;
  });
}

int fvs_pre_advance_hierarchy(fvs_ctx* ctx) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    for (Level& L : ctx->levels) {
      if (L.defined) cc_pre_advance(ctx, L);
    }
    check_last("launch check #9");
    return (int)FVS_OK;
  });
}

namespace {
// device pointer, box and component count of an EB array
double* eb_array(fvs_ctx* ctx, Level& L, Patch& p, int kind, int dir, Box* box, int* ncomp) {
  if (!L.has_eb) throw Error("level has no embedded boundary (fvs_level_set_eb)");
  const int rank = ctx->cfg.rank;
  const bool per_dir = kind == FVS_EB_BETA || (kind >= FVS_EB_UNSHIELDED && kind <= FVS_EB_DOUBLY_SHIELDED) ||
                       kind >= FVS_EB_STABILIZED_FLUXES;
  if (per_dir) check_dir(ctx, dir);
  EbPatch& e = p.eb;
  switch (kind) {
  case FVS_EB_ALPHA: *box = p.sbox; *ncomp = 1; return e.alpha;
  case FVS_EB_BETA: *box = beta_box(p, dir); *ncomp = 1; return e.beta[dir];
  case FVS_EB_BCENT: *box = p.sbox; *ncomp = rank; return e.bcent;
  case FVS_EB_BNORM: *box = p.sbox; *ncomp = rank; return e.bnorm;
  case FVS_EB_UNSHIELDED: *box = beta_box(p, dir); *ncomp = 2; return e.us[dir];
  case FVS_EB_SHIELDED_LEFT: *box = beta_box(p, dir); *ncomp = 2; return e.sl[dir];
  case FVS_EB_SHIELDED_RIGHT: *box = beta_box(p, dir); *ncomp = 2; return e.sr[dir];
  case FVS_EB_DOUBLY_SHIELDED: *box = beta_box(p, dir); *ncomp = 2; return e.ds[dir];
  case FVS_EB_REFERENCE_STATES: *box = p.sbox; *ncomp = ctx->ncomp(); return e.reference;
  case FVS_EB_BOUNDARY_FLUXES: *box = p.sbox; *ncomp = ctx->ncons(); return e.boundary_flux;
  case FVS_EB_STABILIZED_FLUXES: *box = p.fbox[dir]; *ncomp = ctx->ncons(); return e.stab[dir];
  case FVS_EB_SHIELDED_LEFT_FLUXES: *box = p.fbox[dir]; *ncomp = ctx->ncons(); return e.shl[dir];
  case FVS_EB_SHIELDED_RIGHT_FLUXES: *box = p.fbox[dir]; *ncomp = ctx->ncons(); return e.shr[dir];
  }
  throw Error("unknown EB array kind");
}
} // namespace

int fvs_eb_array_box(const fvs_ctx* cctx, int level, int box, int kind, int dir, int lo[3],
                     int hi[3], int* ncomp) {
  return guarded([&] {
    fvs_ctx* ctx = const_cast<fvs_ctx*>(cctx);
    if (!ctx || !lo || !hi) throw Error("null argument");
    Level& L = ctx->level(level);
    Patch& p = ctx->patch(L, box);
    Box b;
    int nc = 0;
    eb_array(ctx, L, p, kind, dir, &b, &nc);
    for (int d = 0; d < 3; ++d) {
      lo[d] = b.lo[d];
      hi[d] = b.hi[d];
    }
    if (ncomp) *ncomp = nc;
    return (int)FVS_OK;
  });
}

int fvs_download_eb_array(fvs_ctx* ctx, int level, int box, int kind, int dir, double* host_fab) {
  return guarded([&] {
    if (!ctx || !host_fab) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    Patch& p = ctx->patch(L, box);
    Box b;
    int nc = 0;
    const double* src = eb_array(ctx, L, p, kind, dir, &b, &nc);
    FVS_CUDA(cudaMemcpyAsync(host_fab, src, (size_t)b.size() * nc * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

long fvs_cut_cell_flags(fvs_ctx* ctx, int level, int box, unsigned char* flags) {
  long count = 0;
  const int rc = guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    Level& L = ctx->level(level);
    if (!L.has_eb) throw Error("level has no embedded boundary (fvs_level_set_eb)");
    Patch& p = ctx->patch(L, box);
    if (flags) {
      FVS_CUDA(cudaMemcpyAsync(flags, p.eb.flags, (size_t)p.cells, cudaMemcpyDeviceToHost,
                               ctx->stream));
      FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    count = p.eb.n_cut;
    return (int)FVS_OK;
  });
  return rc == FVS_OK ? count : (long)FVS_ERROR;
}

// ---- coarse-fine operations ---------------------------------------------------------
int fvs_set_refine_ratio(fvs_ctx* ctx, const int ratio[3]) {
  return guarded([&] {
    if (!ctx || !ratio) throw Error("null argument");
    for (int d = 0; d < ctx->cfg.rank; ++d)
      if (ratio[d] < 1) throw Error("refinement ratio must be >= 1");
    for (int d = 0; d < 3; ++d) ctx->ratio[d] = ratio[d];
    for (Level& L : ctx->levels) L.cf_coarse = -1;
    return (int)FVS_OK;
  });
}

int fvs_reset_coarse_fine_fluxes(fvs_ctx* ctx, int fine, int coarse) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (fine <= coarse || coarse < 0) throw Error("need 0 <= coarse < fine");
    ctx->use_device();
    Level& F = ctx->level(fine);
    Level& Cl = ctx->level(coarse);
    if (F.cf_coarse != coarse) define_flux_register(ctx, F, Cl, coarse);
    for (Level::CfRegister& r : F.cf_regs) {
      FVS_CUDA(cudaMemsetAsync(r.reg, 0, (size_t)r.plane.size() * ctx->ncons() * sizeof(double),
                               ctx->stream));
    }
    return (int)FVS_OK;
  });
}

int fvs_accumulate_coarse_fine_fluxes(fvs_ctx* ctx, int level, double scale, int dir) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    check_dir(ctx, dir);
    Level& L = ctx->level(level);
    ensure_fluxes(ctx, L, dir);
    const double vol = cell_volume(ctx, L);
    const double vol_dt_over_dx = vol * scale / L.dx[dir];
    if (level > 0 && L.cf_coarse >= 0) fine_add(ctx, L, dir, vol_dt_over_dx);
    if (level + 1 < (int)ctx->levels.size() && ctx->levels[level + 1].defined &&
        ctx->levels[level + 1].cf_coarse == level) {
      crse_init(ctx, ctx->levels[level + 1], L, dir, -vol_dt_over_dx);
    }
    return (int)FVS_OK;
  });
}

int fvs_apply_flux_correction(fvs_ctx* ctx, int fine, int coarse, double /*dt*/) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    Level& F = ctx->level(fine);
    Level& Cl = ctx->level(coarse);
    if (F.cf_coarse != coarse) throw Error("flux register is not defined (fvs_reset_coarse_fine_fluxes)");
    for (int d = 0; d < ctx->cfg.rank; ++d) reflux(ctx, F, Cl, d, 1.0);
    return (int)FVS_OK;
  });
}

int fvs_coarsen_conservatively(fvs_ctx* ctx, int fine, int coarse) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (fine <= coarse || coarse < 0) throw Error("need 0 <= coarse < fine");
    ctx->use_device();
    average_down(ctx, ctx->level(fine), ctx->level(coarse));
    return (int)FVS_OK;
  });
}

int fvs_fill_ghost_layer_two_levels(fvs_ctx* ctx, int fine, int coarse) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    if (fine <= coarse || coarse < 0) throw Error("need 0 <= coarse < fine");
    ctx->use_device();
    Level& F = ctx->level(fine);
    pc_interp(ctx, F, ctx->level(coarse));
    fill_ghost_layer(ctx, F);
    return (int)FVS_OK;
  });
}

// ---- tagging ---------------------------------------------------------------------
namespace {
struct DeviceTags {
  char* ptr{nullptr};
  size_t bytes;
  explicit DeviceTags(size_t n) : bytes(n) { FVS_CUDA(cudaMalloc(&ptr, std::max<size_t>(n, 1))); }
  ~DeviceTags() {
    if (ptr) cudaFree(ptr);
  }
};
void call_tag_buffer(fvs_ctx* ctx, int rank, char* tags, const int n[3], int width) {
  int nn[3] = {n[0], rank >= 2 ? n[1] : 1, rank >= 3 ? n[2] : 1};
  const size_t total = (size_t)nn[0] * nn[1] * nn[2];
  if (total == 0) return;
  DeviceTags dt(total), scratch(total);
  FVS_CUDA(cudaMemcpyAsync(dt.ptr, tags, total, cudaMemcpyHostToDevice, ctx->stream));
  const int rc = fvs::launch_tag_buffer(dt.ptr, scratch.ptr, rank, nn, width, ctx->stream);
  check_launch(rc);
  ctx->launches += rc;
  FVS_CUDA(cudaMemcpyAsync(tags, dt.ptr, total, cudaMemcpyDeviceToHost, ctx->stream));
  FVS_CUDA(cudaStreamSynchronize(ctx->stream));
}
} // namespace

int fvs_tag_gradient(fvs_ctx* ctx, int level, int box, int ncond, const int* comps,
                     const double* tols, int use_alpha, char* tags) {
  return guarded([&] {
    if (!ctx || !tags || (ncond > 0 && (!comps || !tols))) throw Error("null argument");
    if (ncond < 0 || ncond > 8) throw Error("at most 8 tagging conditions");
    for (int c = 0; c < ncond; ++c)
      if (comps[c] < 0 || comps[c] >= ctx->ncomp()) throw Error("component out of range");
    ctx->use_device();
    Level& L = ctx->level(level);
    if (L.scratch_gcw < 1) throw Error("gradient tagging needs one ghost cell");
    if (use_alpha && !L.has_eb) throw Error("level has no embedded boundary (fvs_level_set_eb)");
    Patch& p = ctx->patch(L, box);
    int lo[3], len[3];
    for (int d = 0; d < 3; ++d) {
      lo[d] = p.valid.lo[d] - p.sbox.lo[d];
      len[d] = p.valid.len(d);
    }
    const size_t n = (size_t)p.valid.size();
    DeviceTags dt(n);
    FVS_CUDA(cudaMemcpyAsync(dt.ptr, tags, n, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = fvs::launch_tag_gradient(p.buf[L.cur], p.cells, p.n, use_alpha ? p.eb.alpha : nullptr,
                                            ctx->cfg.rank, lo, len, ncond, comps, tols, dt.ptr,
                                            ctx->stream);
    check_launch(rc);
    ctx->launches += rc;
    FVS_CUDA(cudaMemcpyAsync(tags, dt.ptr, n, cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

// The tags of a whole level in one go (the regrid hand-off): GradientDetector on every
// owned box, one read-back, scattered into a dense byte map of the level's domain, then
// TagBuffer on that map.  tags is OR-ed into; on several ranks every rank fills the
// cells of its own boxes (the caller combines the maps).
int fvs_tag_level(fvs_ctx* ctx, int level, int ncond, const int* comps, const double* tols,
                  int use_alpha, int tag_cutcells, int buffer_width, char* tags) {
  return guarded([&] {
    if (!ctx || !tags || (ncond > 0 && (!comps || !tols))) throw Error("null argument");
    if (ncond < 0 || ncond > 8) throw Error("at most 8 tagging conditions");
    for (int c = 0; c < ncond; ++c)
      if (comps[c] < 0 || comps[c] >= ctx->ncomp()) throw Error("component out of range");
    ctx->use_device();
    Level& L = ctx->level(level);
    if (ncond > 0 && L.scratch_gcw < 1) throw Error("gradient tagging needs one ghost cell");
    if ((use_alpha || tag_cutcells) && !L.has_eb) {
      throw Error("level has no embedded boundary (fvs_level_set_eb)");
    }
    size_t total = 0;
    std::vector<size_t> offset;
    for (const Patch& p : L.patches) {
      offset.push_back(total);
      total += (size_t)p.valid.size();
    }
    std::vector<char> host(total);
    if (total > 0) {
      DeviceTags dt(total);
      FVS_CUDA(cudaMemsetAsync(dt.ptr, 0, total, ctx->stream));
      for (size_t i = 0; i < L.patches.size(); ++i) {
        Patch& p = L.patches[i];
        int lo[3], len[3];
        for (int d = 0; d < 3; ++d) {
          lo[d] = p.valid.lo[d] - p.sbox.lo[d];
          len[d] = p.valid.len(d);
        }
        if (ncond > 0) {
          const int rc = fvs::launch_tag_gradient(p.buf[L.cur], p.cells, p.n,
                                                  use_alpha ? p.eb.alpha : nullptr, ctx->cfg.rank, lo,
                                                  len, ncond, comps, tols, dt.ptr + offset[i],
                                                  ctx->stream);
          check_launch(rc);
          ctx->launches += rc;
        }
        if (tag_cutcells) {
          const int rc = fvs::launch_tag_cutcells(p.eb.alpha, p.n, lo, len, dt.ptr + offset[i],
                                                  ctx->stream);
          check_launch(rc);
          ctx->launches += rc;
        }
      }
      FVS_CUDA(cudaMemcpyAsync(host.data(), dt.ptr, total, cudaMemcpyDeviceToHost, ctx->stream));
      FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    const long long nx = L.domain.len(0), ny = L.domain.len(1);
    for (size_t i = 0; i < L.patches.size(); ++i) {
      const Patch& p = L.patches[i];
      const char* src = host.data() + offset[i];
      for (int k = p.valid.lo[2]; k <= p.valid.hi[2]; ++k)
        for (int j = p.valid.lo[1]; j <= p.valid.hi[1]; ++j) {
          char* dst = tags + (p.valid.lo[0] - L.domain.lo[0]) +
                      nx * ((j - L.domain.lo[1]) + ny * (long long)(k - L.domain.lo[2]));
          for (int x = 0; x < p.valid.len(0); ++x) dst[x] |= *src++;
        }
    }
    if (buffer_width > 0) {
      const int n[3] = {L.domain.len(0), L.domain.len(1), L.domain.len(2)};
      call_tag_buffer(ctx, ctx->cfg.rank, tags, n, buffer_width);
    }
    return (int)FVS_OK;
  });
}

int fvs_tag_cutcells(fvs_ctx* ctx, int level, int box, char* tags) {
  return guarded([&] {
    if (!ctx || !tags) throw Error("null argument");
    ctx->use_device();
    Level& L = ctx->level(level);
    if (!L.has_eb) throw Error("level has no embedded boundary (fvs_level_set_eb)");
    Patch& p = ctx->patch(L, box);
    int lo[3], len[3];
    for (int d = 0; d < 3; ++d) {
      lo[d] = p.valid.lo[d] - p.sbox.lo[d];
      len[d] = p.valid.len(d);
    }
    const size_t n = (size_t)p.valid.size();
    DeviceTags dt(n);
    FVS_CUDA(cudaMemcpyAsync(dt.ptr, tags, n, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = fvs::launch_tag_cutcells(p.eb.alpha, p.n, lo, len, dt.ptr, ctx->stream);
    check_launch(rc);
    ctx->launches += rc;
    FVS_CUDA(cudaMemcpyAsync(tags, dt.ptr, n, cudaMemcpyDeviceToHost, ctx->stream));
    FVS_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int)FVS_OK;
  });
}

int fvs_tag_buffer(fvs_ctx* ctx, int rank, char* tags, const int n[3], int width) {
  return guarded([&] {
    if (!ctx || !tags || !n) throw Error("null argument");
    if (rank < 1 || rank > 3 || width < 0) throw Error("invalid argument");
    ctx->use_device();
    call_tag_buffer(ctx, rank, tags, n, width);
    return (int)FVS_OK;
  });
}

// ---- checkpoint output --------------------------------------------------------------
namespace {
std::string box_text(const Box& b, int rank) {
  auto tuple = [rank](const int v[3]) {
    std::string t = "(";
    for (int d = 0; d < rank; ++d) t += (d ? "," : "") + std::to_string(v[d]);
    return t + ")";
  };
  const int zero[3] = {0, 0, 0};
  return "(" + tuple(b.lo) + " " + tuple(b.hi) + " " + tuple(zero) + ")";
}
std::string boxarray_text(const std::vector<Box>& boxes, int rank) { // BoxArray::writeOn
  std::string t = "(" + std::to_string(boxes.size()) + " 0\n";
  for (const Box& b : boxes) t += box_text(b, rank) + "\n";
  return t + ")";
}
// header line of one FAB on disk: native little-endian IEEE doubles (RealDescriptor)
std::string fab_header(const Box& b, int rank, int ncomp) {
  return "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))" + box_text(b, rank) + " " +
         std::to_string(ncomp) + "\n";
}
void make_dir(const std::string& path) {
  if (::mkdir(path.c_str(), 0755) != 0 && errno != EEXIST) {
    throw Error("cannot create directory " + path);
  }
}
std::string five(int n) {
  char buf[16];
  std::snprintf(buf, sizeof buf, "%05d", n);
  return buf;
}
} // namespace

int fvs_write_checkpoint(fvs_ctx* ctx, const char* directory) {
  return guarded([&] {
    if (!ctx || !directory) throw Error("null argument");
    ctx->use_device();
    const int rank = ctx->cfg.rank;
    const int ncomp = ctx->ncomp();
    const int nl = num_levels(ctx);
    const int me = ctx->cfg.comm_rank;
    const std::string root(directory);
    // PreBuildDirectorHierarchy: every rank may arrive first
    make_dir(root);
    for (int l = 0; l < nl; ++l) make_dir(root + "/Level_" + std::to_string(l));
    if (me == 0) { // PatchHierarchy.cpp:324-355
      std::ofstream h(root + "/Header", std::ios::out | std::ios::trunc | std::ios::binary);
      if (!h.good()) throw Error("cannot open " + root + "/Header");
      h.precision(17);
      h << "Checkpoint File\n" << nl - 1 << '\n';
      for (int l = 0; l < nl; ++l) h << ctx->cycles * total_ratio(ctx, l) << ' ';
      h << '\n';
      for (int l = 0; l < nl; ++l) h << ctx->time_point << ' ';
      h << '\n';
      for (int l = 0; l < nl; ++l) h << boxarray_text(ctx->levels[(size_t)l].boxes, rank) << '\n';
    }
    // VisMF::Write of the level data (valid cells, no ghost cells): one data file per rank,
    // the FABs of the rank's boxes in box order; rank 0 writes the VisMF header, whose
    // offsets follow from the (deterministic) FAB header lengths
    for (int l = 0; l < nl; ++l) {
      Level& L = ctx->levels[(size_t)l];
      const std::string prefix = root + "/Level_" + std::to_string(l) + "/data";
      std::vector<std::vector<double>> mins(L.boxes.size()), maxs(L.boxes.size());
      if (!L.patches.empty()) {
        std::ofstream out(prefix + "_D_" + five(me), std::ios::out | std::ios::trunc | std::ios::binary);
        if (!out.good()) throw Error("cannot open " + prefix + "_D_" + five(me));
        std::vector<double> host;
        for (Patch& p : L.patches) {
          const size_t cells = (size_t)p.valid.size();
          host.resize(cells * (size_t)ncomp);
          copy_valid(ctx, L, p, host.data(), false);
          FVS_CUDA(cudaStreamSynchronize(ctx->stream));
          const std::string fh = fab_header(p.valid, rank, ncomp);
          out.write(fh.data(), (std::streamsize)fh.size());
          out.write(reinterpret_cast<const char*>(host.data()), (std::streamsize)(host.size() * sizeof(double)));
          std::vector<double>& mn = mins[(size_t)p.global_index];
          std::vector<double>& mx = maxs[(size_t)p.global_index];
          for (int c = 0; c < ncomp; ++c) {
            const double* v = host.data() + cells * (size_t)c;
            mn.push_back(*std::min_element(v, v + cells));
            mx.push_back(*std::max_element(v, v + cells));
          }
        }
      }
      // the min / max tables of the header need every rank's values
      if (ctx->cfg.comm_size > 1) {
        std::vector<double> table(L.boxes.size() * (size_t)ncomp * 2, -std::numeric_limits<double>::max());
        for (size_t b = 0; b < L.boxes.size(); ++b)
          for (int c = 0; c < ncomp && !mins[b].empty(); ++c) {
            table[(b * (size_t)ncomp + (size_t)c) * 2] = -mins[b][(size_t)c];
            table[(b * (size_t)ncomp + (size_t)c) * 2 + 1] = maxs[b][(size_t)c];
          }
        double* dev = nullptr;
        FVS_CUDA(cudaMalloc(&dev, table.size() * sizeof(double)));
        FVS_CUDA(upload_sync(dev, table.data(), table.size() * sizeof(double)));
        Nccl& n = Nccl::get();
        if (!ctx->comm) throw Error("levels span several ranks but fvs_comm_init was not called");
        n.check(n.AllReduce(dev, dev, table.size(), Nccl::kDouble, Nccl::kMax, ctx->comm, ctx->stream),
                "ncclAllReduce(checkpoint min/max)");
        FVS_CUDA(cudaMemcpyAsync(table.data(), dev, table.size() * sizeof(double),
                                 cudaMemcpyDeviceToHost, ctx->stream));
        FVS_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(dev);
        for (size_t b = 0; b < L.boxes.size(); ++b) {
          mins[b].assign((size_t)ncomp, 0.0);
          maxs[b].assign((size_t)ncomp, 0.0);
          for (int c = 0; c < ncomp; ++c) {
            mins[b][(size_t)c] = -table[(b * (size_t)ncomp + (size_t)c) * 2];
            maxs[b][(size_t)c] = table[(b * (size_t)ncomp + (size_t)c) * 2 + 1];
          }
        }
      }
      if (me != 0) continue;
      std::ofstream h(prefix + "_H", std::ios::out | std::ios::trunc | std::ios::binary);
      if (!h.good()) throw Error("cannot open " + prefix + "_H");
      h.precision(17);
      std::string ngrow = "(";
      for (int d = 0; d < rank; ++d) ngrow += d ? ",0" : "0";
      ngrow += ")";
      h << 1 << '\n' << 1 << '\n' << ncomp << '\n' << ngrow << '\n';
      h << boxarray_text(L.boxes, rank) << '\n' << L.boxes.size() << '\n';
      std::vector<long long> offset((size_t)ctx->cfg.comm_size, 0);
      for (size_t b = 0; b < L.boxes.size(); ++b) {
        const int r = L.owner[b];
        h << "FabOnDisk: data_D_" << five(r) << ' ' << offset[(size_t)r] << '\n';
        offset[(size_t)r] += (long long)fab_header(L.boxes[b], rank, ncomp).size() +
                             L.boxes[b].size() * (long long)ncomp * (long long)sizeof(double);
      }
      h << '\n';
      for (int pass = 0; pass < 2; ++pass) {
        h << L.boxes.size() << ',' << ncomp << '\n';
        for (size_t b = 0; b < L.boxes.size(); ++b) {
          const std::vector<double>& v = pass == 0 ? mins[b] : maxs[b];
          for (int c = 0; c < ncomp; ++c) h << v[(size_t)c] << ',';
          h << '\n';
        }
        h << '\n';
      }
    }
    return (int)FVS_OK;
  });
}

long fvs_kernel_launches(const fvs_ctx* ctx) { return ctx ? ctx->launches : -1; }
int fvs_set_dt_hints(fvs_ctx* ctx, int on) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->dt_hints = on != 0;
    ctx->keep_hints_on_upload = on == 2;
    return (int)FVS_OK;
  });
}
long fvs_dt_hint_redos(const fvs_ctx* ctx) { return ctx ? ctx->hint_redos : -1; }
void* fvs_stream(fvs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int fvs_enable_timing(fvs_ctx* ctx, int on) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->timing = on != 0;
    ctx->sweep_ms = 0.0;
    ctx->sweep_count = 0;
    return (int)FVS_OK;
  });
}

int fvs_sweep_time_ms(fvs_ctx* ctx, double* total_ms, long* n_sweeps) {
  return guarded([&] {
    if (!ctx) throw Error("null context");
    ctx->use_device();
    drain_events(ctx);
    if (total_ms) *total_ms = ctx->sweep_ms;
    if (n_sweeps) *n_sweeps = ctx->sweep_count;
    return (int)FVS_OK;
  });
}

} // extern "C"
