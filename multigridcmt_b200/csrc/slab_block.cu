// slab_block.cu -- native driver of the row-slab multi-GPU step (include/mgcmt_b200.h, "slab block").
//
// multigridcmt_b200/slab.py states the decomposition and drives it from Python over torch.distributed; that is the
// testable reference of the host logic (gloo / single-process emulation), but at 8 GPUs a step is ~100 launches and
// ~12 exchange phases in ~4 ms, and the Python issue time alone is that long.  This file issues the same sequence --
// the same kernels through the same C entry points on the same data -- from C++: k V(4,4) cycles advanced phase by
// phase (each vector on its own stream), one NCCL group of send/recv pairs per halo phase, in-place all-gather of the
// restricted residuals before the replicated coarse part, all-reduce of the Rayleigh sums and of the packed Gram
// matrix.  The vectors form two halves, each with its own communicator, that run half a phase apart so that one half's
// exchange overlaps the other half's kernels.
//
// NCCL is reached through dlopen/dlsym on the library the process already has (the one PyTorch bundles), so this
// shared object has no link-time dependency on it and loads on machines without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mgcmt_b200.h"
#include "kernels.h"

using namespace mgcmt;

namespace {

// the slice of nccl.h this file uses (NCCL 2.x ABI: ncclUniqueId is 128 bytes, ncclFloat64 = 8, ncclSum = 0)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
struct Nccl {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
} g_nccl;

#define CU(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return set_error(MGCMT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)
#define NC(expr)                                                                                      \
  do {                                                                                                \
    int r__ = (expr);                                                                                 \
    if (r__ != 0)                                                                                     \
      return set_error(MGCMT_ERR_CUDA, std::string(#expr) + ": NCCL error " + std::to_string(r__) +   \
                                           (g_nccl.GetErrorString ? std::string(" (") + g_nccl.GetErrorString(r__) + ")" : "")); \
  } while (0)
#define RC(expr)                 \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != MGCMT_OK) return rc__; \
  } while (0)

int need_nccl() {
  if (!g_nccl.lib) return set_error(MGCMT_ERR_STATE, "NCCL is not loaded: call mgcmt_nccl_load first");
  return MGCMT_OK;
}

struct VecState {  // one vector of the block: its hierarchies, level buffers, stream
  mgcmt_hier_t *slab = nullptr, *coarse = nullptr;
  std::vector<double *> v, f, tmp;  // per slab level; entry 0 of v and f is the caller's array during a cycle
  double *fg = nullptr, *vg = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
};

}  // namespace

// The k vectors are split into two halves that advance through the phases of a cycle half a phase apart: while one half
// exchanges halos (NCCL, its own communicator and stream) the other runs its kernels, so the exchange latency of the
// many short phases is hidden instead of idling the GPU.
struct Half {
  int begin = 0, end = 0;        // vectors [begin, end)
  ncclComm_t comm = nullptr;
  cudaStream_t gs = nullptr;     // the half's ordering stream: NCCL calls, fork / join of its vectors' streams
  cudaEvent_t fork = nullptr, comp_done = nullptr;
};

struct mgcmt_slabblock {
  int world = 1, rank = 0, n = 0, own0 = 0, nlev = 0, k = 0;
  double omega = 2.0 / 3.0;
  Half half[2];
  std::vector<VecState> vec;
  cudaEvent_t fork = nullptr;
  double *scal = nullptr;  // 64 doubles
  int smoother = MGCMT_SMOOTH_WJACOBI;  // or MGCMT_SMOOTH_RBGS (mgcmt_slabblock_set_smoother)
  bool fused_rq = true;            // Rayleigh sums inside the finest up leg (MGCMT_SLAB_FUSED_RQ=0: separate stage)
  bool profile = false;            // time every stage of the next cycles with CUDA events (lock-step form only)
  std::vector<cudaEvent_t> prof;   // 2 * nstage + 1 events of the last profiled cycle
};

namespace {

constexpr int kHalo = 10;  // multigridcmt_b200/slab.py: HALO (cone of the longest fused leg: 8 Gauss-Seidel colour stages + transfer)

size_t level_elems(const mgcmt_slabblock *b, int l) { return (size_t)((b->own0 >> l) + 2 * kHalo) * (size_t)(b->n >> l); }

int fork_streams(mgcmt_slabblock *b, Half &h) {
  CU(cudaEventRecord(h.fork, h.gs));
  for (int c = h.begin; c < h.end; ++c) CU(cudaStreamWaitEvent(b->vec[c].stream, h.fork, 0));
  return MGCMT_OK;
}

int join_streams(mgcmt_slabblock *b, Half &h) {
  for (int c = h.begin; c < h.end; ++c) {
    CU(cudaEventRecord(b->vec[c].done, b->vec[c].stream));
    CU(cudaStreamWaitEvent(h.gs, b->vec[c].done, 0));
  }
  return MGCMT_OK;
}

// halo rows of level-l slab arrays: send the first / last kHalo owned rows to the neighbour above / below and receive
// its rows into the halo; all arrays of one phase of one half in one NCCL group
struct HaloItem {
  double *x;
  int level;
};
int exchange(mgcmt_slabblock *b, Half &h, const std::vector<HaloItem> &items) {
  if (b->world == 1 || items.empty()) return MGCMT_OK;
  NC(g_nccl.GroupStart());
  for (const HaloItem &it : items) {
    const size_t cols = (size_t)(b->n >> it.level), own = (size_t)(b->own0 >> it.level), cnt = kHalo * cols;
    double *top_halo = it.x, *top_own = it.x + kHalo * cols, *bot_own = it.x + own * cols, *bot_halo = it.x + (own + kHalo) * cols;
    if (b->rank > 0) {
      NC(g_nccl.Send(top_own, cnt, kNcclFloat64, b->rank - 1, h.comm, h.gs));
      NC(g_nccl.Recv(top_halo, cnt, kNcclFloat64, b->rank - 1, h.comm, h.gs));
    }
    if (b->rank + 1 < b->world) {
      NC(g_nccl.Send(bot_own, cnt, kNcclFloat64, b->rank + 1, h.comm, h.gs));
      NC(g_nccl.Recv(bot_halo, cnt, kNcclFloat64, b->rank + 1, h.comm, h.gs));
    }
  }
  NC(g_nccl.GroupEnd());
  return MGCMT_OK;
}

// The phases of a cycle for one half; comm(i) precedes comp(i).
//   weighted Jacobi (and the 5-point level of the red-black smoother: one pass of 4 sweeps):
//     DOWN(l)    comm: halos of f[l]                           comp: down leg of level l (zero start)      -> tmp[l], f[l+1]
//     UP(l)      comm: halos of tmp[l] (and of v[l+1])         comp: up leg of level l                     -> v[l]
//   red-black on the 9-point levels (two passes of two sweeps per leg, the intermediate iterate exchanged in between):
//     DOWN_A(l)  comm: halos of f[l]                           comp: 2 sweeps from zero                    -> tmp[l]
//     DOWN_B(l)  comm: halos of tmp[l]                         comp: 2 sweeps + residual + restriction     -> v[l], f[l+1]
//     UP_A(l)    comm: halos of v[l] (and of v[l+1])           comp: correction + 2 sweeps                 -> tmp[l]
//     UP_B(l)    comm: halos of tmp[l]                         comp: 2 sweeps                              -> v[l]
//   COARSE       comm: all-gather of the restricted residual   comp: replicated coarse cycle
//   RQ           comm: halos of the result v[0]                comp: Rayleigh sums (only when they are not taken inside
//                                                                     the last up leg)
enum PhaseKind { PH_DOWN, PH_DOWN_A, PH_DOWN_B, PH_COARSE, PH_UP, PH_UP_A, PH_UP_B, PH_RQ };
struct Phase {
  PhaseKind kind;
  int level;
};

std::vector<Phase> build_phases(int nl, bool gs, bool with_rq_stage) {
  std::vector<Phase> ph;
  for (int l = 0; l < nl; ++l) {
    if (gs && l > 0) { ph.push_back({PH_DOWN_A, l}); ph.push_back({PH_DOWN_B, l}); }
    else ph.push_back({PH_DOWN, l});
  }
  ph.push_back({PH_COARSE, nl});
  for (int l = nl - 1; l >= 0; --l) {
    if (gs && l > 0) { ph.push_back({PH_UP_A, l}); ph.push_back({PH_UP_B, l}); }
    else ph.push_back({PH_UP, l});
  }
  if (with_rq_stage) ph.push_back({PH_RQ, 0});
  return ph;
}
std::vector<Phase> build_phases(const mgcmt_slabblock *b, bool with_rq_stage) {
  return build_phases(b->nlev, b->smoother == MGCMT_SMOOTH_RBGS, with_rq_stage);
}

int comm_stage(mgcmt_slabblock *b, Half &h, const Phase &p) {
  const int nl = b->nlev, l = p.level;
  const bool gs = (b->smoother == MGCMT_SMOOTH_RBGS);
  std::vector<HaloItem> items;
  auto add = [&](std::vector<double *> VecState::*member, int lev) {
    for (int c = h.begin; c < h.end; ++c) items.push_back({(b->vec[c].*member)[lev], lev});
  };
  switch (p.kind) {
    case PH_DOWN:
    case PH_DOWN_A:
      add(&VecState::f, l);
      break;
    case PH_DOWN_B:
    case PH_UP_B:
      add(&VecState::tmp, l);
      break;
    case PH_COARSE: {
      if (b->world == 1) return MGCMT_OK;
      const size_t cnt = (size_t)(b->own0 >> nl) * (size_t)(b->n >> nl);
      NC(g_nccl.GroupStart());
      for (int c = h.begin; c < h.end; ++c) {
        VecState &s = b->vec[c];
        NC(g_nccl.AllGather(s.fg + (size_t)b->rank * cnt, s.fg, cnt, kNcclFloat64, h.comm, h.gs));
      }
      NC(g_nccl.GroupEnd());
      return MGCMT_OK;
    }
    case PH_UP:
      add(&VecState::tmp, l);   // the smoothed iterate of the down leg (Jacobi; red-black on the 5-point level)
      if (l + 1 < nl) add(&VecState::v, l + 1);
      break;
    case PH_UP_A:
      add(&VecState::v, l);     // red-black, 9-point level: the down leg left its iterate in v[l]
      if (l + 1 < nl) add(&VecState::v, l + 1);
      break;
    case PH_RQ:
      add(&VecState::v, 0);
      break;
  }
  (void)gs;
  return exchange(b, h, items);
}

int comp_stage(mgcmt_slabblock *b, Half &h, const Phase &p, const double *shifts, double *d_lam) {
  const int nl = b->nlev, l = p.level;
  const bool gs = (b->smoother == MGCMT_SMOOTH_RBGS);
  const int G = gs ? 32 : 0;  // mgcmt_fused_leg: Gauss-Seidel colour sweeps
  RC(fork_streams(b, h));
  for (int c = h.begin; c < h.end; ++c) {
    VecState &s = b->vec[c];
    const double sh = shifts[c];
    switch (p.kind) {
      case PH_DOWN:
        RC(mgcmt_fused_leg(s.slab, l, G | 2 /* down, zero start */, 4, sh, b->omega, nullptr, s.f[l], s.tmp[l], nullptr,
                           (l + 1 == nl) ? s.fg : s.f[l + 1], s.stream));
        break;
      case PH_DOWN_A:
        CU(cudaMemsetAsync(s.v[l], 0, sizeof(double) * level_elems(b, l), s.stream));
        RC(mgcmt_fused_leg(s.slab, l, G | 0 /* smooth */, 2, sh, b->omega, s.v[l], s.f[l], s.tmp[l], nullptr, nullptr, s.stream));
        break;
      case PH_DOWN_B:
        RC(mgcmt_fused_leg(s.slab, l, G | 1 /* down */, 2, sh, b->omega, s.tmp[l], s.f[l], s.v[l], nullptr,
                           (l + 1 == nl) ? s.fg : s.f[l + 1], s.stream));
        break;
      case PH_COARSE:
        RC(mgcmt_vcycle_from(s.coarse, nl, sh, b->smoother, b->omega, s.vg, s.fg, s.stream));
        break;
      case PH_UP: {
        const double *e = (l + 1 == nl) ? s.vg : s.v[l + 1];
        if (l == 0 && d_lam && b->fused_rq)   // the Rayleigh sums of the result are taken inside the last up leg
          RC(mgcmt_slab_up_rq(s.slab, gs ? 1 : 0, sh, b->omega, s.tmp[0], s.f[0], s.v[0], e, d_lam + 2 * c, s.stream));
        else
          RC(mgcmt_fused_leg(s.slab, l, G | 3 /* up */, 4, sh, b->omega, s.tmp[l], s.f[l], s.v[l], e, nullptr, s.stream));
        break;
      }
      case PH_UP_A: {
        const double *e = (l + 1 == nl) ? s.vg : s.v[l + 1];
        RC(mgcmt_fused_leg(s.slab, l, G | 3 /* up */, 2, sh, b->omega, s.v[l], s.f[l], s.tmp[l], e, nullptr, s.stream));
        break;
      }
      case PH_UP_B:
        RC(mgcmt_fused_leg(s.slab, l, G | 0 /* smooth */, 2, sh, b->omega, s.tmp[l], s.f[l], s.v[l], nullptr, nullptr, s.stream));
        break;
      case PH_RQ:
        RC(mgcmt_slab_rayleigh(s.slab, 0, s.v[0], d_lam + 2 * c, s.stream));
        break;
    }
  }
  RC(join_streams(b, h));
  CU(cudaEventRecord(h.comp_done, h.gs));
  return MGCMT_OK;
}

}  // namespace

extern "C" {

int mgcmt_nccl_load(const char *path) {
  if (g_nccl.lib) return MGCMT_OK;
  const char *name = (path && *path) ? path : "libnccl.so.2";
  void *lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return set_error(MGCMT_ERR_STATE, std::string("cannot load NCCL (") + name + "): " + dlerror());
#define SYM(field, sym)                                                                              \
  do {                                                                                               \
    *(void **)(&g_nccl.field) = dlsym(lib, sym);                                                     \
    if (!g_nccl.field) return set_error(MGCMT_ERR_STATE, std::string("NCCL symbol missing: ") + sym); \
  } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllGather, "ncclAllGather");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = lib;
  return MGCMT_OK;
}

int mgcmt_nccl_unique_id(void *out128) {
  RC(need_nccl());
  if (!out128) return set_error(MGCMT_ERR_ARG, "null argument");
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(out128, id.internal, sizeof(id.internal));
  return MGCMT_OK;
}

int mgcmt_nccl_comm_create(const void *id128, int world, int rank, void **out_comm) {
  RC(need_nccl());
  if (!id128 || !out_comm || world < 1 || rank < 0 || rank >= world) return set_error(MGCMT_ERR_ARG, "bad communicator arguments");
  ncclUniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  ncclComm_t comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, world, id, rank));
  *out_comm = comm;
  return MGCMT_OK;
}

int mgcmt_nccl_comm_destroy(void *comm) {
  if (!comm) return MGCMT_OK;
  RC(need_nccl());
  NC(g_nccl.CommDestroy((ncclComm_t)comm));
  return MGCMT_OK;
}

int mgcmt_slabblock_destroy(mgcmt_slabblock_t *b) {
  if (!b) return MGCMT_OK;
  for (VecState &s : b->vec) {
    if (s.slab) mgcmt_hier_destroy(s.slab);
    if (s.coarse) mgcmt_hier_destroy(s.coarse);
    for (size_t l = 1; l < s.v.size(); ++l) cudaFree(s.v[l]);
    for (size_t l = 1; l < s.f.size(); ++l) cudaFree(s.f[l]);
    for (double *p : s.tmp) cudaFree(p);
    cudaFree(s.fg);
    cudaFree(s.vg);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
  }
  for (cudaEvent_t e : b->prof) cudaEventDestroy(e);
  for (Half &h : b->half) {
    if (h.gs) cudaStreamDestroy(h.gs);
    if (h.fork) cudaEventDestroy(h.fork);
    if (h.comp_done) cudaEventDestroy(h.comp_done);
  }
  if (b->fork) cudaEventDestroy(b->fork);
  cudaFree(b->scal);
  delete b;
  return MGCMT_OK;
}

int mgcmt_slabblock_create(void *comm, void *comm2, int world, int rank, int n, int nlev_slab, int lowest_level, int k,
                           const double *row_lo, const double *row_di, const double *row_up, const double *col_lo,
                           const double *col_di, const double *col_up, double omega, void *stream,
                           mgcmt_slabblock_t **out) {
  if (!out) return set_error(MGCMT_ERR_ARG, "null argument");
  if (world < 1 || rank < 0 || rank >= world || n % world) return set_error(MGCMT_ERR_ARG, "rows must divide evenly among the ranks");
  if (world > 1 && !comm) return set_error(MGCMT_ERR_ARG, "a communicator is needed for more than one rank");
  if (k < 1 || k > 16 || nlev_slab < 1) return set_error(MGCMT_ERR_ARG, "need 1 <= k <= 16 vectors and >= 1 slab level");
  const int own0 = n / world;
  if (own0 % (1 << nlev_slab) || (own0 >> (nlev_slab - 1)) < 2 * kHalo)
    return set_error(MGCMT_ERR_ARG, "slab cuts must stay even, and slabs deeper than the halo, on every slab level");
  mgcmt_slabblock *b = new mgcmt_slabblock();
  b->world = world; b->rank = rank; b->n = n; b->own0 = own0; b->nlev = nlev_slab; b->k = k; b->omega = omega;
  // two staggered halves when there is a communicator for each (or none is needed)
  const bool two = k >= 2 && (world == 1 || comm2 != nullptr);
  b->half[0].begin = 0;
  b->half[0].end = two ? (k + 1) / 2 : k;
  b->half[0].comm = (ncclComm_t)comm;
  b->half[1].begin = b->half[0].end;
  b->half[1].end = k;
  b->half[1].comm = (ncclComm_t)comm2;
  b->vec.resize(k);
  if (const char *e = getenv("MGCMT_SLAB_FUSED_RQ")) b->fused_rq = (e[0] != '0');
  auto bail = [&](int rc) {
    std::string msg = mgcmt_last_error();
    mgcmt_slabblock_destroy(b);
    return set_error(rc, msg);
  };
#define CUB(expr)                                                                                     \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      set_error(MGCMT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
      return bail(MGCMT_ERR_CUDA);                                                                    \
    }                                                                                                 \
  } while (0)
  CUB(cudaEventCreateWithFlags(&b->fork, cudaEventDisableTiming));
  for (Half &h : b->half) {
    CUB(cudaStreamCreateWithFlags(&h.gs, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&h.comp_done, cudaEventDisableTiming));
  }
  CUB(cudaMalloc(&b->scal, sizeof(double) * 64));
  const size_t ng = (size_t)(n >> nlev_slab) * (size_t)(n >> nlev_slab);
  for (VecState &s : b->vec) {
    int rc = mgcmt_hier_create_slab(&s.slab, n, n, rank * own0, own0, nlev_slab, kHalo, row_lo, row_di, row_up, col_lo,
                                    col_di, col_up, stream);
    if (rc == MGCMT_OK)
      rc = mgcmt_hier_create2(&s.coarse, n, n, 1, row_lo, row_di, row_up, col_lo, col_di, col_up, lowest_level, nlev_slab,
                              stream);
    if (rc != MGCMT_OK) return bail(rc);
    s.v.assign(nlev_slab, nullptr);
    s.f.assign(nlev_slab, nullptr);
    s.tmp.assign(nlev_slab, nullptr);
    for (int l = 0; l < nlev_slab; ++l) {
      const size_t bytes = sizeof(double) * level_elems(b, l);
      if (l > 0) {
        CUB(cudaMalloc(&s.v[l], bytes));
        CUB(cudaMalloc(&s.f[l], bytes));
        CUB(cudaMemsetAsync(s.v[l], 0, bytes, (cudaStream_t)stream));
        CUB(cudaMemsetAsync(s.f[l], 0, bytes, (cudaStream_t)stream));
      }
      CUB(cudaMalloc(&s.tmp[l], bytes));
      CUB(cudaMemsetAsync(s.tmp[l], 0, bytes, (cudaStream_t)stream));
    }
    CUB(cudaMalloc(&s.fg, sizeof(double) * ng));
    CUB(cudaMalloc(&s.vg, sizeof(double) * ng));
    CUB(cudaMemsetAsync(s.fg, 0, sizeof(double) * ng, (cudaStream_t)stream));
    CUB(cudaMemsetAsync(s.vg, 0, sizeof(double) * ng, (cudaStream_t)stream));
    CUB(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  CUB(cudaStreamSynchronize((cudaStream_t)stream));
#undef CUB
  *out = b;
  return MGCMT_OK;
}

int mgcmt_slabblock_cycle(mgcmt_slabblock_t *b, const double *h_shifts, double *const *h_f0, double *const *h_v0,
                          double *d_lam, void *stream) {
  if (!b || !h_shifts || !h_f0 || !h_v0) return set_error(MGCMT_ERR_ARG, "null argument");
  if (b->world > 1) RC(need_nccl());
  cudaStream_t main = (cudaStream_t)stream;
  const int k = b->k, nl = b->nlev;
  for (int c = 0; c < k; ++c) {
    if (!h_f0[c] || !h_v0[c] || h_f0[c] == h_v0[c]) return set_error(MGCMT_ERR_ARG, "need distinct non-null f and v arrays");
    b->vec[c].f[0] = h_f0[c];
    b->vec[c].v[0] = h_v0[c];
  }
  Half &A = b->half[0], &B = b->half[1];
  const bool two = B.end > B.begin;
  const std::vector<Phase> phases = build_phases(b, d_lam && !b->fused_rq);
  const int nstage = (int)phases.size();
  (void)nl;
  // both halves start after everything already on the caller's stream
  CU(cudaEventRecord(b->fork, main));
  CU(cudaStreamWaitEvent(A.gs, b->fork, 0));
  if (two) CU(cudaStreamWaitEvent(B.gs, b->fork, 0));
  const bool prof = b->profile && !two;
  if (prof) {
    while ((int)b->prof.size() < 2 * nstage + 1) {
      cudaEvent_t e;
      CU(cudaEventCreate(&e));
      b->prof.push_back(e);
    }
    CU(cudaEventRecord(b->prof[0], A.gs));
  }
  RC(comm_stage(b, A, phases[0]));
  if (prof) CU(cudaEventRecord(b->prof[1], A.gs));
  if (two) RC(comm_stage(b, B, phases[0]));
  for (int i = 0; i < nstage; ++i) {
    // A computes stage i (after B's stage i-1 kernels), then exchanges for stage i+1 while B computes stage i, ...
    if (two && i > 0) CU(cudaStreamWaitEvent(A.gs, B.comp_done, 0));
    RC(comp_stage(b, A, phases[i], h_shifts, d_lam));
    if (prof) CU(cudaEventRecord(b->prof[2 * i + 2], A.gs));
    if (i + 1 < nstage) RC(comm_stage(b, A, phases[i + 1]));
    if (prof && i + 1 < nstage) CU(cudaEventRecord(b->prof[2 * i + 3], A.gs));
    if (two) {
      CU(cudaStreamWaitEvent(B.gs, A.comp_done, 0));
      RC(comp_stage(b, B, phases[i], h_shifts, d_lam));
      if (i + 1 < nstage) RC(comm_stage(b, B, phases[i + 1]));
    }
  }
  // back to the caller's stream
  CU(cudaEventRecord(A.fork, A.gs));
  CU(cudaStreamWaitEvent(main, A.fork, 0));
  if (two) {
    CU(cudaEventRecord(B.fork, B.gs));
    CU(cudaStreamWaitEvent(main, B.fork, 0));
  }
  if (d_lam && b->world > 1) NC(g_nccl.AllReduce(d_lam, d_lam, (size_t)2 * k, kNcclFloat64, kNcclSum, A.comm, main));
  for (VecState &s : b->vec) s.f[0] = s.v[0] = nullptr;
  return MGCMT_OK;
}

int mgcmt_slabblock_set_smoother(mgcmt_slabblock_t *b, int smoother, double omega) {
  if (!b) return set_error(MGCMT_ERR_ARG, "null argument");
  if (smoother != MGCMT_SMOOTH_WJACOBI && smoother != MGCMT_SMOOTH_RBGS)
    return set_error(MGCMT_ERR_ARG, "the slab block smooths with weighted Jacobi or red-black Gauss-Seidel");
  b->smoother = smoother;
  b->omega = omega;
  return MGCMT_OK;
}

int mgcmt_debug_slab_phases(int nlev_slab, int smoother, int with_rq_stage, int *h_kinds, int *h_levels, int capacity) {
  if (nlev_slab < 1 || !h_kinds || !h_levels) return -1;
  const std::vector<Phase> ph = build_phases(nlev_slab, smoother == MGCMT_SMOOTH_RBGS, with_rq_stage != 0);
  if ((int)ph.size() > capacity) return -1;
  for (size_t i = 0; i < ph.size(); ++i) {
    h_kinds[i] = (int)ph[i].kind;
    h_levels[i] = ph[i].level;
  }
  return (int)ph.size();
}

int mgcmt_slabblock_profile(mgcmt_slabblock_t *b, int on) {
  if (!b) return set_error(MGCMT_ERR_ARG, "null argument");
  b->profile = on != 0;
  return MGCMT_OK;
}

int mgcmt_slabblock_profile_read(mgcmt_slabblock_t *b, int with_lam, double *ms_comm, double *ms_comp, int *nstage_out) {
  if (!b || !ms_comm || !ms_comp || !nstage_out) return set_error(MGCMT_ERR_ARG, "null argument");
  const int nstage = (int)build_phases(b, with_lam && !b->fused_rq).size();
  if ((int)b->prof.size() < 2 * nstage + 1) return set_error(MGCMT_ERR_STATE, "no profiled cycle (lock-step form, profile on)");
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < nstage; ++i) {
    float a = 0.f, c = 0.f;
    CU(cudaEventElapsedTime(&a, b->prof[2 * i], b->prof[2 * i + 1]));
    CU(cudaEventElapsedTime(&c, b->prof[2 * i + 1], b->prof[2 * i + 2]));
    ms_comm[i] = a;
    ms_comp[i] = c;
  }
  *nstage_out = nstage;
  return MGCMT_OK;
}

int mgcmt_slabblock_gram(mgcmt_slabblock_t *b, double *d_block, long long stride, void *stream) {
  if (!b || !d_block) return set_error(MGCMT_ERR_ARG, "null argument");
  if (b->world > 1) RC(need_nccl());
  if (b->k > 6) return set_error(MGCMT_ERR_ARG, "Gram-matrix orthonormalisation takes k <= 6");
  cudaStream_t main = (cudaStream_t)stream;
  const long long n_own = (long long)b->own0 * b->n;
  double *base = d_block + (size_t)kHalo * b->n;   // owned rows of vector 0; vector c is `stride` doubles further
  RC(mgcmt_gram(n_own, b->k, base, stride, b->scal, stream));
  if (b->world > 1)
    NC(g_nccl.AllReduce(b->scal, b->scal, (size_t)(b->k * (b->k + 1) / 2), kNcclFloat64, kNcclSum, b->half[0].comm, main));
  return mgcmt_cholqr_apply(n_own, b->k, base, stride, b->scal, stream);
}

}  // extern "C"
