// psum.cu -- C ABI (include/psum.h) over the sm_100a kernels in psum_kernels.cuh.
// Host logic: handle, plan upload, launch sequencing, NCCL amplitude exchange, Lanczos driver.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <sched.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/psum.h"
#include "psum_eig.h"
#include "psum_hostcopy.h"
#define PSUM_SMALL_KERNELS
#include "psum_kernels.cuh"
#include "psum_kpass.h"
#include "psum_plan.h"

using namespace psum;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(PSUM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                  __FILE__, __LINE__);                                                         \
  } while (0)
#define PS_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != PSUM_OK) return _r; \
  } while (0)

extern "C" const char* psum_last_error(void) { return g_err; }
extern "C" int psum_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen (only touched by the sharded entry points)
// ---------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid_t*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;
constexpr int kNcclUint8 = 1;  // ncclUint8

int load_nccl() {
  if (g_nccl.lib) return PSUM_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(PSUM_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOAD(sym)                                                              \
  *(void**)(&g_nccl.sym) = dlsym(lib, "nccl" #sym);                             \
  if (!g_nccl.sym) return fail(PSUM_ERR_NCCL, "libnccl lacks symbol nccl" #sym);
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(AllReduce) LOAD(AllGather)
  LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
  g_nccl.lib = lib;
  return PSUM_OK;
}
// cuMemGetAddressRange (driver API, through dlopen): base of the allocation a device pointer lives in
typedef int (*cu_get_range_t)(unsigned long long*, size_t*, unsigned long long);
cu_get_range_t g_cu_get_range = nullptr;
int load_cuda_driver() {
  if (g_cu_get_range) return PSUM_OK;
  void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(PSUM_ERR_CUDA, "cannot dlopen libcuda.so.1: %s", dlerror());
  *(void**)(&g_cu_get_range) = dlsym(lib, "cuMemGetAddressRange_v2");
  if (!g_cu_get_range) return fail(PSUM_ERR_CUDA, "libcuda lacks cuMemGetAddressRange_v2");
  return PSUM_OK;
}
#define NCCL_TRY(expr)                                                                        \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != 0) return fail(PSUM_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
  } while (0)
}  // namespace

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct DevPass {
  PassParams params;
  bool has_im = false;
  int hi_bit = 0;
  uint64_t free_mask = 0;   // local qubits of the pass's free set
  bool has_remote = false;  // some member reads partners from global memory (x-mask outside the free set)
  int rb = 3;
  int threads = 256;
  int grid = 0;
  size_t smem = 0;
};
struct DevPart {
  int g = 0;
  std::vector<DevPass> passes;
  SimpleGroup* sgroups = nullptr;
  SimpleTerm* sterms = nullptr;
  int nsg = 0, nst = 0;
  int grid_stream = 0;
};

struct psum_handle_s {
  int n = 0, p_bits = 0, rank = 0, world = 1, n_loc = 0;
  std::vector<MergedTerm> terms;
  bool hermitian = true, diagonal = true;
  PlanOptions opt;
  int lanczos_probe = 2;  // 0 off, 1 on, 2 auto (on for <= 2^20 amplitudes per rank)
  int term_cache = 1;     // one-chunk passes keep their component terms in shared memory when they fit
  int herm_pairing = 1;   // <psi|H|psi> of a Hermitian operator: evaluate pairs (i, i^x) once
  std::vector<PartHost> parts;
  bool planned = false;
  // device side
  bool uploaded = false;
  int device = -1, num_sms = 0;
  void* arena = nullptr;
  std::vector<DevPart> dparts;
  double2* d_partial = nullptr;
  size_t partial_cap = 0;
  double2* d_scalar = nullptr;  // 64 complex scalars
  double2* h_scalar = nullptr;  // pinned mirror
  // chunk-pipelined <psi|H|psi> from host memory (psum_expect_host): its own plan whose early
  // passes never combine amplitudes of different chunks
  std::vector<PartHost> parts_h;
  std::vector<DevPart> dparts_h;
  void* arena_h = nullptr;
  int late_bits_h = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_chunk[256] = {nullptr};
  cudaEvent_t ev_free = nullptr;
  // pageable host input: ring of pinned staging buffers filled by host threads ahead of the DMA
  static constexpr int kBounceSlots = 4;
  char* bounce_buf[kBounceSlots] = {nullptr};
  cudaEvent_t bounce_done[kBounceSlots] = {nullptr};
  size_t bounce_bytes = 0;
  int bounce_next = 0;
  int copy_threads = 0;     // 0 = auto
  int host_chunk_bits = 0;  // option "host_chunk_bits": the host state is copied in 2^bits chunks (0 = auto)
  int late_split = 1;       // option "late_split": host-state plan gives every late qubit passes of its own (run_ready)
  psum::CopyPool* copy_pool = nullptr;   // persistent copy threads (created by the first pageable transfer)
  // exchange instrumentation (option "profile_exchange") and the compute-only mode used to measure
  // how much of the exchange hides behind the passes (option "skip_exchange")
  int profile_exchange = 0, skip_exchange = 0;
  // peer-memory exchange: every rank maps its partners' source vector through CUDA IPC and PULLS the
  // partner shard with the copy engines (no SM is taken from the passes); NCCL send/recv is the fallback
  int peer_copy = 1;        // option "peer_copy": 1 = copy engines over IPC-mapped peer memory, 0 = NCCL send/recv
  std::vector<std::pair<std::string, void*>> ipc_open;   // opened peer allocations (handle bytes -> base)
  void* d_ipc = nullptr;    // device scratch of the handle all-gather (world x 80 bytes)
  int comm_sms = 0;         // SMs left to the NCCL send/recv kernels while exchanges are in flight
  int grid_reserve = 0;     // (set by the sharded driver) SMs the pass launches leave free
  std::vector<cudaEvent_t> ex_ev;   // start/stop pairs on comm_stream
  int ex_count = 0;
  double ex_stats[4] = {0, 0, 0, 0};
  // sharding
  nccl_comm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_recv[8] = {nullptr}, ev_used[8] = {nullptr};
};

static uint64_t n_local_amps(const psum_handle_s* h) { return 1ull << h->n_loc; }

static void free_device(psum_handle_s* h) {
  if (h->arena) cudaFree(h->arena);
  if (h->arena_h) cudaFree(h->arena_h);
  h->arena_h = nullptr;
  h->dparts_h.clear();
  h->parts_h.clear();
  if (h->d_partial) cudaFree(h->d_partial);
  if (h->d_scalar) cudaFree(h->d_scalar);
  if (h->h_scalar) cudaFreeHost(h->h_scalar);
  h->arena = nullptr;
  h->d_partial = nullptr;
  h->d_scalar = nullptr;
  h->h_scalar = nullptr;
  h->dparts.clear();
  h->uploaded = false;
}

static int plan(psum_handle_s* h) {
  h->parts = build_parts(h->terms, h->n, h->p_bits, h->rank, h->opt);
  h->planned = true;
  if (h->uploaded) free_device(h);
  return PSUM_OK;
}

static size_t pass_smem_bytes(int L, bool big, int quad_cap) { return pass_smem_bytes_c(L, big, quad_cap); }

template <typename T>
static T* arena_put(char* base, size_t& off, const std::vector<T>& v, std::vector<char>& host) {
  off = (off + 255) & ~(size_t)255;
  size_t bytes = v.size() * sizeof(T);
  if (host.size() < off + bytes) host.resize(off + bytes);
  if (bytes) memcpy(host.data() + off, v.data(), bytes);
  T* p = reinterpret_cast<T*>(base + off);
  off += bytes;
  return p;
}

// Uploads one plan (vector of parts) into its own device arena.
static int upload_plan(psum_handle_s* h, const std::vector<PartHost>& parts, void** arena,
                       std::vector<DevPart>* dparts, size_t* partials_out) {
  std::vector<char> host;
  size_t off = 0;
  size_t max_partials = 0;
  for (int phase = 0; phase < 2; ++phase) {
    char* base = (char*)*arena;
    off = 0;
    host.clear();
    dparts->clear();
    max_partials = 0;
    for (const PartHost& P : parts) {
      DevPart D;
      D.g = P.g;
      D.sgroups = arena_put(base, off, P.sgroups, host);
      D.sterms = arena_put(base, off, P.sterms, host);
      D.nsg = (int)P.sgroups.size();
      D.nst = (int)P.sterms.size();
      uint64_t blocks = (n_local_amps(h) + kThreads - 1) / kThreads;
      D.grid_stream = (int)std::min<uint64_t>(blocks, (uint64_t)h->num_sms * 8);
      size_t partials = D.grid_stream;
      size_t pass_partials = 0;
      for (const PassHost& ph : P.passes) {
        DevPass dp;
        memset(&dp.params, 0, sizeof(dp.params));
        dp.params.bundles = arena_put(base, off, ph.bundles, host);
        dp.params.groups = arena_put(base, off, ph.groups, host);
        dp.params.pat_zt = arena_put(base, off, ph.pat_zt, host);
        dp.params.pat_inl = arena_put(base, off, ph.pat_inl, host);
        dp.params.pat_tptr = arena_put(base, off, ph.pat_tptr, host);
        dp.params.term_zb = arena_put(base, off, ph.term_zb, host);
        dp.params.term_val = arena_put(base, off, ph.term_val, host);
        dp.params.chunks = arena_put(base, off, ph.chunks, host);
        dp.params.nchunks = (int)ph.chunks.size();
        dp.params.L = ph.L;
        dp.params.n_nonfree = h->n_loc - ph.L;
        dp.params.ntiles = 1ull << (h->n_loc - ph.L);
        dp.params.tile_begin = 0;
        dp.params.tile_end = dp.params.ntiles;
        int nj = 0;
        for (int b = 0; b < h->n_loc; ++b)
          if (!(ph.free_mask >> b & 1ull)) dp.params.nbit[nj++] = (uint8_t)b;
        for (int j = 0; j < ph.L; ++j) dp.params.fbit[j] = ph.fbit[j];
        for (int r = 0; r < (1 << ph.rb); ++r) {
          uint64_t o = 0;
          for (int b = 0; b < ph.rb; ++b)
            if (r >> b & 1) o |= 1ull << ph.fbit[kRegBitLo + b];
          dp.params.roff[r] = 16ull * o;
        }
        dp.rb = ph.rb;
        dp.has_im = ph.has_im;
        dp.hi_bit = ph.hi_bit;
        dp.free_mask = ph.free_mask;
        dp.has_remote = ph.n_remote > 0;
        dp.params.quad_idx = arena_put(base, off, ph.quad_idx, host);
        dp.params.batch = 1;
        dp.params.batch_stride = 0;
        // 2^RB amplitudes per thread; 2^13 tiles and "big" passes run one CTA per SM, the rest two.
        // The kernel configuration (threads << RB >= 4096) fixes the static staging capacities.
        const bool big_cfg = ph.L >= 13 || ph.caps.big;
        dp.threads = (big_cfg ? 4096 : 2048) >> ph.rb;
        // quad arena: what the plan needs, at least one table set so that the pointer stays in bounds
        int max_nq = 0;
        for (const ChunkRec& cr : ph.chunks) max_nq = std::max<int>(max_nq, cr.nq);
        dp.params.quad_cap = max_nq * quad_table_doubles(ph.L, ph.rb);
        dp.smem = pass_smem_bytes(ph.L, big_cfg, dp.params.quad_cap);
        // a one-chunk pass keeps its component terms in shared memory when they fit beside the tile
        // (two CTAs per SM: 113 KB each; one big CTA: 227 KB)
        dp.params.term_cache = 0;
        if (h->term_cache && ph.chunks.size() == 1 && !ph.term_val.empty() && !ph.has_im) {   // real-only kernels only
          const size_t extra = 16 * ph.term_val.size();
          const size_t limit = (size_t)(big_cfg ? 227 : 113) * 1024;
          if (dp.smem + extra <= limit) {
            dp.params.term_cache = (int)ph.term_val.size();
            dp.smem += extra;
          }
        }
        int per_sm = (dp.smem > 113 * 1024 || big_cfg) ? 1 : 2;
        dp.grid = (int)std::min<uint64_t>(dp.params.ntiles, (uint64_t)h->num_sms * per_sm);
        pass_partials += dp.grid;
        D.passes.push_back(dp);
      }
      partials = std::max(partials, pass_partials);
      max_partials += partials;
      dparts->push_back(std::move(D));
    }
    if (phase == 0) {
      size_t bytes = std::max<size_t>(off, 256);
      CUDA_TRY(cudaMalloc(arena, bytes));
    } else if (off) {
      CUDA_TRY(cudaMemcpy(*arena, host.data(), off, cudaMemcpyHostToDevice));
    }
  }
  *partials_out = max_partials;
  return PSUM_OK;
}

static int upload(psum_handle_s* h) {
  if (h->uploaded) return PSUM_OK;
  if (!h->planned) PS_TRY(plan(h));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PSUM_ERR_CUDA, "no CUDA device available (%s); libpsum has no CPU fallback",
                cudaGetErrorString(e));
  CUDA_TRY(cudaGetDevice(&h->device));
  CUDA_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  size_t max_partials = 0;
  PS_TRY(upload_plan(h, h->parts, &h->arena, &h->dparts, &max_partials));
  h->partial_cap = std::max<size_t>(max_partials, 4096) * 2 + 128 * 4096;
  CUDA_TRY(cudaMalloc(&h->d_partial, h->partial_cap * sizeof(double2)));
  CUDA_TRY(cudaMalloc(&h->d_scalar, 256 * sizeof(double2)));
  CUDA_TRY(cudaMallocHost(&h->h_scalar, 256 * sizeof(double2)));
  static bool attr_done_dev[64] = {false};  // function attributes are per device
  bool& attr_done = attr_done_dev[h->device >= 0 && h->device < 64 ? h->device : 0];
  if (!attr_done) {
    const int max_smem = 227 * 1024;
    CUDA_TRY(kpass_set_smem_3_256(max_smem));
    CUDA_TRY(kpass_set_smem_3_512(max_smem));
    CUDA_TRY(kpass_set_smem_4_128(max_smem));
    CUDA_TRY(kpass_set_smem_4_256(max_smem));
    attr_done = true;
  }
  h->uploaded = true;
  return PSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// create / destroy / info
// ---------------------------------------------------------------------------------------------
extern "C" int psum_create(int n_qubits, int64_t n_terms, const uint64_t* x_mask,
                           const uint64_t* z_mask, const uint8_t* ypow, const double* coeff_re_im,
                           psum_handle_t* out) {
  if (!out) return fail(PSUM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_qubits < 1 || n_qubits > 64) return fail(PSUM_ERR_INVALID, "n_qubits=%d out of [1,64]", n_qubits);
  if (n_terms < 0 || (n_terms > 0 && (!x_mask || !z_mask || !coeff_re_im)))
    return fail(PSUM_ERR_INVALID, "bad term arrays");
  const uint64_t lm = low_mask(n_qubits);
  for (int64_t k = 0; k < n_terms; ++k)
    if ((x_mask[k] & ~lm) || (z_mask[k] & ~lm))
      return fail(PSUM_ERR_INVALID, "term %lld has mask bits above qubit %d", (long long)k, n_qubits - 1);
  psum_handle_s* h = new psum_handle_s();
  h->n = h->n_loc = n_qubits;
  h->terms = merge_terms(n_terms, x_mask, z_mask, ypow, coeff_re_im);
  double cmax = 0.0;
  for (auto& t : h->terms) cmax = std::max(cmax, std::max(fabs(t.re), fabs(t.im)));
  for (auto& t : h->terms) {
    if (t.x) h->diagonal = false;
    // coefficient of the Hermitian label Pauli = c' * i^{nY}
    int ny = popc64(t.x & t.z) & 3;
    double im = (ny == 0) ? t.im : (ny == 1) ? t.re : (ny == 2) ? -t.im : -t.re;
    if (fabs(im) > 1e-13 * cmax) h->hermitian = false;
  }
  *out = h;
  return PSUM_OK;
}

extern "C" int psum_destroy(psum_handle_t h) {
  if (!h) return PSUM_OK;
  free_device(h);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev_free) cudaEventDestroy(h->ev_free);
  delete h->copy_pool;
  h->copy_pool = nullptr;
  for (int i = 0; i < psum_handle_s::kBounceSlots; ++i) {
    if (h->bounce_buf[i]) cudaFreeHost(h->bounce_buf[i]);
    if (h->bounce_done[i]) cudaEventDestroy(h->bounce_done[i]);
  }
  for (auto& ev : h->ex_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& kv : h->ipc_open)
    if (kv.second) cudaIpcCloseMemHandle(kv.second);
  if (h->d_ipc) cudaFree(h->d_ipc);
  for (auto& ev : h->ev_chunk)
    if (ev) cudaEventDestroy(ev);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  for (int i = 0; i < 8; ++i) {
    if (h->ev_recv[i]) cudaEventDestroy(h->ev_recv[i]);
    if (h->ev_used[i]) cudaEventDestroy(h->ev_used[i]);
  }
  delete h;
  return PSUM_OK;
}

extern "C" int psum_set_option(psum_handle_t h, const char* key, int64_t value) {
  if (!h || !key) return fail(PSUM_ERR_INVALID, "null argument");
  std::string k(key);
  if (k == "tile_bits") {
    if (value < 11 || value > kMaxTileBits) return fail(PSUM_ERR_INVALID, "tile_bits must be 11..13");
    h->opt.tile_bits = (int)value;
  } else if (k == "low_bits") {
    if (value < 1 || value > 8) return fail(PSUM_ERR_INVALID, "low_bits must be 1..8");
    h->opt.low_bits = (int)value;
  } else if (k == "kernel") {
    if (value < 0 || value > 2) return fail(PSUM_ERR_INVALID, "kernel must be 0,1,2");
    h->opt.kernel = (int)value;
  } else if (k == "min_cover") {
    if (value < 1) return fail(PSUM_ERR_INVALID, "min_cover must be >= 1");
    h->opt.min_cover = (int)value;
  } else if (k == "plan_tries") {
    if (value < 1 || value > 4096) return fail(PSUM_ERR_INVALID, "plan_tries must be 1..4096");
    h->opt.plan_tries = (int)value;
  } else if (k == "reg_bits") {
    if (value != 3 && value != 4) return fail(PSUM_ERR_INVALID, "reg_bits must be 3 or 4");
    h->opt.reg_bits = (int)value;
  } else if (k == "share_weight") {
    if (value < 0 || value > 10000) return fail(PSUM_ERR_INVALID, "share_weight must be 0..10000 (percent)");
    h->opt.share_weight = (int)value;
  } else if (k == "min_bundle") {
    if (value < 2 || value > 255) return fail(PSUM_ERR_INVALID, "min_bundle must be 2..255");
    h->opt.min_bundle = (int)value;
  } else if (k == "cta_mode") {
    if (value < 0 || value > 2) return fail(PSUM_ERR_INVALID, "cta_mode must be 0 (auto), 1 (two CTAs per SM) or 2 (one big CTA)");
    h->opt.cta_mode = (int)value;
  } else if (k == "host_chunk_bits") {
    if (value < 0 || value > 6) return fail(PSUM_ERR_INVALID, "host_chunk_bits must be 0 (auto) .. 6");
    h->host_chunk_bits = (int)value;
    h->parts_h.clear();
    return PSUM_OK;
  } else if (k == "late_split") {
    h->late_split = h->opt.late_split = value ? 1 : 0;
    h->parts_h.clear();   // the host-state plan is rebuilt on its next use
  } else if (k == "late_bits") {
    if (value < 0 || value > 8) return fail(PSUM_ERR_INVALID, "late_bits must be 0..8");
    h->opt.late_bits = (int)value;
  } else if (k == "herm_pairing") {
    h->herm_pairing = value ? 1 : 0;
  } else if (k == "term_cache") {
    h->term_cache = value ? 1 : 0;   // shared-memory sizes are fixed when the plan is uploaded: fall through
  } else if (k == "profile_exchange") {
    h->profile_exchange = value ? 1 : 0;
    return PSUM_OK;
  } else if (k == "skip_exchange") {
    h->skip_exchange = value ? 1 : 0;
    return PSUM_OK;
  } else if (k == "peer_copy") {
    h->peer_copy = value ? 1 : 0;
    return PSUM_OK;
  } else if (k == "comm_sms") {
    if (value < 0 || value > 64) return fail(PSUM_ERR_INVALID, "comm_sms must be 0..64");
    h->comm_sms = (int)value;
    return PSUM_OK;
  } else if (k == "copy_threads") {
    if (value < 0 || value > 256) return fail(PSUM_ERR_INVALID, "copy_threads must be 0..256");
    h->copy_threads = (int)value;
    return PSUM_OK;
  } else if (k == "lanczos_probe") {
    if (value < 0 || value > 2) return fail(PSUM_ERR_INVALID, "lanczos_probe must be 0, 1 or 2");
    h->lanczos_probe = (int)value;
    return PSUM_OK;
  } else {
    return fail(PSUM_ERR_INVALID, "unknown option %s", key);
  }
  h->planned = false;
  if (h->uploaded) free_device(h);
  return PSUM_OK;
}

extern "C" int psum_info(psum_handle_t h, int64_t* info, int n_info) {
  if (!h || !info) return fail(PSUM_ERR_INVALID, "null argument");
  if (!h->planned) PS_TRY(plan(h));
  int64_t v[17] = {0};
  int64_t v16 = 0;
  v[0] = h->n;
  v[1] = (int64_t)h->terms.size();
  {
    std::vector<uint64_t> xs;
    for (auto& t : h->terms) xs.push_back(t.x);
    std::sort(xs.begin(), xs.end());
    v[2] = (int64_t)(std::unique(xs.begin(), xs.end()) - xs.begin());
  }
  int64_t gparts = 0;
  for (auto& P : h->parts) {
    v[3] += (int64_t)P.passes.size();
    for (auto& ps : P.passes) {
      v[4] += ps.n_remote;
      v[8] += (int64_t)ps.pat_zt.size();
      v[12] += ps.n_flat;
      v[14] += ps.n_fast;
      v[15] += (int64_t)ps.bundles.size() + ps.n_singles;
      v16 += ps.n_quad;
      v[13] += (int64_t)ps.groups.size();
      v[10] = ps.L;
    }
    v[9] += (int64_t)P.comp.size();
    if (P.g != 0) ++gparts;
  }
  v[5] = h->hermitian;
  v[6] = h->diagonal;
  v[7] = h->n_loc;
  v[11] = gparts;
  v[16] = v16;
  for (int i = 0; i < n_info && i < 17; ++i) info[i] = v[i];
  return PSUM_OK;
}

static PartHost* find_part(psum_handle_s* h, int g) {
  for (auto& P : h->parts)
    if (P.g == g) return &P;
  return nullptr;
}

extern "C" int psum_pass_info(psum_handle_t h, int g, int p, uint64_t* free_mask, int64_t* n_groups,
                              int64_t* n_remote_groups, int64_t* n_patterns) {
  if (!h) return fail(PSUM_ERR_INVALID, "null handle");
  if (!h->planned) PS_TRY(plan(h));
  PartHost* P = find_part(h, g);
  if (!P || p < 0 || p >= (int)P->passes.size()) return fail(PSUM_ERR_INVALID, "no pass %d of part %d", p, g);
  const PassHost& ps = P->passes[p];
  if (free_mask) *free_mask = ps.free_mask;
  if (n_groups) *n_groups = ps.n_xgroups;
  if (n_remote_groups) *n_remote_groups = ps.n_remote;
  if (n_patterns) *n_patterns = (int64_t)ps.pat_zt.size();
  return PSUM_OK;
}

// Lower bounds of one H.psi on this rank, from the plan the kernels execute:
//   out[0] DRAM bytes: a pass reads its source tile once, reads the old y (every pass but the first of
//          the step; none without `store`) and writes y; a remote bundle re-reads one partner vector
//   out[1] shared-memory wavefronts: per warp and round a partner load set (bundle or single) costs
//          2^RB LDS.128 of 4 wavefronts, a fast member 2 record reads, a general member 3 + one per
//          pattern, a bundle header 1; staging the tile writes 16 B per amplitude
//   out[2] group records, out[3] partner load sets
extern "C" int psum_plan_floors(psum_handle_t h, int store, double out[4]) {
  if (!h || !out) return fail(PSUM_ERR_INVALID, "null argument");
  if (!h->planned) PS_TRY(plan(h));
  const double N = ldexp(1.0, h->n_loc);
  double dram = 0.0, wf = 0.0, ngr = 0.0, nsets = 0.0;
  bool first = true;
  for (const PartHost& P : h->parts) {
    if (P.passes.empty()) {   // streaming kernel: one partner read per x-group
      dram += 16.0 * N * ((double)P.sgroups.size() + (store ? (first ? 1.0 : 2.0) : 0.0));
      first = false;
      continue;
    }
    for (const PassHost& ps : P.passes) {
      const double na = (double)(1 << ps.rb);
      const double warp_rounds = N / (32.0 * na);
      const double local_sets = (double)ps.n_singles + (double)ps.bundles.size() - (double)ps.n_remote_bundles;
      dram += 16.0 * N * (1.0 + (store ? (first ? 1.0 : 2.0) : 0.0) + (double)ps.n_remote_bundles);
      wf += warp_rounds * (4.0 * na * local_sets + 2.0 * (double)ps.n_fast + 3.0 * (double)ps.n_general +
                           (double)ps.n_general_patterns + (double)ps.bundles.size());
      {  // quadratic member: 2 record reads, (1 + nW + RB) lane-table reads of 2 wavefronts, the rest broadcast
        const int nW = ps.L - kRegBitLo - ps.rb;
        wf += warp_rounds * (double)ps.n_quad * (2.0 + 2.0 * (1 + nW + ps.rb) + 1.0 + ps.rb + 0.5 * ps.rb * (ps.rb - 1));
      }
      wf += N * 16.0 / 128.0;
      ngr += (double)ps.groups.size();
      nsets += (double)ps.n_singles + (double)ps.bundles.size();
      first = false;
    }
  }
  out[0] = dram;
  out[1] = wf;
  out[2] = ngr;
  out[3] = nsets;
  return PSUM_OK;
}

// Plan introspection (host only): hands out the device-facing arrays of one pass exactly as the
// kernels read them, so that the planner can be checked without a device (tests/plan_interpreter.py).
extern "C" int psum_plan_export(psum_handle_t h, int g, int p, int64_t counts[12], void* groups,
                                uint16_t* pat_zt, uint16_t* pat_inl, uint32_t* pat_tptr, uint64_t* term_zb,
                                double* term_val, void* chunks, uint8_t fbit[16], uint8_t nbit[64], void* bundles,
                                uint16_t* quad_idx) {
  if (!h || !counts) return fail(PSUM_ERR_INVALID, "null argument");
  if (!h->planned) PS_TRY(plan(h));
  PartHost* P = find_part(h, g);
  if (!P || p < 0 || p >= (int)P->passes.size()) return fail(PSUM_ERR_INVALID, "no pass %d of part %d", p, g);
  const PassHost& ps = P->passes[p];
  counts[0] = (int64_t)ps.groups.size();
  counts[1] = (int64_t)ps.pat_zt.size();
  counts[2] = (int64_t)ps.term_val.size();
  counts[3] = (int64_t)ps.chunks.size();
  counts[4] = ps.L;
  counts[5] = h->n_loc - ps.L;
  counts[6] = ps.has_im ? 1 : 0;
  counts[7] = ps.hi_bit;
  counts[8] = (int64_t)ps.bundles.size();
  counts[9] = ps.rb;
  counts[10] = (int64_t)ps.quad_idx.size();
  counts[11] = ps.caps.big ? 1 : 0;
  if (quad_idx) memcpy(quad_idx, ps.quad_idx.data(), ps.quad_idx.size() * sizeof(uint16_t));
  if (groups) memcpy(groups, ps.groups.data(), ps.groups.size() * sizeof(GroupRec));
  if (bundles) memcpy(bundles, ps.bundles.data(), ps.bundles.size() * sizeof(BundleRec));
  if (pat_zt) memcpy(pat_zt, ps.pat_zt.data(), ps.pat_zt.size() * sizeof(uint16_t));
  if (pat_inl) memcpy(pat_inl, ps.pat_inl.data(), ps.pat_inl.size() * sizeof(uint16_t));
  if (pat_tptr) memcpy(pat_tptr, ps.pat_tptr.data(), ps.pat_tptr.size() * sizeof(uint32_t));
  if (term_zb) memcpy(term_zb, ps.term_zb.data(), ps.term_zb.size() * sizeof(uint64_t));
  if (term_val) memcpy(term_val, ps.term_val.data(), ps.term_val.size() * sizeof(double));
  if (chunks) memcpy(chunks, ps.chunks.data(), ps.chunks.size() * sizeof(ChunkRec));
  if (fbit) memcpy(fbit, ps.fbit, 16);
  if (nbit) {
    int nj = 0;
    for (int b = 0; b < h->n_loc; ++b)
      if (!(ps.free_mask >> b & 1ull)) nbit[nj++] = (uint8_t)b;
  }
  return PSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// launch sequencing
// ---------------------------------------------------------------------------------------------
// Runs one part.  With a y buffer every pass accumulates into y and only the launch flagged
// `last` reduces <bra|y> over the FINAL y; without y each pass appends the partials of its own
// contribution at d_partial[*pcount...].
// pass_lo/pass_hi and chunk/nchunks restrict the launch to a pass range and to the tiles of one
// 2^(n_loc - log2 nchunks)-amplitude chunk (only valid for passes whose hi_bit lies inside a chunk).
static int run_part(psum_handle_s* h, const DevPart& D, const double2* src, const double2* bra,
                    double2* y, int accumulate, bool want_expect, bool last_part, size_t* pcount,
                    cudaStream_t st, size_t pass_lo = 0, size_t pass_hi = (size_t)-1, uint64_t chunk = 0,
                    uint64_t nchunks = 1) {
  const bool bra_is_src = (bra == src) || bra == nullptr;
  if (!D.passes.empty()) {
    for (size_t pi = pass_lo; pi < std::min(pass_hi, D.passes.size()); ++pi) {
      const DevPass& dp = D.passes[pi];
      PassParams pp = dp.params;
      if (nchunks > 1) {
        const uint64_t per = pp.ntiles / nchunks;
        pp.tile_begin = chunk * per;
        pp.tile_end = pp.tile_begin + per;
      }
      pp.accumulate = (accumulate || pi > 0) ? 1 : 0;
      // the persistent grid fills every SM; while an exchange is in flight a few SMs are left to the
      // NCCL kernels, otherwise CTAs that cannot become resident wait for a whole pass
      int grid = dp.grid;
      if (h->grid_reserve > 0) {
        const int per_sm = std::max(1, dp.grid / std::max(1, h->num_sms));
        grid = std::max(per_sm, std::min(dp.grid, (h->num_sms - h->grid_reserve) * per_sm));
      }
      const bool last = last_part && (pi + 1 == D.passes.size());
      const bool ex = want_expect && (y ? last : true);
      double2* part = ex ? h->d_partial + *pcount : nullptr;
      const double2* b = bra_is_src ? nullptr : bra;
      if (ex && *pcount + grid > h->partial_cap)
        return fail(PSUM_ERR_NOMEM, "expectation partial buffer too small");
      // Hermitian pairing: expectation only, <psi|H|psi> of a Hermitian operator, local part
      const bool herm = !y && h->hermitian && h->herm_pairing && bra_is_src && D.g == 0;
      const int mode = (y ? kModeStore : 0) | (ex ? kModeExpect : 0) | (dp.has_im ? kModeIm : 0) |
                       ((herm && ex) ? kModeHerm : 0);
      const cudaError_t le = kpass_launch(dp.rb, dp.threads, mode, grid, dp.smem, st, src, b, y, pp, part);
      if (le != cudaSuccess)
        return fail(PSUM_ERR_CUDA, "k_pass launch (mode %d, rb %d, %d threads) failed: %s", mode, dp.rb, dp.threads,
                    cudaGetErrorString(le));
      if (ex) *pcount += grid;
    }
  } else {
    const bool ex = want_expect && (y ? last_part : true);
    double2* part = ex ? h->d_partial + *pcount : nullptr;
    if (ex && *pcount + D.grid_stream > h->partial_cap)
      return fail(PSUM_ERR_NOMEM, "expectation partial buffer too small");
    k_stream<<<D.grid_stream, kThreads, 0, st>>>(src, bra ? bra : src, y, D.sgroups, D.nsg, D.sterms,
                                                 n_local_amps(h), accumulate, part);
    if (ex) *pcount += D.grid_stream;
  }
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

static int fetch_scalars(psum_handle_s* h, int count, double* out, cudaStream_t st) {
  CUDA_TRY(cudaMemcpyAsync(h->h_scalar, h->d_scalar, sizeof(double2) * count, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i < count; ++i) {
    out[2 * i] = h->h_scalar[i].x;
    out[2 * i + 1] = h->h_scalar[i].y;
  }
  return PSUM_OK;
}

static int apply_local(psum_handle_s* h, const double2* psi, const double2* bra, double2* y,
                       double* expect_out, cudaStream_t st) {
  if (h->world != 1) return fail(PSUM_ERR_INVALID, "handle is sharded: use psum_apply_sharded / psum_apply_part");
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  if (!psi || (!y && !expect_out)) return fail(PSUM_ERR_INVALID, "null device pointer");
  if (y == psi) return fail(PSUM_ERR_INVALID, "y must not alias psi");
  if (((uintptr_t)psi | (uintptr_t)y | (uintptr_t)bra) & 15u)
    return fail(PSUM_ERR_INVALID, "state vectors must be 16-byte aligned");
  PS_TRY(upload(h));
  size_t pcount = 0;
  for (const DevPart& D : h->dparts)
    PS_TRY(run_part(h, D, psi, bra, y, 0, expect_out != nullptr, &D == &h->dparts.back(), &pcount, st));
  if (expect_out) {
    k_reduce_partials<<<1, kThreads, 0, st>>>(h->d_partial, (int)pcount, h->d_scalar, 0);
    CUDA_TRY(cudaGetLastError());
    PS_TRY(fetch_scalars(h, 1, expect_out, st));
  }
  return PSUM_OK;
}

extern "C" int psum_apply(psum_handle_t h, const void* psi_dev, void* y_dev, double* expect_out,
                          psum_stream_t stream) {
  if (!h) return fail(PSUM_ERR_INVALID, "null handle");
  if (!y_dev) return fail(PSUM_ERR_INVALID, "y_dev is NULL");
  return apply_local(h, (const double2*)psi_dev, (const double2*)psi_dev, (double2*)y_dev, expect_out,
                     (cudaStream_t)stream);
}

extern "C" int psum_expect(psum_handle_t h, const void* psi_dev, double out[2], psum_stream_t stream) {
  if (!h || !out) return fail(PSUM_ERR_INVALID, "null argument");
  return apply_local(h, (const double2*)psi_dev, (const double2*)psi_dev, nullptr, out, (cudaStream_t)stream);
}

static int ensure_host_plan(psum_handle_s* h, int late);
static int host_late_bits(const psum_handle_s* h);

// ---- host -> device copies of caller memory ------------------------------------------------------
// The reference API hands over a pageable numpy array.  cudaMemcpyAsync from pageable memory is
// staged by the driver one piece at a time and blocks the calling thread, which would serialise
// the chunk pipeline; instead host threads copy ahead into a ring of pinned buffers and the DMA
// engine drains them (one cudaMemcpyAsync per slot).
static bool host_is_pageable(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

static int copy_thread_count(const psum_handle_s* h) {
  if (h->copy_threads > 0) return h->copy_threads;
  cpu_set_t set;
  int cores = 8;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
  // ranks of one node share the cores; one core stays free for the thread that feeds the GPU
  const int share = cores / std::max(1, h->world);
  // measured on a 16-core B200 host (30 qubits, pageable numpy): 8 threads 359 ms, 12 threads 370 ms, 15 threads 420 ms
  return std::max(2, std::min(8, share > 4 ? share - 2 : share));
}

// dst_dev[0..bytes) <- src_host (pageable) through the pinned ring, DMA on stream cs
static int bounce_copy(psum_handle_s* h, const char* src, char* dst, size_t bytes, cudaStream_t cs) {
  const size_t slot_bytes = (size_t)128 << 20;
  if (!h->bounce_buf[0]) {
    for (int i = 0; i < psum_handle_s::kBounceSlots; ++i) {
      CUDA_TRY(cudaHostAlloc((void**)&h->bounce_buf[i], slot_bytes, cudaHostAllocDefault));
      CUDA_TRY(cudaEventCreateWithFlags(&h->bounce_done[i], cudaEventDisableTiming));
    }
    h->bounce_bytes = slot_bytes;
  }
  const int nthreads = copy_thread_count(h);
  if (!h->copy_pool || h->copy_pool->threads() != nthreads) {
    delete h->copy_pool;
    h->copy_pool = new psum::CopyPool(nthreads);
  }
  for (size_t off = 0; off < bytes; off += slot_bytes) {
    const size_t nb = std::min(slot_bytes, bytes - off);
    const int s = h->bounce_next;
    h->bounce_next = (s + 1) % psum_handle_s::kBounceSlots;
    CUDA_TRY(cudaEventSynchronize(h->bounce_done[s]));   // the DMA that last read this slot is done
    h->copy_pool->copy(h->bounce_buf[s], src + off, nb);   // streaming stores, see psum_hostcopy.h
    CUDA_TRY(cudaMemcpyAsync(dst + off, h->bounce_buf[s], nb, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaEventRecord(h->bounce_done[s], cs));
  }
  return PSUM_OK;
}

// Once copies FROM the caller's host buffer are queued, an error return must not leave them in
// flight (the caller may free the buffer): the guard drains the copy stream unless disarmed.
struct CopyGuard {
  cudaStream_t cs = nullptr;
  ~CopyGuard() {
    if (cs) cudaStreamSynchronize(cs);
  }
};

// Host-state pipeline: the state arrives in 2^late chunks, in index order (chunk = the top `late` local
// qubits).  A pass whose free set holds the late qubits T combines, in one tile, the 2^|T| chunks that
// differ in T; the tiles with given values of the OTHER late qubits form one contiguous tile range (the
// non-free late qubits are the top bits of the tile id), complete as soon as its last chunk -- the one
// with every bit of T set -- has landed.  After chunk k this launches every such range: chunk-local
// passes (T = 0) run on chunk k itself, a pass that frees only late qubit t runs its range when the
// upper chunk of the pair arrives, and only passes that free ALL late qubits (or read partners from
// global memory) wait for the whole state.  Expectation only (no y): every launch adds its own partials.
static int run_ready(psum_handle_s* h, const DevPart& D, const double2* psi, uint64_t k, int late, int cbits,
                     size_t* pcount, cudaStream_t st) {
  const uint64_t lmask = (1ull << late) - 1;
  for (size_t pi = 0; pi < D.passes.size(); ++pi) {
    const DevPass& dp = D.passes[pi];
    const uint64_t T = dp.has_remote ? lmask : ((dp.free_mask >> cbits) & lmask);
    if ((k & T) != T) continue;
    uint64_t u = 0;
    int j = 0;
    for (int b = 0; b < late; ++b)
      if (!(T >> b & 1ull)) u |= ((k >> b) & 1ull) << j++;
    PS_TRY(run_part(h, D, psi, psi, nullptr, 0, true, true, pcount, st, pi, pi + 1, u, 1ull << j));
  }
  return PSUM_OK;
}

// <psi|H|psi> for a state that lives in HOST memory: the H2D copy is cut into 2^late_bits chunks
// on a copy stream and every "early" pass (hi_bit inside a chunk) runs chunk by chunk as the data
// arrives; only the passes that combine different chunks wait for the whole state.
extern "C" int psum_expect_host(psum_handle_t h, const void* psi_host, void* psi_dev, double out[2],
                                psum_stream_t stream) {
  if (!h || !psi_host || !psi_dev || !out) return fail(PSUM_ERR_INVALID, "null argument");
  if (h->world != 1) return fail(PSUM_ERR_INVALID, "handle is sharded");
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  if ((uintptr_t)psi_dev & 15u) return fail(PSUM_ERR_INVALID, "state vectors must be 16-byte aligned");
  PS_TRY(upload(h));
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t N = n_local_amps(h);
  const size_t bytes = N * sizeof(double2);
  const int late = host_late_bits(h);
  if (late <= 0) {  // small state: one copy, then the ordinary expectation
    CUDA_TRY(cudaMemcpyAsync(psi_dev, psi_host, bytes, cudaMemcpyHostToDevice, st));
    return apply_local(h, (const double2*)psi_dev, (const double2*)psi_dev, nullptr, out, st);
  }
  PS_TRY(ensure_host_plan(h, late));
  if (!h->copy_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_free, cudaEventDisableTiming));
  }
  const uint64_t nchunks = 1ull << late;
  for (uint64_t k = 0; k < nchunks; ++k)
    if (!h->ev_chunk[k]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_chunk[k], cudaEventDisableTiming));
  // the copy may only start once earlier work on `stream` has released psi_dev
  CUDA_TRY(cudaEventRecord(h->ev_free, st));
  CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_free, 0));
  const size_t cbytes = bytes >> late;
  if (h->dparts_h.size() != 1) return fail(PSUM_ERR_INVALID, "unexpected plan shape");
  const DevPart& D = h->dparts_h[0];
  const int cbits = h->n_loc - late;
  const double2* psi = (const double2*)psi_dev;
  const bool pageable = host_is_pageable(psi_host);
  size_t pcount = 0;
  CopyGuard guard;
  guard.cs = h->copy_stream;
  for (uint64_t k = 0; k < nchunks; ++k) {
    char* dst = (char*)psi_dev + k * cbytes;
    const char* srcp = (const char*)psi_host + k * cbytes;
    if (pageable) PS_TRY(bounce_copy(h, srcp, dst, cbytes, h->copy_stream));
    else CUDA_TRY(cudaMemcpyAsync(dst, srcp, cbytes, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev_chunk[k], h->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chunk[k], 0));
    PS_TRY(run_ready(h, D, psi, k, late, cbits, &pcount, st));
  }
  k_reduce_partials<<<1, kThreads, 0, st>>>(h->d_partial, (int)pcount, h->d_scalar, 0);
  CUDA_TRY(cudaGetLastError());
  PS_TRY(fetch_scalars(h, 1, out, st));   // st waited for every chunk: the copies are complete
  guard.cs = nullptr;
  return PSUM_OK;
}

extern "C" int psum_bra_apply(psum_handle_t h, const void* phi_dev, const void* psi_dev, double out[2],
                              psum_stream_t stream) {
  if (!h || !out || !phi_dev) return fail(PSUM_ERR_INVALID, "null argument");
  return apply_local(h, (const double2*)psi_dev, (const double2*)phi_dev, nullptr, out, (cudaStream_t)stream);
}

// Launches one pass for `nb` states stored back to back (expectation only, bra = source): the CTA
// stages the plan of a tile once and loops over the states.  partial[b * grid + block].
static int launch_pass_batch(psum_handle_s* h, const DevPart& D, const DevPass& dp, const double2* src, int nb,
                             uint64_t stride, double2* part, cudaStream_t st) {
  PassParams pp = dp.params;
  pp.accumulate = 0;
  pp.batch = nb;
  pp.batch_stride = stride;
  pp.term_cache = 0;   // the batched launch sizes its own shared memory (per-warp accumulators instead)
  const bool herm = h->hermitian && h->herm_pairing && D.g == 0;
  const int mode = kModeExpect | (dp.has_im ? kModeIm : 0) | (herm ? kModeHerm : 0);
  const bool big_cfg = (dp.threads << dp.rb) >= 4096;
  const size_t smem = pass_smem_bytes_c(pp.L, big_cfg, pp.quad_cap, nb);
  const cudaError_t le = kpass_launch(dp.rb, dp.threads, mode, dp.grid, smem, st, src, nullptr, nullptr, pp, part);
  if (le != cudaSuccess) return fail(PSUM_ERR_CUDA, "batched k_pass launch failed: %s", cudaGetErrorString(le));
  return PSUM_OK;
}

// <psi_b|H|psi_b> for `batch` states stored back to back (VQE max_evals_grouped, vqe.py:518; ListOp
// fronts, matrix_op.py:199-201).  One-chunk passes are launched ONCE for up to kMaxBatch states: each
// CTA folds the plan of a tile (and builds the quadratic-group tables) once and applies it to every
// state; the per-state sums are reduced on the device and read back with one synchronisation per
// group of states.
extern "C" int psum_expect_batch(psum_handle_t h, int batch, const void* psi_dev, double* out,
                                 psum_stream_t stream) {
  if (!h || !out || batch < 0) return fail(PSUM_ERR_INVALID, "bad argument");
  if (h->world != 1) return fail(PSUM_ERR_INVALID, "handle is sharded");
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  if (batch > 0 && (!psi_dev || ((uintptr_t)psi_dev & 15u))) return fail(PSUM_ERR_INVALID, "bad state pointer");
  PS_TRY(upload(h));
  cudaStream_t st = (cudaStream_t)stream;
  const double2* base = (const double2*)psi_dev;
  const uint64_t N = n_local_amps(h);
  size_t max_grid = 1;
  for (const DevPart& D : h->dparts) {
    max_grid = std::max<size_t>(max_grid, (size_t)D.grid_stream);
    for (const DevPass& dp : D.passes) max_grid = std::max<size_t>(max_grid, (size_t)dp.grid);
  }
  if ((size_t)kMaxBatch * max_grid > h->partial_cap) {
    CUDA_TRY(cudaDeviceSynchronize());
    if (h->d_partial) cudaFree(h->d_partial);
    h->d_partial = nullptr;
    h->partial_cap = (size_t)kMaxBatch * max_grid + 4096;
    CUDA_TRY(cudaMalloc(&h->d_partial, h->partial_cap * sizeof(double2)));
  }
  for (int b0 = 0; b0 < batch; b0 += kMaxBatch) {
    const int nb = std::min(kMaxBatch, batch - b0);
    const double2* src0 = base + (uint64_t)b0 * N;
    bool first = true;     // first contribution to d_scalar[0 .. nb)
    for (const DevPart& D : h->dparts) {
      if (D.passes.empty()) {   // streaming kernel: state by state
        for (int b = 0; b < nb; ++b) {
          k_stream<<<D.grid_stream, kThreads, 0, st>>>(src0 + (uint64_t)b * N, src0 + (uint64_t)b * N, nullptr, D.sgroups,
                                                       D.nsg, D.sterms, N, 0, h->d_partial + (size_t)b * D.grid_stream);
        }
        k_accum_batch<<<nb, kThreads, 0, st>>>(h->d_partial, D.grid_stream, h->d_scalar, first ? 0 : 1);
        first = false;
        continue;
      }
      for (const DevPass& dp : D.passes) {
        if (dp.params.nchunks == 1 && nb > 1) {
          PS_TRY(launch_pass_batch(h, D, dp, src0, nb, N, h->d_partial, st));
        } else {                  // several chunks (or a single state): one launch per state
          for (int b = 0; b < nb; ++b)
            PS_TRY(launch_pass_batch(h, D, dp, src0 + (uint64_t)b * N, 1, N, h->d_partial + (size_t)b * dp.grid, st));
        }
        k_accum_batch<<<nb, kThreads, 0, st>>>(h->d_partial, dp.grid, h->d_scalar, first ? 0 : 1);
        first = false;
      }
    }
    CUDA_TRY(cudaGetLastError());
    PS_TRY(fetch_scalars(h, nb, out + 2 * b0, st));
  }
  return PSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// diagonal operators
// ---------------------------------------------------------------------------------------------
extern "C" int psum_diag(psum_handle_t h, double* d_dev, psum_stream_t stream) {
  if (!h || !d_dev) return fail(PSUM_ERR_INVALID, "null argument");
  if (!h->diagonal) return fail(PSUM_ERR_NOT_DIAGONAL, "operator has X/Y terms");
  if (!h->hermitian)
    return fail(PSUM_ERR_NOT_HERMITIAN, "psum_diag returns a real diagonal: the operator has complex weights "
                "(apply it to the all-ones vector for the complex diagonal)");
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  PS_TRY(upload(h));
  const DevPart* D0 = nullptr;
  for (auto& D : h->dparts)
    if (D.g == 0) D0 = &D;
  if (!D0) return fail(PSUM_ERR_INVALID, "no local part");
  // the streaming plan of a diagonal operator is one group holding every term (z local part);
  // the rank sign of the global z-part is already folded into the coefficients.
  k_diag<<<D0->grid_stream, kThreads, 0, (cudaStream_t)stream>>>(d_dev, D0->sterms, D0->nst,
                                                                n_local_amps(h), 0);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// Dense matrix of the operator on the device (row-major 2^n x 2^n complex128): the matrix the
// reference builds with to_matrix() (operator_base / list_op.py:320-329).  Small n only.
extern "C" int psum_to_dense(psum_handle_t h, void* mat_dev, psum_stream_t stream) {
  if (!h || !mat_dev) return fail(PSUM_ERR_INVALID, "null argument");
  if (h->world != 1) return fail(PSUM_ERR_INVALID, "handle is sharded");
  if (h->n > 14) return fail(PSUM_ERR_INVALID, "dense matrix of %d qubits refused (2^%d entries)", h->n, 2 * h->n);
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  PS_TRY(upload(h));
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t N = n_local_amps(h);
  CUDA_TRY(cudaMemsetAsync(mat_dev, 0, N * N * sizeof(double2), st));
  for (const DevPart& D : h->dparts) {
    if (D.g != 0 || D.nsg == 0) continue;
    const uint64_t total = N * (uint64_t)D.nsg;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((total + kThreads - 1) / kThreads, (uint64_t)h->num_sms * 16));
    k_to_dense<<<grid, kThreads, 0, st>>>((double2*)mat_dev, D.sgroups, D.nsg, D.sterms, N);
  }
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// Inverse transform: dense matrix -> Pauli weights (legacy/op_converter.py:35-100 computes
// tr(P M) / 2^n with 4^n sparse products).  mat_dev and w_dev are 2^n x 2^n complex128 row-major on the
// device; w_dev[x][z] = weight of the label Pauli with masks (x, z) times (-i)^{popc(x & z)}.
extern "C" int psum_pauli_decompose(const void* mat_dev, int n_qubits, void* w_dev, psum_stream_t stream) {
  if (!mat_dev || !w_dev) return fail(PSUM_ERR_INVALID, "null argument");
  if (n_qubits < 1 || n_qubits > 12) return fail(PSUM_ERR_INVALID, "pauli_decompose: n_qubits must be 1..12");
  const size_t smem = sizeof(double2) << n_qubits;
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(k_pauli_decompose, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_done = true;
  }
  int dev = 0, sms = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)std::min<uint64_t>(1ull << n_qubits, (uint64_t)sms * 2);
  k_pauli_decompose<<<grid, kThreads, smem, (cudaStream_t)stream>>>((const double2*)mat_dev, (double2*)w_dev, n_qubits);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// <psi|H_j|psi> for several operators (aux operators: numpy_eigen_solver.py:208-226; qEOM commutator
// lists) against ONE state: the operators' passes run back to back on the stream and the host reads
// the n_ops scalars with a single synchronisation.  out = 2 * n_ops doubles.
extern "C" int psum_expect_multi(const psum_handle_t* hs, int n_ops, const void* psi_dev, double* out,
                                 psum_stream_t stream) {
  if (!hs || !out || n_ops < 1 || n_ops > 256 || !psi_dev) return fail(PSUM_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const double2* psi = (const double2*)psi_dev;
  psum_handle_s* first = nullptr;
  for (int j = 0; j < n_ops; ++j) {
    psum_handle_s* h = hs[j];
    if (!h) return fail(PSUM_ERR_INVALID, "null handle %d", j);
    if (h->world != 1) return fail(PSUM_ERR_INVALID, "handle %d is sharded", j);
    if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
    if (first && h->n != first->n) return fail(PSUM_ERR_INVALID, "operators act on different numbers of qubits");
    PS_TRY(upload(h));
    if (!first) first = h;
    size_t pcount = 0;
    for (const DevPart& D : h->dparts) PS_TRY(run_part(h, D, psi, psi, nullptr, 0, true, true, &pcount, st));
    k_reduce_partials<<<1, kThreads, 0, st>>>(h->d_partial, (int)pcount, first->d_scalar + j, 0);
    CUDA_TRY(cudaGetLastError());
  }
  return fetch_scalars(first, n_ops, out, st);
}

extern "C" int psum_diag_kmin(psum_handle_t h, int k, double* vals, uint64_t* idx, psum_stream_t stream) {
  if (!h || !vals || !idx || k < 1) return fail(PSUM_ERR_INVALID, "bad argument");
  if (h->world != 1) return fail(PSUM_ERR_INVALID, "diag_kmin on a sharded handle: reduce per-rank results on the host");
  const uint64_t N = n_local_amps(h);
  if ((uint64_t)k > N) return fail(PSUM_ERR_INVALID, "k > 2^n");
  cudaStream_t st = (cudaStream_t)stream;
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, N * sizeof(double)));
  int rc = psum_diag(h, d, stream);
  if (rc != PSUM_OK) {
    cudaFree(d);
    return rc;
  }
  const int grid = (int)std::min<uint64_t>((N + kThreads - 1) / kThreads, (uint64_t)h->num_sms * 8);
  ValIdx* part = nullptr;
  ValIdx* hpart = (ValIdx*)malloc(sizeof(ValIdx) * grid);
  cudaError_t e = cudaMalloc(&part, sizeof(ValIdx) * grid);
  if (e != cudaSuccess) {
    cudaFree(d);
    free(hpart);
    return fail(PSUM_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  ValIdx lo{0.0, 0};
  for (int j = 0; j < k; ++j) {
    k_argmin_after<<<grid, kThreads, 0, st>>>(d, N, lo, j > 0, part);
    cudaMemcpyAsync(hpart, part, sizeof(ValIdx) * grid, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) break;
    ValIdx best = hpart[0];
    for (int b = 1; b < grid; ++b)
      if (hpart[b].v < best.v || (hpart[b].v == best.v && hpart[b].i < best.i)) best = hpart[b];
    vals[j] = best.v;
    idx[j] = best.i;
    lo = best;
  }
  cudaFree(part);
  cudaFree(d);
  free(hpart);
  if (e != cudaSuccess) return fail(PSUM_ERR_CUDA, "diag_kmin: %s", cudaGetErrorString(e));
  return PSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// vector helpers
// ---------------------------------------------------------------------------------------------
static int vec_grid(uint64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + kThreads - 1) / kThreads, (uint64_t)sms * 8));
}

extern "C" int psum_fill_state(void* psi_dev, uint64_t n_amps, uint64_t first_index, uint64_t seed,
                               psum_stream_t stream) {
  if (!psi_dev) return fail(PSUM_ERR_INVALID, "null pointer");
  k_fill_state<<<vec_grid(n_amps), kThreads, 0, (cudaStream_t)stream>>>((double2*)psi_dev, n_amps, first_index, seed);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// scratch for the handle-less vector helpers
struct VecScratch {
  double2* d = nullptr;
  double2* h = nullptr;
  int cap = 0;
};
static thread_local VecScratch g_vs;
static int vec_scratch(int need) {
  if (g_vs.cap >= need) return PSUM_OK;
  if (g_vs.d) cudaFree(g_vs.d);
  if (g_vs.h) cudaFreeHost(g_vs.h);
  g_vs.cap = 0;
  CUDA_TRY(cudaMalloc(&g_vs.d, sizeof(double2) * (need + 1)));
  CUDA_TRY(cudaMallocHost(&g_vs.h, sizeof(double2) * 4));
  g_vs.cap = need;
  return PSUM_OK;
}

extern "C" int psum_vec_dot(const void* a_dev, const void* b_dev, uint64_t n_amps, double out[2],
                            psum_stream_t stream) {
  if (!a_dev || !b_dev || !out) return fail(PSUM_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = vec_grid(n_amps);
  PS_TRY(vec_scratch(grid));
  k_dot<<<grid, kThreads, 0, st>>>((const double2*)a_dev, (const double2*)b_dev, n_amps, g_vs.d);
  k_reduce_partials<<<1, kThreads, 0, st>>>(g_vs.d, grid, g_vs.d + grid, 0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(g_vs.h, g_vs.d + grid, sizeof(double2), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  out[0] = g_vs.h[0].x;
  out[1] = g_vs.h[0].y;
  return PSUM_OK;
}

extern "C" int psum_vec_scale(void* a_dev, uint64_t n_amps, double s_re, double s_im, psum_stream_t stream) {
  if (!a_dev) return fail(PSUM_ERR_INVALID, "null pointer");
  k_scale<<<vec_grid(n_amps), kThreads, 0, (cudaStream_t)stream>>>((double2*)a_dev, n_amps, s_re, s_im);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

extern "C" int psum_vec_axpy(void* y_dev, const void* x_dev, uint64_t n_amps, double a_re, double a_im,
                             psum_stream_t stream) {
  if (!y_dev || !x_dev) return fail(PSUM_ERR_INVALID, "null pointer");
  k_axpy<<<vec_grid(n_amps), kThreads, 0, (cudaStream_t)stream>>>((double2*)y_dev, (const double2*)x_dev, n_amps, a_re, a_im);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// gate-level state preparation (single GPU)
// ---------------------------------------------------------------------------------------------
extern "C" int psum_state_basis(void* psi_dev, int n_qubits, uint64_t index, psum_stream_t stream) {
  if (!psi_dev || n_qubits < 1 || n_qubits > 40) return fail(PSUM_ERR_INVALID, "bad argument");
  const uint64_t N = 1ull << n_qubits;
  if (index >= N) return fail(PSUM_ERR_INVALID, "basis index out of range");
  k_set_basis_state<<<vec_grid(N), kThreads, 0, (cudaStream_t)stream>>>((double2*)psi_dev, N, index);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

extern "C" int psum_gate_1q(void* psi_dev, int n_qubits, int q, const double* u_re_im, psum_stream_t stream) {
  if (!psi_dev || !u_re_im || n_qubits < 1 || q < 0 || q >= n_qubits) return fail(PSUM_ERR_INVALID, "bad argument");
  Gate2x2 U;
  U.u00 = make_double2(u_re_im[0], u_re_im[1]);
  U.u01 = make_double2(u_re_im[2], u_re_im[3]);
  U.u10 = make_double2(u_re_im[4], u_re_im[5]);
  U.u11 = make_double2(u_re_im[6], u_re_im[7]);
  const uint64_t pairs = 1ull << (n_qubits - 1);
  k_gate_1q<<<vec_grid(pairs), kThreads, 0, (cudaStream_t)stream>>>((double2*)psi_dev, pairs, q, U);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

extern "C" int psum_gate_2q(void* psi_dev, int n_qubits, int kind, int control, int target, psum_stream_t stream) {
  if (!psi_dev || n_qubits < 2 || control < 0 || target < 0 || control >= n_qubits || target >= n_qubits ||
      control == target || kind < 0 || kind > 1)
    return fail(PSUM_ERR_INVALID, "bad argument");
  const uint64_t quads = 1ull << (n_qubits - 2);
  k_gate_cx_cz<<<vec_grid(quads), kThreads, 0, (cudaStream_t)stream>>>((double2*)psi_dev, quads, control, target, kind);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// A block of gates on at most 12 qubits in one HBM sweep (the tile of the block's qubits is staged in
// shared memory).  free_qubits: the L block qubits, ascending; when n_qubits >= 3 they must start with
// 0, 1, 2 (128-byte runs).  gates_host: n_gates records of 80 bytes {int32 kind, a, b, pad; double u[8]}
// with a, b = positions INSIDE free_qubits (tile bits).
extern "C" int psum_gate_block(void* psi_dev, int n_qubits, const int32_t* free_qubits, int n_free,
                               const void* gates_host, int n_gates, psum_stream_t stream) {
  if (!psi_dev || !free_qubits || !gates_host || n_gates < 1) return fail(PSUM_ERR_INVALID, "bad argument");
  if (n_qubits < 1 || n_qubits > 40 || n_free < 1 || n_free > 12 || n_free > n_qubits)
    return fail(PSUM_ERR_INVALID, "gate block: 1 <= n_free <= min(12, n_qubits)");
  GateBlockParams P;
  memset(&P, 0, sizeof(P));
  uint64_t fm = 0;
  for (int j = 0; j < n_free; ++j) {
    const int q = free_qubits[j];
    if (q < 0 || q >= n_qubits || (j > 0 && q <= free_qubits[j - 1])) return fail(PSUM_ERR_INVALID, "free_qubits must ascend");
    if (n_qubits >= 3 && j < 3 && q != j && n_free >= 3) return fail(PSUM_ERR_INVALID, "the three lowest qubits must be free");
    P.fbit[j] = (uint8_t)q;
    fm |= 1ull << q;
  }
  const GateRec* gh = (const GateRec*)gates_host;
  for (int k = 0; k < n_gates; ++k) {
    const bool two = gh[k].kind != 0;
    if (gh[k].kind < 0 || gh[k].kind > 2 || gh[k].a < 0 || gh[k].a >= n_free ||
        (two && (gh[k].b < 0 || gh[k].b >= n_free || gh[k].b == gh[k].a)))
      return fail(PSUM_ERR_INVALID, "gate %d is malformed", k);
  }
  P.L = n_free;
  P.n_nonfree = n_qubits - n_free;
  P.ntiles = 1ull << P.n_nonfree;
  int nj = 0;
  for (int b = 0; b < n_qubits; ++b)
    if (!(fm >> b & 1ull)) P.nbit[nj++] = (uint8_t)b;
  cudaStream_t st = (cudaStream_t)stream;
  // device copy of the gate list (per-thread scratch, grows on demand)
  static thread_local GateRec* d_gates = nullptr;
  static thread_local int d_cap = 0;
  if (d_cap < n_gates) {
    CUDA_TRY(cudaStreamSynchronize(st));
    if (d_gates) cudaFree(d_gates);
    d_cap = std::max(256, 2 * n_gates);
    CUDA_TRY(cudaMalloc(&d_gates, sizeof(GateRec) * (size_t)d_cap));
  } else {
    CUDA_TRY(cudaStreamSynchronize(st));   // the previous block may still read the scratch
  }
  CUDA_TRY(cudaMemcpyAsync(d_gates, gh, sizeof(GateRec) * (size_t)n_gates, cudaMemcpyHostToDevice, st));
  static bool attr_done = false;
  const size_t smem = ((size_t)16 << n_free) + 8 * 128 + 8 * 9 * 64;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(k_gate_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)16 << 12) + 8 * 128 + 8 * 9 * 64)));
    attr_done = true;
  }
  int dev = 0, sms = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)std::min<uint64_t>(P.ntiles, (uint64_t)sms * 2);
  k_gate_block<<<grid, kThreads, smem, st>>>((double2*)psi_dev, P, d_gates, n_gates);
  CUDA_TRY(cudaGetLastError());
  return PSUM_OK;
}

// ---------------------------------------------------------------------------------------------
// sharding
// ---------------------------------------------------------------------------------------------
extern "C" int psum_nccl_unique_id(void* out128) {
  if (!out128) return fail(PSUM_ERR_INVALID, "null pointer");
  PS_TRY(load_nccl());
  nccl_uid_t id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return PSUM_OK;
}

extern "C" int psum_shard_init(psum_handle_t h, int rank, int world, const void* unique_id128) {
  if (!h) return fail(PSUM_ERR_INVALID, "null handle");
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
    return fail(PSUM_ERR_INVALID, "world must be a power of two and 0 <= rank < world");
  int p = 0;
  while ((1 << p) < world) ++p;
  if (p >= h->n) return fail(PSUM_ERR_INVALID, "world too large for %d qubits", h->n);
  h->p_bits = p;
  h->rank = rank;
  h->world = world;
  h->n_loc = h->n - p;
  PS_TRY(plan(h));
  if (unique_id128 && world > 1) {
    PS_TRY(load_nccl());
    nccl_uid_t id;
    memcpy(&id, unique_id128, 128);
    NCCL_TRY(g_nccl.CommInitRank(&h->comm, world, id, rank));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    for (int i = 0; i < 8; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_recv[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_used[i], cudaEventDisableTiming));
    }
  }
  return PSUM_OK;
}

extern "C" int psum_shard_terms(psum_handle_t h, int g, int64_t cap, uint64_t* x_local, uint64_t* z_local,
                                double* coeff_re_im, int64_t* n_out) {
  if (!h || !n_out) return fail(PSUM_ERR_INVALID, "null argument");
  if (!h->planned) PS_TRY(plan(h));
  PartHost* P = find_part(h, g);
  int64_t n = P ? (int64_t)P->comp.size() : 0;
  *n_out = n;
  if (!P || cap < n || !x_local || !z_local || !coeff_re_im) return PSUM_OK;
  for (int64_t i = 0; i < n; ++i) {
    x_local[i] = P->comp[i].xl;
    z_local[i] = P->comp[i].zl;
    coeff_re_im[2 * i] = P->comp[i].im ? 0.0 : P->comp[i].val;
    coeff_re_im[2 * i + 1] = P->comp[i].im ? P->comp[i].val : 0.0;
  }
  return PSUM_OK;
}

extern "C" int psum_apply_part(psum_handle_t h, int g, const void* src_dev, const void* bra_dev, void* y_dev,
                               int accumulate, double* expect_acc, psum_stream_t stream) {
  if (!h || !src_dev) return fail(PSUM_ERR_INVALID, "null argument");
  if (!y_dev && !expect_acc) return fail(PSUM_ERR_INVALID, "nothing to compute");
  if (((uintptr_t)src_dev | (uintptr_t)y_dev | (uintptr_t)bra_dev) & 15u)
    return fail(PSUM_ERR_INVALID, "state vectors must be 16-byte aligned");
  PS_TRY(upload(h));
  cudaStream_t st = (cudaStream_t)stream;
  const DevPart* D = nullptr;
  for (auto& d : h->dparts)
    if (d.g == g) D = &d;
  if (!D) {  // no term has this global part: contribution is zero
    if (y_dev && !accumulate) CUDA_TRY(cudaMemsetAsync(y_dev, 0, n_local_amps(h) * sizeof(double2), st));
    return PSUM_OK;
  }
  size_t pcount = 0;
  const double2* bra = bra_dev ? (const double2*)bra_dev : (const double2*)src_dev;
  PS_TRY(run_part(h, *D, (const double2*)src_dev, bra, (double2*)y_dev, accumulate, expect_acc != nullptr, true, &pcount, st));
  if (expect_acc) {
    double v[2];
    k_reduce_partials<<<1, kThreads, 0, st>>>(h->d_partial, (int)pcount, h->d_scalar, 0);
    CUDA_TRY(cudaGetLastError());
    PS_TRY(fetch_scalars(h, 1, v, st));
    expect_acc[0] = v[0];
    expect_acc[1] = v[1];
  }
  return PSUM_OK;
}

extern "C" int psum_allreduce_sum(psum_handle_t h, double* host_vals, int n) {
  if (!h || !host_vals || n < 0 || n > 256) return fail(PSUM_ERR_INVALID, "bad argument");
  if (h->world == 1) return PSUM_OK;
  if (!h->comm) return fail(PSUM_ERR_NCCL, "no communicator (psum_shard_init without id)");
  PS_TRY(upload(h));
  double* d = (double*)h->d_scalar;
  double* hs = (double*)h->h_scalar;
  for (int i = 0; i < n; ++i) hs[i] = host_vals[i];
  CUDA_TRY(cudaMemcpyAsync(d, hs, sizeof(double) * n, cudaMemcpyHostToDevice, h->comm_stream));
  NCCL_TRY(g_nccl.AllReduce(d, d, n, kNcclFloat64, kNcclSum, h->comm, h->comm_stream));
  CUDA_TRY(cudaMemcpyAsync(hs, d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->comm_stream));
  CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  for (int i = 0; i < n; ++i) host_vals[i] = hs[i];
  return PSUM_OK;
}

// Sharded H.psi: the local part runs on `stream` while the first partner shard is exchanged on the
// communication stream; each further global part g alternates between the two halves of recv_dev
// when it holds >= 2 shards, so exchange g+1 overlaps the passes over the shard of g.
// Builds (once) the chunk-friendly plan used when the state arrives from host memory.
static int ensure_host_plan(psum_handle_s* h, int late) {
  if (!h->parts_h.empty() && h->late_bits_h == late) return PSUM_OK;
  PlanOptions o = h->opt;
  o.late_bits = late;
  o.late_split = h->late_split;
  h->parts_h = build_parts(h->terms, h->n, h->p_bits, h->rank, o);
  if (h->arena_h) cudaFree(h->arena_h);
  h->arena_h = nullptr;
  size_t need = 0;
  PS_TRY(upload_plan(h, h->parts_h, &h->arena_h, &h->dparts_h, &need));
  h->late_bits_h = late;
  // the chunked path launches every chunk-local pass once per chunk, each launch with its own
  // block of partials: grow the partial buffer when a large plan needs more than upload() reserved
  size_t need_h = 0;
  const size_t nchunks = (size_t)1 << late;
  for (const DevPart& D : h->dparts_h) {
    need_h += (size_t)D.grid_stream;
    for (const DevPass& dp : D.passes) need_h += (size_t)dp.grid * nchunks;   // upper bound: one block per tile range
  }
  if (need_h + 4096 > h->partial_cap) {
    CUDA_TRY(cudaDeviceSynchronize());  // earlier launches may still read the old buffer
    if (h->d_partial) cudaFree(h->d_partial);
    h->d_partial = nullptr;
    h->partial_cap = need_h + 4096;
    CUDA_TRY(cudaMalloc(&h->d_partial, h->partial_cap * sizeof(double2)));
  }
  return PSUM_OK;
}

static int host_late_bits(const psum_handle_s* h) {
  if (h->n_loc < 20 || h->opt.kernel == 1) return 0;
  const int want = h->host_chunk_bits > 0 ? h->host_chunk_bits : 4;
  return std::max(0, std::min(want, h->n_loc - std::max(11, std::min(h->opt.tile_bits, kMaxTileBits))));
}

// Peer pointers of the source vector: every rank exports the allocation that holds its `psi` through
// CUDA IPC, the (handle, offset) records are all-gathered over NCCL, and the partners' allocations are
// opened (cached).  peers[r] = device pointer, valid on THIS device, of rank r's shard.  Returns
// PSUM_OK with peers empty when IPC is not available for this buffer (the caller falls back to NCCL).
static int resolve_peers(psum_handle_s* h, const void* psi_dev, std::vector<const double2*>* peers) {
  peers->clear();
  struct Rec {
    cudaIpcMemHandle_t handle;
    unsigned long long offset;
    unsigned long long ok;
  };
  static_assert(sizeof(Rec) == 80, "IPC record layout");
  Rec mine;
  memset(&mine, 0, sizeof(mine));
  if (load_cuda_driver() == PSUM_OK) {
    unsigned long long base = 0;
    size_t size = 0;
    if (g_cu_get_range(&base, &size, (unsigned long long)(uintptr_t)psi_dev) == 0 &&
        cudaIpcGetMemHandle(&mine.handle, (void*)(uintptr_t)base) == cudaSuccess) {
      mine.offset = (unsigned long long)(uintptr_t)psi_dev - base;
      mine.ok = 1;
    } else {
      cudaGetLastError();
    }
  }
  if (!h->d_ipc) CUDA_TRY(cudaMalloc(&h->d_ipc, sizeof(Rec) * (size_t)(h->world + 1)));
  std::vector<Rec> all((size_t)h->world);
  char* d = (char*)h->d_ipc;
  CUDA_TRY(cudaMemcpyAsync(d + sizeof(Rec) * h->world, &mine, sizeof(Rec), cudaMemcpyHostToDevice, h->comm_stream));
  NCCL_TRY(g_nccl.AllGather(d + sizeof(Rec) * h->world, d, sizeof(Rec), kNcclUint8, h->comm, h->comm_stream));
  CUDA_TRY(cudaMemcpyAsync(all.data(), d, sizeof(Rec) * h->world, cudaMemcpyDeviceToHost, h->comm_stream));
  CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  for (int r = 0; r < h->world; ++r)
    if (!all[r].ok) return PSUM_OK;   // some rank cannot export: everybody falls back
  peers->resize((size_t)h->world, nullptr);
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) {
      (*peers)[r] = (const double2*)psi_dev;
      continue;
    }
    const std::string key((const char*)&all[r].handle, sizeof(cudaIpcMemHandle_t));
    void* base = nullptr;
    for (auto& kv : h->ipc_open)
      if (kv.first == key) base = kv.second;
    if (!base) {
      cudaError_t e = cudaIpcOpenMemHandle(&base, all[r].handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        peers->clear();
        // the decision must be the same on every rank: agree through the flag all-reduce below
        base = nullptr;
      } else {
        h->ipc_open.emplace_back(key, base);
      }
    }
    if (peers->empty()) break;
    (*peers)[r] = (const double2*)((const char*)base + all[r].offset);
  }
  // all ranks must take the same path
  double flag = peers->empty() ? 1.0 : 0.0;
  double* dflag = (double*)h->d_ipc;
  CUDA_TRY(cudaMemcpyAsync(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice, h->comm_stream));
  NCCL_TRY(g_nccl.AllReduce(dflag, dflag, 1, kNcclFloat64, kNcclSum, h->comm, h->comm_stream));
  CUDA_TRY(cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, h->comm_stream));
  CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  if (flag != 0.0) peers->clear();
  return PSUM_OK;
}

// Sharded H.psi / <psi|H|psi>.  host_src != NULL: this rank's shard comes from host memory (y must
// be NULL): the H2D copy is chunked on the copy stream, the chunk-local passes of the local part
// follow it chunk by chunk, and the exchanges start when the last chunk has landed.
static int sharded_impl(psum_handle_s* h, const void* psi_dev, void* y_dev, void* recv_dev, uint64_t recv_amps,
                        double* expect_out, cudaStream_t st, const void* host_src) {
  const uint64_t N = n_local_amps(h);
  const double2* psi = (const double2*)psi_dev;
  double2* y = (double2*)y_dev;
  const int late = host_src ? host_late_bits(h) : 0;
  if (host_src && late > 0) PS_TRY(ensure_host_plan(h, late));
  const std::vector<DevPart>& parts = (host_src && late > 0) ? h->dparts_h : h->dparts;
  bool any_remote = false;
  for (auto& D : parts)
    if (D.g != 0) any_remote = true;
  if (any_remote) {
    if (!h->comm) return fail(PSUM_ERR_NCCL, "terms with global x-parts need a communicator");
    if (!recv_dev || recv_amps < N) return fail(PSUM_ERR_INVALID, "recv_dev must hold one shard");
  }
  std::vector<const double2*> peers;
  if (any_remote && h->peer_copy && !h->skip_exchange) PS_TRY(resolve_peers(h, psi_dev, &peers));
  // ring of receive buffers: up to nbuf exchanges are in flight ahead of the part being computed
  const int nbuf = (int)std::max<uint64_t>(1, std::min<uint64_t>(8, recv_amps / std::max<uint64_t>(N, 1)));
  size_t pcount = 0;
  bool wrote_y = false;
  const DevPart* local = nullptr;
  std::vector<const DevPart*> remote;
  for (auto& D : parts) {
    if (D.g == 0) local = &D;
    else remote.push_back(&D);
  }
  size_t n_early = 0;
  CopyGuard guard;
  if (host_src) {
    // ---- chunked H2D + the chunk-local passes of the local part -------------------------------
    if (!h->copy_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_free, cudaEventDisableTiming));
    }
    const uint64_t nchunks = 1ull << late;
    for (uint64_t k = 0; k < nchunks; ++k)
      if (!h->ev_chunk[k]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_chunk[k], cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(h->ev_free, st));
    CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_free, 0));
    const size_t cbytes = (N * sizeof(double2)) >> late;
    const int cbits = h->n_loc - late;
    const bool pipelined = local && late > 0 && !local->passes.empty() && !y;
    const bool pageable = host_is_pageable(host_src);
    guard.cs = h->copy_stream;
    for (uint64_t k = 0; k < nchunks; ++k) {
      char* dst = (char*)psi_dev + k * cbytes;
      const char* srcp = (const char*)host_src + k * cbytes;
      if (pageable) PS_TRY(bounce_copy(h, srcp, dst, cbytes, h->copy_stream));
      else CUDA_TRY(cudaMemcpyAsync(dst, srcp, cbytes, cudaMemcpyHostToDevice, h->copy_stream));
      CUDA_TRY(cudaEventRecord(h->ev_chunk[k], h->copy_stream));
      CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chunk[k], 0));
      if (pipelined) PS_TRY(run_ready(h, *local, psi, k, late, cbits, &pcount, st));
    }
    if (pipelined) n_early = local->passes.size();   // every tile range has been launched
  }
  if (any_remote) {
    CUDA_TRY(cudaEventRecord(h->ev_ready, st));   // st has waited for every chunk by now
    CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    // pulling needs the PARTNER's source vector to be complete: a one-word all-reduce on the
    // communication streams is the cross-rank "psi is ready" barrier (send/recv pairs imply it)
    if (!peers.empty())
      NCCL_TRY(g_nccl.AllReduce(h->d_ipc, h->d_ipc, 1, kNcclFloat64, kNcclSum, h->comm, h->comm_stream));
  }
  auto post_exchange = [&](const DevPart* D, int s, bool must_wait_used) -> int {
    if (must_wait_used) CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_used[s], 0));
    double2* buf = (double2*)recv_dev + (uint64_t)s * N;
    const int peer = h->rank ^ D->g;
    if (h->profile_exchange) {
      while (h->ex_ev.size() < (size_t)(2 * (h->ex_count + 1))) {
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreate(&ev));
        h->ex_ev.push_back(ev);
      }
      CUDA_TRY(cudaEventRecord(h->ex_ev[2 * h->ex_count], h->comm_stream));
    }
    if (!h->skip_exchange && !peers.empty()) {
      // pull the partner shard with the copy engines over NVLink
      CUDA_TRY(cudaMemcpyAsync(buf, peers[(size_t)peer], N * sizeof(double2), cudaMemcpyDeviceToDevice, h->comm_stream));
    } else if (!h->skip_exchange) {
      NCCL_TRY(g_nccl.GroupStart());
      NCCL_TRY(g_nccl.Send(psi, 2 * N, kNcclFloat64, peer, h->comm, h->comm_stream));
      NCCL_TRY(g_nccl.Recv(buf, 2 * N, kNcclFloat64, peer, h->comm, h->comm_stream));
      NCCL_TRY(g_nccl.GroupEnd());
    }
    if (h->profile_exchange) {
      CUDA_TRY(cudaEventRecord(h->ex_ev[2 * h->ex_count + 1], h->comm_stream));
      ++h->ex_count;
    }
    CUDA_TRY(cudaEventRecord(h->ev_recv[s], h->comm_stream));
    return PSUM_OK;
  };
  size_t posted = 0;
  h->ex_count = 0;
  struct ReserveGuard {
    psum_handle_s* h;
    ~ReserveGuard() { h->grid_reserve = 0; }
  } reserve_guard{h};
  h->grid_reserve = (any_remote && !h->skip_exchange) ? h->comm_sms : 0;
  for (; posted < remote.size() && posted < (size_t)nbuf; ++posted)
    PS_TRY(post_exchange(remote[posted], (int)(posted % nbuf), false));
  // (rest of the) local part: overlaps the exchanges posted above
  if (local) {
    if (!host_src || n_early < local->passes.size() || local->passes.empty())
      PS_TRY(run_part(h, *local, psi, psi, y, 0, expect_out != nullptr, remote.empty(), &pcount, st, n_early,
                      (size_t)-1));
    wrote_y = true;
  }
  if (!wrote_y && y) CUDA_TRY(cudaMemsetAsync(y, 0, N * sizeof(double2), st));
  for (size_t r = 0; r < remote.size(); ++r) {
    const int slot = (int)(r % nbuf);
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_recv[slot], 0));
    const double2* buf = (const double2*)recv_dev + (uint64_t)slot * N;
    PS_TRY(run_part(h, *remote[r], buf, psi, y, 1, expect_out != nullptr, r + 1 == remote.size(), &pcount, st));
    CUDA_TRY(cudaEventRecord(h->ev_used[slot], st));
    if (posted < remote.size()) {  // refill the slot that was just consumed
      PS_TRY(post_exchange(remote[posted], slot, true));
      ++posted;
    }
  }
  // a partner may still be pulling this rank's psi: nobody leaves before every rank has consumed its
  // pulls (with an expectation the all-reduce below is that barrier)
  if (!peers.empty() && !expect_out)
    NCCL_TRY(g_nccl.AllReduce(h->d_scalar + 128, h->d_scalar + 128, 1, kNcclFloat64, kNcclSum, h->comm, st));
  if (expect_out) {
    k_reduce_partials<<<1, kThreads, 0, st>>>(h->d_partial, (int)pcount, h->d_scalar, 0);
    CUDA_TRY(cudaGetLastError());
    if (h->comm) NCCL_TRY(g_nccl.AllReduce(h->d_scalar, h->d_scalar, 2, kNcclFloat64, kNcclSum, h->comm, st));
    PS_TRY(fetch_scalars(h, 1, expect_out, st));
    guard.cs = nullptr;   // st waited for every chunk and is drained: the host copies are complete
  }
  if (h->profile_exchange && h->ex_count > 0) {
    // per-exchange durations on the communication stream (the stream is serial, so their sum is
    // its busy time) and the span from the first post to the last arrival
    CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
    double sum = 0.0;
    for (int i = 0; i < h->ex_count; ++i) {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, h->ex_ev[2 * i], h->ex_ev[2 * i + 1]));
      sum += ms;
    }
    float span = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&span, h->ex_ev[0], h->ex_ev[2 * h->ex_count - 1]));
    h->ex_stats[0] = h->ex_count;
    h->ex_stats[1] = (double)(N * sizeof(double2));
    h->ex_stats[2] = sum;
    h->ex_stats[3] = span;
  }
  return PSUM_OK;
}

extern "C" int psum_exchange_stats(psum_handle_t h, double out[4]) {
  if (!h || !out) return fail(PSUM_ERR_INVALID, "null argument");
  for (int i = 0; i < 4; ++i) out[i] = h->ex_stats[i];
  return PSUM_OK;
}

extern "C" int psum_apply_sharded(psum_handle_t h, const void* psi_dev, void* y_dev, void* recv_dev,
                                  uint64_t recv_amps, double* expect_out, psum_stream_t stream) {
  if (!h || !psi_dev) return fail(PSUM_ERR_INVALID, "null argument");
  if (!y_dev && !expect_out) return fail(PSUM_ERR_INVALID, "nothing to compute");
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  if (((uintptr_t)psi_dev | (uintptr_t)y_dev | (uintptr_t)recv_dev) & 15u)
    return fail(PSUM_ERR_INVALID, "state vectors must be 16-byte aligned");
  PS_TRY(upload(h));
  return sharded_impl(h, psi_dev, y_dev, recv_dev, recv_amps, expect_out, (cudaStream_t)stream, nullptr);
}

// Sharded <psi|H|psi> for a shard that lives in HOST memory (pinned for full speed); psi_dev
// receives the copy.  The multi-GPU form of psum_expect_host.
extern "C" int psum_expect_sharded_host(psum_handle_t h, const void* psi_host, void* psi_dev, void* recv_dev,
                                        uint64_t recv_amps, double out[2], psum_stream_t stream) {
  if (!h || !psi_host || !psi_dev || !out) return fail(PSUM_ERR_INVALID, "null argument");
  if (h->terms.empty()) return fail(PSUM_ERR_EMPTY, "Operator is empty, check the operator.");
  if (((uintptr_t)psi_dev | (uintptr_t)recv_dev) & 15u)
    return fail(PSUM_ERR_INVALID, "state vectors must be 16-byte aligned");
  PS_TRY(upload(h));
  return sharded_impl(h, psi_dev, nullptr, recv_dev, recv_amps, out, (cudaStream_t)stream, psi_host);
}

#include "psum_lanczos.inc"
