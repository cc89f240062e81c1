// nsgym_abi.cu -- the C ABI of include/nsgym_b200.h: handle life cycle, buffer layout,
// launches, and the host-buffer (end-to-end) step pipeline.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "nsgym_host.h"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define NSG_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) return fail(-10, "%s: %s", #call, cudaGetErrorString(e__));      \
  } while (0)

constexpr int kHostStreams = 4;
// NSGYM_OPT_SPECIALIZE auto: the ~0.3 s NVRTC compilation pays for itself on throughput-sized batches; a
// small batch is launch-latency bound either way and keeps the precompiled kernel
constexpr int64_t kSpecializeMinEnvs = 32768;

struct KindInfo { int state_words, obs_words, n_theta; bool box; };
const KindInfo kKinds[NSGYM_ENV_COUNT] = {
    {4, 4, 6, false}, {4, 6, 8, false}, {2, 2, 2, false}, {2, 2, 1, true},
    {2, 3, 4, true},  {1, 0, 1, false}, {1, 0, 1, false}, {1, 0, 3, false},
};

}  // namespace

struct NsgymHandle {
  NsgymSpec spec;
  nsg::DevicePools pools{nullptr, nullptr, nullptr, nullptr, 0};
  NsgymBuffers buf{};
  bool bound = false;
  bool initialised = false;
  uint64_t step_index = 0;
  int64_t launches = 0;
  int n_istate = 0;
  cudaStream_t streams[kHostStreams]{};
  cudaEvent_t host_event{};        // orders nsgym_step_host's side streams after the caller's stream
  bool streams_ready = false;
  nsg::RowTable rows;              // heterogeneous handles (nsgym_create_rows)
  int32_t plan_elapsed = -1;       // planning copy (nsgym_fanout): TimeLimit steps since the copy
  int32_t general_kernels = 0;     // NSGYM_OPT_GENERAL_KERNELS
  int32_t kernel_class = -1;       // instantiation picked by the last step / rollout launch (NSGYM_KERNEL_*)
  int32_t specialize = -1;         // NSGYM_OPT_SPECIALIZE: -1 auto (batches of >= kSpecializeMinEnvs envs), 0 never, 1 always
  int32_t specialized = 0;         // the last step / rollout launch went to a program-specialised kernel
  nsg::SpecCache spec_cache;       // program-specialised kernels of this handle (nsgym_jit.cu)

  bool grid() const { return nsg::is_grid_kind(spec.env_kind); }
  size_t real_bytes() const { return grid() ? 8 : (spec.precision == NSGYM_F64 ? 8 : 4); }
  size_t action_bytes() const { return kKinds[spec.env_kind].box ? real_bytes() : 4; }
  size_t state_bytes_per_env() const {
    return grid() ? 4 : size_t(kKinds[spec.env_kind].state_words) * real_bytes();
  }
  int theta_planes() const { return grid() ? spec.n_slots * spec.n_dist : spec.n_slots; }
};

namespace {

nsg::LaunchIO base_io(const NsgymHandle* h) {
  nsg::LaunchIO io{};
  io.state = h->buf.d_state; io.theta = h->buf.d_theta; io.t = h->buf.d_t; io.istate = h->buf.d_istate;
  io.action = h->buf.d_action;
  io.reward = h->buf.d_reward; io.flags = h->buf.d_flags; io.change = h->buf.d_change;
  io.delta = h->buf.d_delta; io.obs = h->buf.d_obs;
  io.n = h->spec.n_envs; io.begin = 0; io.count = h->spec.n_envs;
  io.gid_offset = uint64_t(h->spec.env_id_offset); io.seed = h->spec.seed; io.step_index = h->step_index;
  io.gamma = 1.f;
  io.rows = h->rows.active ? &h->rows : nullptr;
  io.plan_elapsed = h->plan_elapsed;
  io.general_kernels = h->general_kernels;
  io.sched_replay = h->spec.persistent_params ? 0 : 1;
  io.specialize = nsg::jit::enabled() && (h->specialize > 0 || (h->specialize < 0 && h->spec.n_envs >= kSpecializeMinEnvs));
  io.spec_cache = const_cast<nsg::SpecCache*>(&h->spec_cache);
  return io;
}

cudaError_t dispatch(NsgymHandle* h, nsg::LaunchOp op, const nsg::LaunchIO& io_in, cudaStream_t s) {
  h->launches += 1;
  nsg::LaunchIO io = io_in;
  int32_t cls = -1;
  int32_t spec = 0;
  io.kernel_class = &cls;
  io.specialized = &spec;
  struct Keep {
    NsgymHandle* h; nsg::LaunchOp op; int32_t* c; int32_t* s;
    ~Keep() { if (op != nsg::OP_RESET && *c >= 0) { h->kernel_class = *c; h->specialized = *s; } }
  } keep{h, op, &cls, &spec};
  if (h->grid()) return nsg::launch_grid(op, h->spec, h->pools, io, s);
  if (h->spec.precision == NSGYM_F64) return nsg::launch_classic_f64(op, h->spec, h->pools, io, s);
  return nsg::launch_classic_f32(op, h->spec, h->pools, io, s);
}

int validate_slot(const NsgymSpec* s, const NsgymSlot& sl, int j) {
  const KindInfo& k = kKinds[s->env_kind];
  const bool grid = nsg::is_grid_kind(s->env_kind);
  if (sl.theta_index < 0 || sl.theta_index >= k.n_theta) return fail(-1, "slot %d: theta_index out of range", j);
  if (sl.sched_op < 0 || sl.sched_op >= NSGYM_SCHED_COUNT) return fail(-1, "slot %d: bad sched_op", j);
  const bool dist_op = sl.upd_op >= NSGYM_UPD_D_NOP;
  if (dist_op != grid) return fail(-1, "slot %d: update opcode %d does not fit this env kind", j, sl.upd_op);
  // the Acrobot cross-checks are wired to fixed partners (classic_control.py:241-265, :307-357)
  if (sl.constraint == NSGYM_CONS_ACRO_LENGTH1 && !(s->env_kind == NSGYM_ENV_ACROBOT && sl.theta_index == 1))
    return fail(-1, "slot %d: NSGYM_CONS_ACRO_LENGTH1 belongs to Acrobot LINK_LENGTH_1", j);
  if (sl.constraint == NSGYM_CONS_ACRO_COM &&
      !(s->env_kind == NSGYM_ENV_ACROBOT && (sl.theta_index == 5 || sl.theta_index == 6)))
    return fail(-1, "slot %d: NSGYM_CONS_ACRO_COM belongs to Acrobot LINK_COM_POS_1 / LINK_COM_POS_2", j);
  if (sl.sched_op == NSGYM_SCHED_PERIODIC && sl.si[0] <= 0) return fail(-1, "slot %d: period must be > 0", j);
  if (sl.sched_op == NSGYM_SCHED_BURST && sl.si[1] <= 0) return fail(-1, "slot %d: burst cycle must be > 0", j);
  if (sl.sched_op == NSGYM_SCHED_BITMAP &&
      (sl.si[0] < 0 || (sl.si[0] + (sl.si[1] + 31) / 32) > s->n_bitmap_words))
    return fail(-1, "slot %d: bitmap range outside the pool", j);
  if (sl.sched_op == NSGYM_SCHED_WINDOW && (sl.si[0] < 0 || sl.si[0] + 2 * sl.si[1] > s->n_pool_i))
    return fail(-1, "slot %d: window list outside the pool", j);
  const int per = dist_op ? s->n_dist : 1;
  switch (sl.upd_op) {
    case NSGYM_UPD_POLY: case NSGYM_UPD_STEPWISE: case NSGYM_UPD_CYCLIC:
    case NSGYM_UPD_D_STEPWISE: case NSGYM_UPD_D_CYCLIC:
      if (sl.ui[0] < 0 || sl.ui[1] < 0 || sl.ui[0] + sl.ui[1] * per > s->n_pool_f)
        return fail(-1, "slot %d: value list outside the pool", j);
      if ((sl.upd_op == NSGYM_UPD_CYCLIC || sl.upd_op == NSGYM_UPD_D_CYCLIC) && sl.ui[1] == 0)
        return fail(-1, "slot %d: cyclic list is empty", j);
      break;
    case NSGYM_UPD_D_LERP:
      if (sl.ui[0] < 0 || sl.ui[0] + 2 * per > s->n_pool_f) return fail(-1, "slot %d: lerp data outside the pool", j);
      break;
    default: break;
  }
  return 0;
}

int validate(const NsgymSpec* s) {
  if (!s) return fail(-1, "spec is NULL");
  if (s->abi_version != NSGYM_ABI_VERSION)
    return fail(-1, "ABI version mismatch: caller %d, library %d", s->abi_version, NSGYM_ABI_VERSION);
  if (s->env_kind < 0 || s->env_kind >= NSGYM_ENV_COUNT) return fail(-1, "unknown env_kind %d", s->env_kind);
  if (s->n_envs <= 0) return fail(-1, "n_envs must be positive");
  if (s->n_envs > (int64_t(1) << 28)) return fail(-1, "n_envs per handle is limited to 2^28 (32-bit plane indexing)");
  if (s->n_slots < 0 || s->n_slots > NSGYM_MAX_SLOTS) return fail(-1, "n_slots %d out of range", s->n_slots);
  const KindInfo& k = kKinds[s->env_kind];
  if (s->n_slots > k.n_theta) return fail(-1, "%d slots > %d tunable parameters of this env", s->n_slots, k.n_theta);
  const bool grid = nsg::is_grid_kind(s->env_kind);
  if (grid) {
    const int want = s->env_kind == NSGYM_ENV_CLIFFWALKING ? 4 : 3;
    if (s->n_dist != want) return fail(-1, "n_dist %d, this env needs %d", s->n_dist, want);
    if (s->nrow <= 0 || s->ncol <= 0 || s->nrow * s->ncol > 256)
      return fail(-1, "gridworld maps are limited to 256 cells (got %dx%d)", s->nrow, s->ncol);
    // start_cell == -1: FrozenLake map with several 'S' cells, a reset samples one (toy_text.py:314-319 accepts any desc)
    if (s->start_cell >= s->nrow * s->ncol || s->start_cell < -1 || (s->start_cell < 0 && s->env_kind != NSGYM_ENV_FROZENLAKE))
      return fail(-1, "start_cell out of range");
    if (s->env_kind != NSGYM_ENV_BRIDGE && s->n_slots != 1) return fail(-1, "this env has exactly one parameter, P");
  } else if (s->precision != NSGYM_F32 && s->precision != NSGYM_F64) {
    return fail(-1, "precision must be NSGYM_F32 or NSGYM_F64");
  }
  for (int j = 0; j < s->n_slots; ++j) {
    if (int rc = validate_slot(s, s->slots[j], j)) return rc;
    for (int q = 0; q < j; ++q)
      if (s->slots[q].theta_index == s->slots[j].theta_index) return fail(-1, "slot %d: parameter bound twice", j);
  }
  return 0;
}

template <typename T>
int upload(const T* src, int n, const T** dst) {
  *dst = nullptr;
  if (n <= 0 || !src) return 0;
  T* d = nullptr;
  NSG_CUDA(cudaMalloc(&d, sizeof(T) * size_t(n)));
  NSG_CUDA(cudaMemcpy(d, src, sizeof(T) * size_t(n), cudaMemcpyHostToDevice));
  *dst = d;
  return 0;
}

}  // namespace

namespace {

// lane i of dst <- env i / fanout of src, `words` 32-bit words per env record (packed records)
__global__ void fanout_records_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t total,
                                      uint32_t words, uint32_t fanout) {
  const uint64_t idx = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const uint64_t env = idx / words;
  const uint32_t w = uint32_t(idx - env * words);
  dst[idx] = src[(env / fanout) * words + w];
}
// plane layout [planes][n]: dst[p][i] <- src[p][i / fanout] (or a constant per plane)
__global__ void fanout_planes_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t n_dst,
                                     uint32_t n_src, uint32_t planes, uint32_t words, uint32_t fanout,
                                     uint32_t and_mask) {
  const uint64_t idx = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;     // over n_dst * words
  if (idx >= uint64_t(n_dst) * words) return;
  const uint32_t i = uint32_t(idx / words), w = uint32_t(idx % words);
  for (uint32_t p = 0; p < planes; ++p)
    dst[(uint64_t(p) * n_dst + i) * words + w] = src[(uint64_t(p) * n_src + i / fanout) * words + w] & and_mask;
}
struct PlaneConsts { uint32_t lo[NSGYM_MAX_SLOTS * NSGYM_MAX_DIST], hi[NSGYM_MAX_SLOTS * NSGYM_MAX_DIST]; };
__global__ void fill_planes_kernel(uint32_t* __restrict__ dst, uint32_t n_dst, uint32_t planes, uint32_t words,
                                   const __grid_constant__ PlaneConsts c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dst) return;
  for (uint32_t p = 0; p < planes; ++p) {
    dst[(uint64_t(p) * n_dst + i) * words] = c.lo[p];
    if (words == 2) dst[(uint64_t(p) * n_dst + i) * words + 1] = c.hi[p];
  }
}

// NSWrapper.step packaging (base.py:314-361): flags / change mask / time word -> per-field arrays.
// Four envs per thread: 32-bit accesses to the byte arrays, 128-bit accesses to the time words
// (torch allocations are 256-byte aligned and every plane offset j * n keeps 4-byte alignment
// when n is a multiple of 4; other batch sizes take the scalar tail path for every env).
__device__ __forceinline__ void unpack_one(uint32_t f, uint32_t c, int32_t tw, uint32_t i, uint32_t n, int n_slots,
                                           uint8_t* terminated, uint8_t* truncated, uint8_t* was_reset,
                                           int32_t* rel_time, uint8_t* env_change) {
  if (terminated) terminated[i] = (f & NSGYM_FLAG_TERMINATED) ? 1 : 0;
  if (truncated) truncated[i] = (f & NSGYM_FLAG_TRUNCATED) ? 1 : 0;
  if (was_reset) was_reset[i] = (f & NSGYM_FLAG_RESET) ? 1 : 0;
  if (rel_time) rel_time[i] = tw & 0x0FFFFFFF;
  if (env_change)
    for (int j = 0; j < n_slots; ++j) env_change[uint32_t(j) * n + i] = (c >> j) & 1u;
}

__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t* __restrict__ flags, const uint8_t* __restrict__ change, const int32_t* __restrict__ t,
              uint32_t n, int n_slots, uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated,
              uint8_t* __restrict__ was_reset, int32_t* __restrict__ rel_time, uint8_t* __restrict__ env_change) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;      // group of four envs
  const uint32_t i0 = q * 4;
  if (i0 >= n) return;
  if ((n & 3u) == 0) {
    const uint32_t f4 = reinterpret_cast<const uint32_t*>(flags)[q];
    const uint32_t c4 = reinterpret_cast<const uint32_t*>(change)[q];
    const int4 t4 = reinterpret_cast<const int4*>(t)[q];
    // bit k of byte b of f4 -> byte b of the result, as 0 / 1
    if (terminated) reinterpret_cast<uint32_t*>(terminated)[q] = f4 & 0x01010101u;
    if (truncated) reinterpret_cast<uint32_t*>(truncated)[q] = (f4 >> 1) & 0x01010101u;
    if (was_reset) reinterpret_cast<uint32_t*>(was_reset)[q] = (f4 >> 2) & 0x01010101u;
    if (rel_time)
      reinterpret_cast<int4*>(rel_time)[q] = make_int4(t4.x & 0x0FFFFFFF, t4.y & 0x0FFFFFFF, t4.z & 0x0FFFFFFF,
                                                       t4.w & 0x0FFFFFFF);
    if (env_change)
      for (int j = 0; j < n_slots; ++j)
        reinterpret_cast<uint32_t*>(env_change + size_t(j) * n)[q] = (c4 >> j) & 0x01010101u;
    return;
  }
  for (uint32_t i = i0; i < n && i < i0 + 4; ++i)
    unpack_one(flags[i], change[i], t[i], i, n, n_slots, terminated, truncated, was_reset, rel_time, env_change);
}

inline unsigned blocks_for(uint64_t n) { return unsigned((n + 255) / 256); }

// Episode bookkeeping of a batch in one pass over the step's outputs (nsgym_episode_stats): running
// return / length per env, finished episodes folded into the totals.  Persistent grid (a few CTAs
// per SM, grid-stride): per-thread partial sums, warp shuffle + shared-memory reduction, one fp64
// atomicAdd per total per CTA.
constexpr int kStatTotals = NSGYM_STAT_COUNT;
__global__ void __launch_bounds__(256)
episode_stats_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ flags, uint32_t n,
                     double* __restrict__ run_ret, int32_t* __restrict__ run_len, double* __restrict__ totals) {
  double acc[kStatTotals];
#pragma unroll
  for (int k = 0; k < kStatTotals; ++k) acc[k] = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t f = flags[i];
    const bool stepped = !(f & NSGYM_FLAG_RESET);          // an autoreset call is not an env step
    const bool ended = (f & (NSGYM_FLAG_TERMINATED | NSGYM_FLAG_TRUNCATED)) != 0;
    double rr = run_ret[i] + (stepped ? double(reward[i]) : 0.0);
    int32_t rl = run_len[i] + (stepped ? 1 : 0);
    acc[NSGYM_STAT_STEPS] += stepped ? 1.0 : 0.0;
    acc[NSGYM_STAT_EPISODES] += ended ? 1.0 : 0.0;
    acc[NSGYM_STAT_RETURN_SUM] += ended ? rr : 0.0;
    acc[NSGYM_STAT_LENGTH_SUM] += ended ? double(rl) : 0.0;
    acc[NSGYM_STAT_TERMINATED] += (f & NSGYM_FLAG_TERMINATED) ? 1.0 : 0.0;
    acc[NSGYM_STAT_TRUNCATED] += (f & NSGYM_FLAG_TRUNCATED) ? 1.0 : 0.0;
    acc[NSGYM_STAT_REJECTED] += (f & NSGYM_FLAG_REJECTED) ? 1.0 : 0.0;
    acc[NSGYM_STAT_BAD_DIST] += (f & NSGYM_FLAG_BAD_DIST) ? 1.0 : 0.0;
    run_ret[i] = ended ? 0.0 : rr;
    run_len[i] = ended ? 0 : rl;
  }
  __shared__ double part[8][kStatTotals];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kStatTotals; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kStatTotals) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[w][threadIdx.x];
    if (v != 0.0) atomicAdd(totals + threadIdx.x, v);
  }
}

}  // namespace

namespace nsg {
int validate_row_slot(const NsgymSpec* spec, const NsgymSlot* slot, int j, char* err, size_t err_len) {
  const int rc = validate_slot(spec, *slot, j);
  if (rc && err && err_len) snprintf(err, err_len, "%s", g_error.c_str());
  return rc;
}
}  // namespace nsg

extern "C" {

int nsgym_abi_version(void) { return NSGYM_ABI_VERSION; }

size_t nsgym_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(NsgymSlot);
    case 1: return sizeof(NsgymSpec);
    case 2: return sizeof(NsgymLayout);
    case 3: return sizeof(NsgymBuffers);
    case 4: return sizeof(NsgymHostOut);
    case 5: return sizeof(NsgymSnapshotInfo);
    default: return 0;
  }
}

const char* nsgym_last_error(void) { return g_error.c_str(); }

int nsgym_create(const NsgymSpec* spec, NsgymHandle** out) {
  if (!out) return fail(-1, "out is NULL");
  *out = nullptr;
  if (int rc = validate(spec)) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(-2, "no CUDA device: ns_gym_b200 has no CPU execution path");
  NsgymHandle* h = new (std::nothrow) NsgymHandle();
  if (!h) return fail(-3, "out of host memory");
  h->spec = *spec;
  for (int j = 0; j < spec->n_slots; ++j) nsg::set_mod_magic(&h->spec.slots[j], *spec);
  int planes = 0;
  for (int j = 0; j < spec->n_slots; ++j)
    if (spec->slots[j].istate_plane >= 0) planes = planes > spec->slots[j].istate_plane + 1 ? planes : spec->slots[j].istate_plane + 1;
  h->n_istate = planes;
  int rc = upload(spec->pool_f, spec->n_pool_f, &h->pools.pool_f);
  if (!rc) rc = upload(spec->pool_i, spec->n_pool_i, &h->pools.pool_i);
  if (!rc) rc = upload(spec->bitmap, spec->n_bitmap_words, &h->pools.bitmap);
  if (!rc && h->grid()) {
    NsgymSpec with_map = *spec;                       // (h->spec keeps no host pointers)
    h->pools.grid_n_start = nsg::resolve_start_cells(&with_map);
    h->spec.start_cell = with_map.start_cell;
    uint32_t words[nsg::kGridTabWords];
    char err[256] = "";
    const int n_words = nsg::build_grid_tables(*spec, words, err, sizeof err);
    rc = n_words < 0 ? fail(-1, "%s", err) : upload(words, n_words, &h->pools.grid_tab);
  }
  if (rc) { nsgym_destroy(h); return rc; }
  h->spec.pool_f = nullptr; h->spec.pool_i = nullptr; h->spec.bitmap = nullptr; h->spec.cell_class = nullptr;
  *out = h;
  return 0;
}

int nsgym_create_rows(const NsgymSpec* spec, const NsgymSlot* rows, NsgymHandle** out) {
  if (!rows) return fail(-1, "rows is NULL");
  if (spec && spec->n_slots <= 0) return fail(-1, "a heterogeneous batch needs at least one bound parameter");
  NsgymHandle* h = nullptr;
  if (int rc = nsgym_create(spec, &h)) return rc;
  char err[384] = "";
  const bool dbl = h->grid() || spec->precision == NSGYM_F64;
  // h->spec keeps the device-side pool pointers out; the rows are lowered against the caller's spec
  if (int rc = nsg::build_rows(h->spec, rows, dbl, &h->rows, err, sizeof err)) {
    nsgym_destroy(h);
    return fail(rc, "%s", err);
  }
  *out = h;
  return 0;
}

void nsgym_destroy(NsgymHandle* h) {
  if (!h) return;
  nsg::free_rows(&h->rows);
  cudaFree(const_cast<double*>(h->pools.pool_f));
  cudaFree(const_cast<int32_t*>(h->pools.pool_i));
  cudaFree(const_cast<uint32_t*>(h->pools.bitmap));
  cudaFree(const_cast<uint32_t*>(h->pools.grid_tab));
  if (h->streams_ready) {
    for (auto& s : h->streams) cudaStreamDestroy(s);
    cudaEventDestroy(h->host_event);
  }
  delete h;
}

int nsgym_layout(const NsgymHandle* h, int want_delta, int want_obs, NsgymLayout* out) {
  if (!h || !out) return fail(-1, "NULL argument");
  const KindInfo& k = kKinds[h->spec.env_kind];
  const size_t n = size_t(h->spec.n_envs), w = h->real_bytes();
  std::memset(out, 0, sizeof *out);
  out->state = n * h->state_bytes_per_env();
  out->theta = n * size_t(h->theta_planes()) * w;
  out->t = n * 4;
  out->istate = n * size_t(h->n_istate) * 4;
  out->action = n * h->action_bytes();
  out->reward = n * 4;
  out->flags = n;
  out->change = n;
  out->delta = want_delta ? n * size_t(h->spec.n_slots) * w : 0;
  out->obs = (want_obs && k.obs_words) ? n * size_t(k.obs_words) * 4 : 0;
  out->state_words = k.state_words;
  out->obs_words = k.obs_words;
  out->n_istate = h->n_istate;
  out->theta_planes = h->theta_planes();
  // algorithmic bytes per env-step: state r+w, bound theta r+w, action r, t r+w,
  // reward w, flags w, change w [+ obs w] [+ delta w] [+ cursor planes r+w]
  double b = 2.0 * double(h->state_bytes_per_env()) + 2.0 * double(h->theta_planes()) * double(w) +
             double(h->action_bytes()) + 8.0 + 4.0 + 1.0 + 1.0;
  if (out->obs) b += 4.0 * k.obs_words;
  if (out->delta) b += double(h->spec.n_slots) * double(w);
  b += 8.0 * h->n_istate;
  out->row_bytes_per_env = h->rows.active ? h->rows.bytes_per_env : 0.0;
  b += out->row_bytes_per_env;
  out->bytes_per_step = b;
  return 0;
}

int nsgym_bind(NsgymHandle* h, const NsgymBuffers* b) {
  if (!h || !b) return fail(-1, "NULL argument");
  if (!b->d_state || !b->d_t || !b->d_action || !b->d_reward || !b->d_flags || !b->d_change)
    return fail(-1, "state, t, action, reward, flags and change buffers are mandatory");
  if (h->theta_planes() > 0 && !b->d_theta) return fail(-1, "theta buffer is mandatory when parameters are bound");
  if (h->n_istate > 0 && !b->d_istate) return fail(-1, "istate buffer is mandatory for list / Memoryless slots");
  if ((reinterpret_cast<uintptr_t>(b->d_state) & 31u) != 0) return fail(-1, "state buffer must be 32-byte aligned");
  h->buf = *b;
  h->bound = true;
  h->initialised = false;
  return 0;
}

int nsgym_reset(NsgymHandle* h, const uint8_t* d_mask, const double* d_inj_uniform, void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  if (!h->initialised && d_mask) return fail(-1, "the first reset must cover all envs (mask must be NULL)");
  nsg::LaunchIO io = base_io(h);
  io.mask = d_mask;
  io.inj_u = d_inj_uniform;
  io.force_init = h->initialised ? 0 : 1;
  cudaError_t e = dispatch(h, nsg::OP_RESET, io, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-10, "reset launch: %s", cudaGetErrorString(e));
  h->initialised = true;
  h->step_index += 1;
  if (!d_mask) h->plan_elapsed = -1;   // a full reset restarts TimeLimit and t together
  return 0;
}

int nsgym_step(NsgymHandle* h, const void* d_action, const double* d_inj_uniform, const double* d_inj_normal,
               int skip_updates, void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  if (!h->initialised) return fail(-4, "step before reset");
  nsg::LaunchIO io = base_io(h);
  if (d_action) io.action = d_action;
  io.inj_u = d_inj_uniform;
  io.inj_z = d_inj_normal;
  io.skip_updates = skip_updates;
  cudaError_t e = dispatch(h, nsg::OP_STEP, io, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-10, "step launch: %s", cudaGetErrorString(e));
  h->step_index += 1;
  if (h->plan_elapsed >= 0) h->plan_elapsed += 1;
  return 0;
}

static int ensure_streams(NsgymHandle* h) {
  if (!h->streams_ready) {
    for (auto& s : h->streams) NSG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    NSG_CUDA(cudaEventCreateWithFlags(&h->host_event, cudaEventDisableTiming));
    h->streams_ready = true;
  }
  return 0;
}

int nsgym_step_many(NsgymHandle* const* handles, const void* const* d_actions, int n, int skip_updates, void* stream) {
  if (!handles || n <= 0) return fail(-1, "no handles");
  for (int k = 0; k < n; ++k) {
    if (!handles[k] || !handles[k]->bound) return fail(-1, "handle %d not bound", k);
    if (!handles[k]->initialised) return fail(-4, "handle %d: step before reset", k);
  }
  // One host call for a batch of several env kinds (BASELINE config C4): the host-side cost of a step is
  // paid once, and the shards' kernels run CONCURRENTLY -- shard 0 on the caller's stream, every other shard
  // on a stream of its own, forked from and joined back into the caller's stream with events -- so the tail
  // of one kernel overlaps the head of the next instead of leaving SMs idle between two short launches.
  cudaStream_t s0 = static_cast<cudaStream_t>(stream);
  if (n > 1) {
    for (int k = 0; k < n; ++k)
      if (int rc = ensure_streams(handles[k])) return rc;
    NSG_CUDA(cudaEventRecord(handles[0]->host_event, s0));
  }
  for (int k = 0; k < n; ++k) {
    cudaStream_t sk = k == 0 ? s0 : handles[k]->streams[0];
    if (k > 0) NSG_CUDA(cudaStreamWaitEvent(sk, handles[0]->host_event, 0));
    if (int rc = nsgym_step(handles[k], d_actions ? d_actions[k] : nullptr, nullptr, nullptr, skip_updates, sk)) {
      for (int q = 1; q <= k; ++q) cudaStreamSynchronize(handles[q]->streams[0]);
      return rc;
    }
    if (k > 0) {
      NSG_CUDA(cudaEventRecord(handles[k]->host_event, sk));
      NSG_CUDA(cudaStreamWaitEvent(s0, handles[k]->host_event, 0));
    }
  }
  return 0;
}

int nsgym_unpack(NsgymHandle* h, uint8_t* d_terminated, uint8_t* d_truncated, uint8_t* d_was_reset,
                 int32_t* d_relative_time, uint8_t* d_env_change, void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  const uint32_t n = uint32_t(h->spec.n_envs);
  unpack_kernel<<<blocks_for((uint64_t(n) + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      h->buf.d_flags, h->buf.d_change, h->buf.d_t, n, h->spec.n_slots, d_terminated, d_truncated, d_was_reset,
      d_relative_time, d_env_change);
  NSG_CUDA(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int nsgym_step_host(NsgymHandle* h, const void* h_action, const NsgymHostOut* out, int n_chunks, void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  if (!h->initialised) return fail(-4, "step before reset");
  if (!h_action || !out) return fail(-1, "NULL argument");
  if (int rc = ensure_streams(h)) return rc;
  // order the pipeline after whatever the caller's stream still has in flight on this handle's
  // buffers (a reset, a device-side step, a planning copy): the side streams are non-blocking
  // streams and would not wait for it by themselves
  NSG_CUDA(cudaEventRecord(h->host_event, static_cast<cudaStream_t>(stream)));
  for (auto& s : h->streams) NSG_CUDA(cudaStreamWaitEvent(s, h->host_event, 0));
  const int64_t n = h->spec.n_envs;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > n) n_chunks = int(n);
  // chunk boundaries on multiples of 256 envs keep every sub-range 128-byte aligned
  int64_t per = ((n + n_chunks - 1) / n_chunks + 255) / 256 * 256;
  const size_t ab = h->action_bytes(), sb = h->state_bytes_per_env(), w = h->real_bytes();
  const int obs_words = kKinds[h->spec.env_kind].obs_words;
  cudaError_t err = cudaSuccess;
  const char* what = "";
  auto copy = [&](void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s, const char* name) {
    if (err != cudaSuccess) return;
    err = cudaMemcpyAsync(dst, src, bytes, kind, s);
    if (err != cudaSuccess) what = name;
  };
  int c = 0;
  for (int64_t b = 0; b < n && err == cudaSuccess; b += per, ++c) {
    const int64_t cnt = (n - b) < per ? (n - b) : per;
    cudaStream_t s = h->streams[c % kHostStreams];
    copy(static_cast<char*>(h->buf.d_action) + size_t(b) * ab, static_cast<const char*>(h_action) + size_t(b) * ab,
         size_t(cnt) * ab, cudaMemcpyHostToDevice, s, "action H2D");
    if (err != cudaSuccess) break;
    nsg::LaunchIO io = base_io(h);
    io.begin = b;
    io.count = cnt;
    err = dispatch(h, nsg::OP_STEP, io, s);
    if (err != cudaSuccess) { what = "step launch"; break; }
    if (out->h_reward) copy(out->h_reward + b, h->buf.d_reward + b, size_t(cnt) * 4, cudaMemcpyDeviceToHost, s, "reward D2H");
    if (out->h_flags) copy(out->h_flags + b, h->buf.d_flags + b, size_t(cnt), cudaMemcpyDeviceToHost, s, "flags D2H");
    if (out->h_change) copy(out->h_change + b, h->buf.d_change + b, size_t(cnt), cudaMemcpyDeviceToHost, s, "change D2H");
    if (out->h_state)
      copy(static_cast<char*>(out->h_state) + size_t(b) * sb, static_cast<char*>(h->buf.d_state) + size_t(b) * sb,
           size_t(cnt) * sb, cudaMemcpyDeviceToHost, s, "state D2H");
    if (out->h_obs && h->buf.d_obs && obs_words)
      copy(out->h_obs + b * obs_words, h->buf.d_obs + b * obs_words, size_t(cnt) * obs_words * 4,
           cudaMemcpyDeviceToHost, s, "obs D2H");
    if (out->h_delta && h->buf.d_delta)
      for (int j = 0; j < h->spec.n_slots; ++j)
        copy(static_cast<char*>(out->h_delta) + (size_t(j) * n + b) * w,
             static_cast<char*>(h->buf.d_delta) + (size_t(j) * n + b) * w, size_t(cnt) * w, cudaMemcpyDeviceToHost, s,
             "delta D2H");
  }
  // synchronous call: nothing of this step is left in flight on the side streams, error or not
  for (auto& s : h->streams) {
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (err == cudaSuccess && e2 != cudaSuccess) { err = e2; what = "stream synchronize"; }
  }
  if (err != cudaSuccess) return fail(-10, "nsgym_step_host: %s: %s", what, cudaGetErrorString(err));
  h->step_index += 1;
  if (h->plan_elapsed >= 0) h->plan_elapsed += 1;
  return 0;
}

void* nsgym_alloc_host(size_t bytes, int write_combined) {
  void* p = nullptr;
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  if (cudaHostAlloc(&p, bytes ? bytes : 1, flags) != cudaSuccess) {
    fail(-3, "cudaHostAlloc(%zu bytes) failed", bytes);
    return nullptr;
  }
  return p;
}

void nsgym_free_host(void* p) { if (p) cudaFreeHost(p); }

int nsgym_rollout(NsgymHandle* h, int k_steps, int policy, float gamma, float* d_return, int32_t* d_length,
                  int skip_updates, void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  if (!h->initialised) return fail(-4, "rollout before reset");
  if (policy != 0) return fail(-1, "only policy 0 (uniform random) is implemented");
  if (k_steps <= 0) return fail(-1, "k_steps must be positive");
  nsg::LaunchIO io = base_io(h);
  io.k_steps = k_steps;
  io.gamma = gamma;
  io.ret = d_return;
  io.len = d_length;
  io.skip_updates = skip_updates;
  cudaError_t e = dispatch(h, nsg::OP_ROLLOUT, io, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-10, "rollout launch: %s", cudaGetErrorString(e));
  h->step_index += uint64_t(k_steps);
  if (h->plan_elapsed >= 0) h->plan_elapsed += k_steps;
  return 0;
}

int nsgym_rollout_linear(NsgymHandle* h, int k_steps, const void* d_policy, int per_env, float gamma,
                         float* d_return, int32_t* d_length, int skip_updates, void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  if (!h->initialised) return fail(-4, "rollout before reset");
  if (!d_policy) return fail(-1, "NULL policy (nsgym_rollout runs the uniform-random policy)");
  if (k_steps <= 0) return fail(-1, "k_steps must be positive");
  nsg::LaunchIO io = base_io(h);
  io.k_steps = k_steps;
  io.gamma = gamma;
  io.ret = d_return;
  io.len = d_length;
  io.skip_updates = skip_updates;
  io.policy = d_policy;
  io.policy_per_env = per_env ? 1 : 0;
  cudaError_t e = dispatch(h, nsg::OP_ROLLOUT, io, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-10, "rollout launch: %s", cudaGetErrorString(e));
  h->step_index += uint64_t(k_steps);
  if (h->plan_elapsed >= 0) h->plan_elapsed += k_steps;
  return 0;
}

int nsgym_fanout(const NsgymHandle* src, NsgymHandle* dst, int fanout, int theta_from_init, void* stream) {
  if (!src || !dst || !src->bound || !dst->bound) return fail(-1, "both handles must be bound");
  if (!src->initialised) return fail(-4, "the source must be reset before a planning copy is taken");
  if (fanout < 1) return fail(-1, "fanout must be >= 1");
  const NsgymSpec &a = src->spec, &b = dst->spec;
  if (a.env_kind != b.env_kind || a.precision != b.precision || a.n_slots != b.n_slots || a.n_dist != b.n_dist ||
      src->n_istate != dst->n_istate)
    return fail(-1, "planning copy: source and destination differ in env kind / precision / slot list");
  for (int j = 0; j < a.n_slots; ++j)
    if (a.slots[j].theta_index != b.slots[j].theta_index || a.slots[j].istate_plane != b.slots[j].istate_plane)
      return fail(-1, "planning copy: slot %d differs between source and destination", j);
  if (b.n_envs != a.n_envs * int64_t(fanout))
    return fail(-1, "planning copy: destination has %lld envs, expected %lld x %d", (long long)b.n_envs,
                (long long)a.n_envs, fanout);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint32_t n_src = uint32_t(a.n_envs), n_dst = uint32_t(b.n_envs), m = uint32_t(fanout);
  // state: one packed record per env
  {
    const uint32_t words = uint32_t(src->state_bytes_per_env() / 4);
    const uint64_t total = uint64_t(n_dst) * words;
    fanout_records_kernel<<<blocks_for(total), 256, 0, s>>>(static_cast<const uint32_t*>(src->buf.d_state),
                                                           static_cast<uint32_t*>(dst->buf.d_state), total, words, m);
  }
  // episode time word: time, "ended" and "table fresh" bits carry over; CartPole's "terminated
  // once" does not (the copy is a freshly reset env whose state is overwritten)
  fanout_planes_kernel<<<blocks_for(n_dst), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(src->buf.d_t),
                                                        reinterpret_cast<uint32_t*>(dst->buf.d_t), n_dst, n_src, 1, 1, m,
                                                        ~0x40000000u);
  if (src->n_istate > 0)
    fanout_planes_kernel<<<blocks_for(n_dst), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(src->buf.d_istate),
                                                          reinterpret_cast<uint32_t*>(dst->buf.d_istate), n_dst, n_src,
                                                          uint32_t(src->n_istate), 1, m, ~0u);
  const int planes = src->theta_planes();
  if (planes > 0) {
    const uint32_t words = uint32_t(src->real_bytes() / 4);
    if (!theta_from_init) {
      fanout_planes_kernel<<<blocks_for(uint64_t(n_dst) * words), 256, 0, s>>>(
          static_cast<const uint32_t*>(src->buf.d_theta), static_cast<uint32_t*>(dst->buf.d_theta), n_dst, n_src,
          uint32_t(planes), words, m, ~0u);
    } else {
      // classic_control.py:131-135 / toy_text.py:477-481, 675-682: parameters back to their initial values
      PlaneConsts c{};
      const int per = src->grid() ? a.n_dist : 1;
      for (int j = 0; j < a.n_slots; ++j)
        for (int k = 0; k < per; ++k) {
          const double v = a.theta_init[a.slots[j].theta_index][k];
          uint32_t lo, hi = 0;
          if (words == 2) { uint64_t u; std::memcpy(&u, &v, 8); lo = uint32_t(u); hi = uint32_t(u >> 32); }
          else { const float f = float(v); std::memcpy(&lo, &f, 4); }
          c.lo[j * per + k] = lo; c.hi[j * per + k] = hi;
        }
      fill_planes_kernel<<<blocks_for(n_dst), 256, 0, s>>>(static_cast<uint32_t*>(dst->buf.d_theta), n_dst,
                                                          uint32_t(planes), words, c);
    }
  }
  NSG_CUDA(cudaGetLastError());
  dst->initialised = true;
  dst->plan_elapsed = 0;
  dst->launches += 3;
  return 0;
}

int nsgym_transition_table(NsgymHandle* h, int64_t env, int n_times, double* d_prob, int32_t* d_next,
                           float* d_reward, uint8_t* d_done, void* stream) {
  if (!h) return fail(-1, "NULL handle");
  if (!h->grid()) return fail(-1, "transition tables exist for the gridworld kinds only");
  if (env < 0 || env >= h->spec.n_envs) return fail(-1, "env index out of range");
  if (n_times <= 0 || n_times > (1 << 20)) return fail(-1, "n_times out of range");
  if (!d_prob || !d_next || !d_reward || !d_done) return fail(-1, "NULL output");
  nsg::LaunchIO io{};
  io.n = h->spec.n_envs; io.count = h->spec.n_envs;
  io.gid_offset = uint64_t(h->spec.env_id_offset); io.seed = h->spec.seed; io.step_index = h->step_index;
  io.rows = h->rows.active ? &h->rows : nullptr;
  io.sched_replay = h->spec.persistent_params ? 0 : 1;
  io.plan_elapsed = -1;
  cudaError_t e = nsg::launch_table(h->spec, h->pools, io, env, n_times, d_prob, d_next, d_reward, d_done,
                                    static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-10, "transition table: %s", cudaGetErrorString(e));
  h->launches += 2;
  return 0;
}

size_t nsgym_snapshot_bytes(const NsgymHandle* h) {
  if (!h) return 0;
  const size_t n = size_t(h->spec.n_envs);
  return n * h->state_bytes_per_env() + n * size_t(h->theta_planes()) * h->real_bytes() + n * 4 +
         n * size_t(h->n_istate) * 4;
}

namespace {
int snapshot_copy(NsgymHandle* h, char* packed, bool to_packed, cudaStream_t s) {
  const size_t n = size_t(h->spec.n_envs);
  struct Part { void* dev; size_t bytes; } parts[4] = {
      {h->buf.d_state, n * h->state_bytes_per_env()},
      {h->buf.d_theta, n * size_t(h->theta_planes()) * h->real_bytes()},
      {h->buf.d_t, n * 4},
      {h->buf.d_istate, n * size_t(h->n_istate) * 4}};
  size_t off = 0;
  for (const Part& p : parts) {
    if (p.bytes && p.dev) {
      if (to_packed) NSG_CUDA(cudaMemcpyAsync(packed + off, p.dev, p.bytes, cudaMemcpyDeviceToDevice, s));
      else NSG_CUDA(cudaMemcpyAsync(p.dev, packed + off, p.bytes, cudaMemcpyDeviceToDevice, s));
    }
    off += p.bytes;
  }
  return 0;
}
}  // namespace

int nsgym_snapshot(NsgymHandle* h, void* d_dst, NsgymSnapshotInfo* info, void* stream) {
  if (!h || !h->bound || !d_dst || !info) return fail(-1, "NULL argument / handle not bound");
  info->step_index = h->step_index;
  info->plan_elapsed = h->plan_elapsed;
  info->_reserved = 0;
  return snapshot_copy(h, static_cast<char*>(d_dst), true, static_cast<cudaStream_t>(stream));
}

int nsgym_restore(NsgymHandle* h, const void* d_src, const NsgymSnapshotInfo* info, void* stream) {
  if (!h || !h->bound || !d_src || !info) return fail(-1, "NULL argument / handle not bound");
  if (int rc = snapshot_copy(h, const_cast<char*>(static_cast<const char*>(d_src)), false,
                             static_cast<cudaStream_t>(stream)))
    return rc;
  h->step_index = info->step_index;
  h->plan_elapsed = info->plan_elapsed;     // the TimeLimit count of a planning copy rewinds with the state
  h->initialised = true;
  return 0;
}

int nsgym_eval_update(NsgymHandle* h, int slot, void* d_param, const int32_t* d_time, int32_t* d_istate,
                      uint8_t* d_flag, void* d_delta, const double* d_inj_uniform, const double* d_inj_normal,
                      int64_t n, void* stream) {
  if (!h) return fail(-1, "NULL handle");
  if (slot < 0 || slot >= h->spec.n_slots) return fail(-1, "slot out of range");
  if (h->rows.active) return fail(-5, "nsgym_eval_update works on homogeneous handles");
  if (!d_param || !d_time || !d_flag) return fail(-1, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  h->launches += 1;
  if (h->grid())
    e = nsg::launch_eval_dist(h->spec, h->pools, slot, static_cast<double*>(d_param), d_time, d_istate, d_flag,
                              static_cast<double*>(d_delta), d_inj_uniform, n, h->spec.seed, h->step_index, s);
  else if (h->spec.precision == NSGYM_F64)
    e = nsg::launch_eval_scalar_f64(h->spec, h->pools, slot, d_param, d_time, d_istate, d_flag, d_delta,
                                    d_inj_uniform, d_inj_normal, n, h->spec.seed, h->step_index, s);
  else
    e = nsg::launch_eval_scalar_f32(h->spec, h->pools, slot, d_param, d_time, d_istate, d_flag, d_delta,
                                    d_inj_uniform, d_inj_normal, n, h->spec.seed, h->step_index, s);
  if (e != cudaSuccess) return fail(-10, "eval launch: %s", cudaGetErrorString(e));
  h->step_index += 1;
  return 0;
}

int nsgym_eval_w1(int dim, const double* d_u, const double* d_v, double* d_out, double* d_ref, int64_t n,
                  void* stream) {
  if (!d_u || !d_v || !d_out || !d_ref || n < 0) return fail(-1, "bad argument");
  if (dim != 3 && dim != 4) return fail(-1, "dim must be 3 or 4");
  const cudaError_t e = nsg::launch_eval_w1(dim, d_u, d_v, d_out, d_ref, n, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-10, "eval_w1 launch: %s", cudaGetErrorString(e));
  return 0;
}

int nsgym_eval_draws(NsgymHandle* h, int what, int lane, int t, double p, uint64_t step_index, double* d_out,
                     int64_t n, void* stream) {
  if (!h || !d_out) return fail(-1, "NULL argument");
  if (n <= 0 || n > (int64_t(1) << 28)) return fail(-1, "n out of range");
  if (lane < 0 || lane >= NSGYM_MAX_SLOTS) return fail(-1, "lane out of range");
  const bool grid = h->grid();
  const bool ok = grid ? (what == NSGYM_DRAW_SCHED_UNIFORM || what == NSGYM_DRAW_GEOMETRIC ||
                          what == NSGYM_DRAW_DYN_UNIFORM || what == NSGYM_DRAW_DIRICHLET)
                       : ((what >= NSGYM_DRAW_NORMAL && what <= NSGYM_DRAW_GEOMETRIC) ||
                          (what == NSGYM_DRAW_BOX_MULLER_SWEEP && h->spec.precision == NSGYM_F32));
  if (!ok) return fail(-1, "draw kind %d does not exist for this env kind", what);
  nsg::LaunchIO io{};
  io.n = n; io.count = n;
  io.gid_offset = uint64_t(h->spec.env_id_offset); io.seed = h->spec.seed; io.step_index = step_index;
  io.sched_replay = h->spec.persistent_params ? 0 : 1;
  io.plan_elapsed = -1;
  if (what == NSGYM_DRAW_BOX_MULLER_SWEEP) io.begin = int64_t(step_index);   // first radius word of the sweep
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (grid) e = nsg::launch_eval_draws_grid(io, h->spec.n_dist, what, lane, t, p, d_out, s);
  else if (h->spec.precision == NSGYM_F64) e = nsg::launch_eval_draws_f64(io, what, lane, t, p, d_out, s);
  else e = nsg::launch_eval_draws_f32(io, what, lane, t, p, d_out, s);
  if (e != cudaSuccess) return fail(-10, "eval_draws launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return 0;
}

int nsgym_episode_stats(NsgymHandle* h, double* d_running_return, int32_t* d_running_length, double* d_totals,
                        void* stream) {
  if (!h || !h->bound) return fail(-1, "handle not bound");
  if (!d_running_return || !d_running_length || !d_totals) return fail(-1, "NULL argument");
  const uint32_t n = uint32_t(h->spec.n_envs);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned want = blocks_for(n), cap = unsigned(sms) * 8u;
  episode_stats_kernel<<<want < cap ? want : cap, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      h->buf.d_reward, h->buf.d_flags, n, d_running_return, d_running_length, d_totals);
  NSG_CUDA(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int nsgym_set_option(NsgymHandle* h, int option, int64_t value) {
  if (!h) return fail(-1, "NULL handle");
  switch (option) {
    case NSGYM_OPT_GENERAL_KERNELS: h->general_kernels = value != 0; return 0;
    case NSGYM_OPT_SPECIALIZE: h->specialize = value < 0 ? -1 : (value != 0); return 0;
    default: return fail(-1, "unknown option %d", option);
  }
}

void nsgym_set_seed(NsgymHandle* h, uint64_t seed) { if (h) h->spec.seed = seed; }
uint64_t nsgym_step_index(const NsgymHandle* h) { return h ? h->step_index : 0; }
void nsgym_set_step_index(NsgymHandle* h, uint64_t v) { if (h) h->step_index = v; }
int64_t nsgym_launch_count(const NsgymHandle* h) { return h ? h->launches : 0; }
int nsgym_last_kernel_class(const NsgymHandle* h) { return h ? h->kernel_class : -1; }
int nsgym_last_kernel_specialized(const NsgymHandle* h) { return h ? h->specialized : 0; }

int nsgym_jit_check(const NsgymSpec* spec, int rollout, int want_delta, int want_obs, char* source, size_t source_len,
                    char* log, size_t log_len) {
  if (int rc = validate(spec)) return rc;
  NsgymSpec s = *spec;   // as nsgym_create keeps it (no device needed up to here)
  for (int j = 0; j < s.n_slots; ++j) nsg::set_mod_magic(&s.slots[j], *spec);
  nsg::LaunchIO io{};
  io.n = io.count = s.n_envs > 0 ? s.n_envs : 1;
  io.seed = s.seed;
  io.plan_elapsed = -1;
  io.sched_replay = s.persistent_params ? 0 : 1;
  // non-NULL markers only: the generator tests the pointers, nothing dereferences them
  static float marker[2];
  if (want_delta) io.delta = marker;
  if (want_obs) io.obs = marker;
  std::string src;
  io.spec_source = &src;
  nsg::DevicePools pools{nullptr, nullptr, nullptr, nullptr, 0};
  if (nsg::is_grid_kind(s.env_kind)) pools.grid_n_start = nsg::resolve_start_cells(&s);
  const nsg::LaunchOp op = rollout ? nsg::OP_ROLLOUT : nsg::OP_STEP;
  io.k_steps = 1;
  io.gamma = 1.f;
  cudaError_t e;
  if (nsg::is_grid_kind(s.env_kind)) e = nsg::launch_grid(op, s, pools, io, nullptr);
  else if (s.precision == NSGYM_F64) e = nsg::launch_classic_f64(op, s, pools, io, nullptr);
  else e = nsg::launch_classic_f32(op, s, pools, io, nullptr);
  if (e != cudaSuccess || src.empty()) return fail(-2, "this program does not specialise (general kernel class or per-env rows)");
  if (source && source_len) { std::strncpy(source, src.c_str(), source_len - 1); source[source_len - 1] = 0; }
  std::vector<char> cubin;
  std::string text;
  const bool fmad = !nsg::is_grid_kind(s.env_kind) && s.precision != NSGYM_F64;
  const int rc = nsg::jit::compile(src, fmad, &cubin, &text);
  if (log && log_len) { std::strncpy(log, text.c_str(), log_len - 1); log[log_len - 1] = 0; }
  if (rc != 0) return fail(-3, "specialised kernel does not compile: %.400s", text.c_str());
  return int(cubin.size());
}

int nsgym_jit_stats(int64_t* compiled, int64_t* hits, int64_t* failed, char* last_failure, size_t len) {
  const nsg::jit::Stats st = nsg::jit::stats();
  if (compiled) *compiled = st.compiled;
  if (hits) *hits = st.hits;
  if (failed) *failed = st.failed;
  if (last_failure && len) { std::strncpy(last_failure, st.last_failure.c_str(), len - 1); last_failure[len - 1] = 0; }
  return nsg::jit::enabled() ? 1 : 0;
}

}  // extern "C"
