// nsgym_rows.cu -- lowering of per-env rows (nsgym_create_rows) into SoA planes.
//
// Every row is lowered exactly like a homogeneous slot (lower_slot), turned into canonical row
// words, and compared with env 0: a word that never differs stays a launch constant, a word
// that differs somewhere becomes one coalesced plane [n_envs] in device memory.  The traffic a
// heterogeneous step adds is therefore exactly the per-env information the batch carries.
#include <climits>
#include <cstdio>
#include <cstring>
#include <vector>

#include "nsgym_classic_launch.cuh"

namespace nsg {

namespace {

template <typename R>
void lower_row(const NsgymSpec& spec, const NsgymSlot& src, int j, int32_t (&iw)[kRowInt], double (&rw)[kRowReal],
               double (&dw)[kRowDbl]) {
  NsgymSlot a = src;
  set_mod_magic(&a, spec);
  const SlotT<R> b = lower_slot<R>(a, j, reachable_t_max(spec));
  row_words<R>(b, a, iw, rw, dw);
}

template <typename R>
int build_rows_t(const NsgymSpec& spec, const NsgymSlot* rows, RowTable* out, char* err, size_t err_len) {
  const int64_t n = spec.n_envs;
  const int np = spec.n_slots;
  RowTable t;
  t.active = true;
  t.lean = true;
  t.precision = std::is_same<R, double>::value ? NSGYM_F64 : NSGYM_F32;
  int32_t iw[kRowInt];
  double rw[kRowReal], dw[kRowDbl];
  // ---- pass 1: validate, defaults from env 0, which words vary, value range of the int words ----
  int32_t lo[NSGYM_MAX_SLOTS][kRowInt], hi_v[NSGYM_MAX_SLOTS][kRowInt];
  for (int j = 0; j < np; ++j) {
    lower_row<R>(spec, rows[j], j, t.def_int[j], t.def_real[j], t.def_dbl[j]);
    for (int w = 0; w < kRowInt; ++w) { lo[j][w] = INT32_MAX; hi_v[j][w] = INT32_MIN; }
  }
  for (int64_t e = 0; e < n; ++e) {
    for (int j = 0; j < np; ++j) {
      const NsgymSlot& a = rows[e * np + j];
      const NsgymSlot& key = spec.slots[j];
      if (a.theta_index != key.theta_index || a.constraint != key.constraint ||
          a.partner_slot != key.partner_slot || a.partner_index != key.partner_index ||
          (a.istate_plane >= 0 && a.istate_plane != key.istate_plane)) {
        snprintf(err, err_len, "row (env %lld, slot %d): theta_index / constraint / partner / istate_plane differ "
                 "from spec->slots[%d] (the key set is shared by the batch)", (long long)e, j, j);
        return -1;
      }
      if (a.ui[2] != key.ui[2] || a.uf[5] != key.uf[5]) {
        snprintf(err, err_len, "row (env %lld, slot %d): the Lipschitz bound (ui[2], uf[5]) is shared by the batch",
                 (long long)e, j);
        return -1;
      }
      if (a.ui[3] != key.ui[3]) {
        snprintf(err, err_len, "row (env %lld, slot %d): a Memoryless scheduler driving a list update (packed cursor, "
                 "ui[3]) must be declared by spec->slots[%d] for the whole batch", (long long)e, j, j);
        return -1;
      }
      if (int rc = validate_row_slot(&spec, &a, j, err, err_len)) return rc;
      lower_row<R>(spec, a, j, iw, rw, dw);
      {   // lean kernels: deterministic schedulers; fast / medium scalar rules or deterministic distribution rules
        const bool det_sched = a.sched_op != NSGYM_SCHED_RANDOM && a.sched_op != NSGYM_SCHED_DECAY &&
                               a.sched_op != NSGYM_SCHED_MEMORYLESS;
        const bool dist = a.upd_op >= NSGYM_UPD_D_NOP;
        const bool lean_upd = dist ? (a.upd_op != NSGYM_UPD_D_RANDOM && a.ui[2] == 0)
                                   : !((iw[RI_OPS] & 0x1F) & SF_SLOW_UPD);
        t.lean = t.lean && det_sched && lean_upd;
      }
      uint32_t m = 0;
      for (int w = 0; w < kRowInt; ++w) {
        m |= (iw[w] != t.def_int[j][w]) ? (1u << w) : 0u;
        if (iw[w] < lo[j][w]) lo[j][w] = iw[w];
        if (iw[w] > hi_v[j][w]) hi_v[j][w] = iw[w];
      }
      for (int w = 0; w < kRowReal; ++w)
        m |= (std::memcmp(&rw[w], &t.def_real[j][w], sizeof(double)) != 0) ? (1u << (kRowInt + w)) : 0u;
      for (int w = 0; w < kRowDbl; ++w)
        m |= (std::memcmp(&dw[w], &t.def_dbl[j][w], sizeof(double)) != 0) ? (1u << (kRowInt + kRowReal + w)) : 0u;
      t.mask[j] |= m;
    }
  }
  // ---- plane assignment ----
  for (int j = 0; j < np; ++j) {
    // int words: bit-packed, in word order, into the planes of this slot -- a word takes the bits its
    // largest value over the batch needs (all 32 when it is negative somewhere)
    int used = 32;                                      // bits taken in the current plane (32: open a new one)
    for (int w = 0; w < kRowInt; ++w) {
      if (!((t.mask[j] >> w) & 1u)) continue;
      int bits = 32;
      if (lo[j][w] >= 0) {
        bits = 1;
        while (bits < 32 && (uint32_t(hi_v[j][w]) >> bits) != 0u) ++bits;
      }
      if (used + bits > 32) { t.n_int++; used = 0; }
      t.plane[j][w] = uint8_t(t.n_int - 1);
      t.shift[j][w] = uint8_t(used);
      t.bits[j][w] = uint8_t(bits);
      used += bits;
    }
    for (int w = 0; w < kRowReal; ++w)
      if ((t.mask[j] >> (kRowInt + w)) & 1u) t.plane[j][kRowInt + w] = uint8_t(t.n_real++);
    for (int w = 0; w < kRowDbl; ++w)
      if ((t.mask[j] >> (kRowInt + kRowReal + w)) & 1u) t.plane[j][kRowInt + kRowReal + w] = uint8_t(t.n_dbl++);
  }
  t.bytes_per_env = 4.0 * t.n_int + double(sizeof(R)) * t.n_real + 8.0 * t.n_dbl;
  {   // the kernels index planes with 32-bit arithmetic
    const int widest = t.n_int > t.n_real ? (t.n_int > t.n_dbl ? t.n_int : t.n_dbl) : (t.n_real > t.n_dbl ? t.n_real : t.n_dbl);
    if (uint64_t(widest) * uint64_t(n) >= (1ull << 32)) {
      snprintf(err, err_len, "%d row planes x %lld envs exceed 32-bit plane indexing: split the batch into shards",
               widest, (long long)n);
      return -1;
    }
  }
  // ---- pass 2: fill the planes ----
  std::vector<int32_t> hi(size_t(t.n_int) * n);
  std::vector<R> hr(size_t(t.n_real) * n);
  std::vector<double> hd(size_t(t.n_dbl) * n);
  for (int64_t e = 0; e < n; ++e) {
    for (int j = 0; j < np; ++j) {
      lower_row<R>(spec, rows[e * np + j], j, iw, rw, dw);
      const uint32_t m = t.mask[j];
      for (int w = 0; w < kRowInt; ++w)
        if ((m >> w) & 1u)
          hi[size_t(t.plane[j][w]) * n + e] |= int32_t(uint32_t(iw[w]) << t.shift[j][w]);
      for (int w = 0; w < kRowReal; ++w)
        if ((m >> (kRowInt + w)) & 1u) hr[size_t(t.plane[j][kRowInt + w]) * n + e] = R(rw[w]);
      for (int w = 0; w < kRowDbl; ++w)
        if ((m >> (kRowInt + kRowReal + w)) & 1u) hd[size_t(t.plane[j][kRowInt + kRowReal + w]) * n + e] = dw[w];
    }
  }
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    *dst = nullptr;
    if (!bytes) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t e = up(hi.data(), hi.size() * 4, reinterpret_cast<void**>(&t.d_int));
  if (e == cudaSuccess) e = up(hr.data(), hr.size() * sizeof(R), &t.d_real);
  if (e == cudaSuccess) e = up(hd.data(), hd.size() * 8, reinterpret_cast<void**>(&t.d_dbl));
  if (e != cudaSuccess) {
    free_rows(&t);
    snprintf(err, err_len, "row planes: %s", cudaGetErrorString(e));
    return -10;
  }
  *out = t;
  return 0;
}

}  // namespace

int build_rows(const NsgymSpec& spec, const NsgymSlot* rows, bool real_is_double, RowTable* out, char* err,
               size_t err_len) {
  return real_is_double ? build_rows_t<double>(spec, rows, out, err, err_len)
                        : build_rows_t<float>(spec, rows, out, err, err_len);
}

void free_rows(RowTable* t) {
  if (!t) return;
  cudaFree(t->d_int);
  cudaFree(t->d_real);
  cudaFree(t->d_dbl);
  t->d_int = nullptr; t->d_real = nullptr; t->d_dbl = nullptr;
  t->active = false;
}

}  // namespace nsg
