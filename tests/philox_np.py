"""NumPy restatement of the device Philox4x32-10 stream layout (tests only): an independent
check of the counter-based RNG and of the rollout policy's action draws."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, seed):
    """Counters: uint32 arrays (broadcastable); seed: python int (64 bit).  Returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & MASK,
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & MASK]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def block(gids, step_index, blk, seed):
    gids = np.asarray(gids, dtype=np.uint64)
    return philox4x32_10(gids & MASK, gids >> np.uint64(32), np.uint64(step_index & 0xFFFFFFFF),
                         np.uint64((((step_index >> 32) << 8) | blk) & 0xFFFFFFFF), seed)


BLK_POLICY = 13
BLK_PAIR = 14


def policy_actions(kind, gids, step_index, seed, precision="fp64"):
    """The rollout kernels' uniform-random policy (nsgym_device.cuh / nsgym_grid.cuh)."""
    if kind == "grid":      # block 14 of the step PAIR, half = step & 1; the two lowest bits of the half's low word
        w = block(gids, step_index >> 1, BLK_PAIR, seed)
        return (w[3 if step_index & 1 else 1] & np.uint32(3)).astype(np.int32)
    if precision == "fp32":  # the low bytes of block 0's words (the fp32 draws use the top 24 bits)
        b = block(gids, step_index, 0, seed)
        x = ((b[0] & np.uint32(0xFF)) | ((b[1] & np.uint32(0xFF)) << np.uint32(8)) |
             ((b[2] & np.uint32(0xFF)) << np.uint32(16)) | (b[3] << np.uint32(24))).astype(np.uint32)
    else:
        x = block(gids, step_index, BLK_POLICY, seed)[0]
    if kind == "cartpole":
        return (x >> np.uint32(31)).astype(np.int32)
    if kind in ("acrobot", "mountaincar"):
        return ((x.astype(np.uint64) * np.uint64(3)) >> np.uint64(32)).astype(np.int32)
    u = (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    if kind == "pendulum":
        return np.float32(-2) + np.float32(4) * u
    if kind == "mountaincar_cont":
        return np.float32(-1) + np.float32(2) * u
    raise KeyError(kind)


# ---------------------------------------------------------------------------------------------
# Native draw transforms of the step kernels (nsgym_device.cuh: Rng<R>; nsgym_grid.cuh:
# dirichlet_ones), restated so that the oracle port can be fed the very numbers the throughput
# kernels draw: the benched lean kernels are then compared with the oracle directly
# (tests/test_gpu_native_parity.py).  nsgym_eval_draws returns the device's own values; the test
# test_gpu_native_draws.py holds these restatements to them value by value.
# ---------------------------------------------------------------------------------------------
BLK_MAIN, BLK_SCHED0, BLK_DIRICHLET0, BLK_RESET2 = 0, 4, 8, 12
SCHED_REPLAY_TAG = 0x80


def _half(b, half):
    return (b[2], b[3]) if half else (b[0], b[1])


def unit53(hi, lo):
    v = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (v >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def unit24(x):
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def std_normal(gids, step_index, lane, seed, precision):
    """Rng<R>::std_normal: Box-Muller from one 64-bit half block.  fp32: 24 + 24 bits; the device
    evaluates lg2 / sqrt / cos with MUFU approximations (abs error ~1e-6), here in float64 from the
    same float32 inputs.  fp64: 32 + 32 bits, accurate log / sqrt / sincospi."""
    wx, wy = _half(block(gids, step_index, lane >> 1, seed), lane & 1)
    if precision == "fp32":
        u1 = ((wx >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
        ang = (wy >> np.uint32(8)).astype(np.float32) * np.float32(6.283185307179586 / 16777216.0)
        return np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos(ang.astype(np.float64))
    u1 = (wx.astype(np.float64) + 1.0) * (1.0 / 4294967296.0)
    u2 = wy.astype(np.float64) * (1.0 / 4294967296.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def reset_uniforms(gids, step_index, seed, precision):
    """Rng<R>::reset_uniforms: [4, n] initial-state uniforms (fp32: the top 24 bits of block 0's
    words; fp64: 53 bits from each half of blocks 0 and 12)."""
    a = block(gids, step_index, BLK_MAIN, seed)
    if precision == "fp32":
        return np.stack([unit24(w).astype(np.float64) for w in a])
    b = block(gids, step_index, BLK_RESET2, seed)
    return np.stack([unit53(a[0], a[1]), unit53(a[2], a[3]), unit53(b[0], b[1]), unit53(b[2], b[3])])


def sched_uniform(gids, step_index, lane, seed, t=None):
    """Rng::sched_uniform: keyed by the episode time t (counter (env, t, 0x80 | block)) when the
    scheduler replays per episode (persistent_params False), by the step index otherwise."""
    blk = BLK_SCHED0 + (lane >> 1)
    if t is None:
        b = block(gids, step_index, blk, seed)
    else:
        g = np.asarray(gids, dtype=np.uint64)
        b = philox4x32_10(g & MASK, g >> np.uint64(32), np.asarray(t, dtype=np.uint64), np.uint64(SCHED_REPLAY_TAG | blk), seed)
    hi, lo = _half(b, lane & 1)
    return unit53(hi, lo)


def dyn_uniform(gids, step_index, seed):
    """Rng::dyn_uniform (gridworlds): block 14 of the step PAIR, half = step & 1, top 53 bits."""
    hi, lo = _half(block(gids, step_index >> 1, BLK_PAIR, seed), step_index & 1)
    return unit53(hi, lo)


def dirichlet_ones(gids, step_index, lane, seed, attempt, dim):
    """dirichlet_ones (native branch): standard exponentials e_k = -log((w_k + 0.5) / 2^32) from block
    8 + lane (the attempt index enters counter word 1), scaled by the reciprocal of their sum."""
    g = np.asarray(gids, dtype=np.uint64)
    c1 = (g >> np.uint64(32)) ^ np.uint64((attempt << 12) & 0xFFFFFFFF)
    w = philox4x32_10(g & MASK, c1, np.uint64(step_index & 0xFFFFFFFF),
                      np.uint64((((step_index >> 32) << 8) | (BLK_DIRICHLET0 + lane)) & 0xFFFFFFFF), seed)
    e = [-np.log((w[k].astype(np.float64) + 0.5) * (1.0 / 4294967296.0)) for k in range(dim)]
    acc = e[0]
    for k in range(1, dim):
        acc = acc + e[k]
    inv = 1.0 / acc
    return np.stack([v * inv for v in e])


def native_tables(n_envs, n_rows, n_slots, seed, precision, grid, gid_offset=0):
    """The injected-table layout of oracle/streams.py filled with the kernels' NATIVE draws:
    uniforms[K, L, N] (lane 0 slip draw, lanes 1..4 reset draws, Dirichlet lanes as the
    exponentials' uniforms are not representable -> see NativeEnvStreams.dirichlet) and
    normals[K, P, N]; row r is Philox step index r (reset = 0, step k = k + 1)."""
    from oracle import streams as S_

    gids = np.arange(n_envs, dtype=np.uint64) + np.uint64(gid_offset)
    L = S_.n_uniform_lanes(n_slots)
    u = np.zeros((n_rows, L, n_envs))
    z = np.zeros((n_rows, max(n_slots, 1), n_envs))
    for r in range(n_rows):
        if grid:
            u[r, S_.LANE_DYN] = dyn_uniform(gids, r, seed)
        else:
            u[r, S_.LANE_RESET0:S_.LANE_RESET0 + 4] = reset_uniforms(gids, r, seed, precision)
            for j in range(min(n_slots, 8)):
                z[r, j] = std_normal(gids, r, j, seed, precision)
    return u, z


def native_streams(n_envs, n_rows, n_slots, seed, precision, grid, replay=True, gid_offset=0, t_max=1024):
    """(clock, [per-env streams], u, z): per-env stream objects for the oracle port whose draws are
    the kernels' native ones.  Scheduler uniforms are keyed by the episode time when `replay`."""
    from oracle import streams as S_

    u, z = native_tables(n_envs, n_rows, n_slots, seed, precision, grid, gid_offset)
    gids = np.arange(n_envs, dtype=np.uint64) + np.uint64(gid_offset)
    clock = S_.Clock()
    if replay:      # su[t, slot, env]
        su = np.stack([np.stack([sched_uniform(gids, 0, j, seed, t=t) for j in range(max(n_slots, 1))])
                       for t in range(t_max)])
    else:           # su[row, slot, env]
        su = np.stack([np.stack([sched_uniform(gids, r, j, seed) for j in range(max(n_slots, 1))])
                       for r in range(n_rows)])

    class NativeEnvStreams(S_.EnvStreams):
        def __init__(self, i):
            super().__init__(u[:, :, i], z[:, :, i], clock)
            self.i = i

        def sched_uniform(self, slot, t):
            return float(su[t if replay else self.clock.k, slot, self.i])

        def dirichlet(self, slot, n):
            key = (self.clock.k, slot)
            attempt = self._dir_seen.get(key, 0)
            self._dir_seen = {key: attempt + 1}
            p = dirichlet_ones(gids[self.i:self.i + 1], self.clock.k, slot, seed, attempt, n)
            return [float(v) for v in p[:, 0]]

    return clock, [NativeEnvStreams(i) for i in range(n_envs)], u, z
