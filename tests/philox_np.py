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
