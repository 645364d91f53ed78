"""CPU ORACLE for the two steps either side of the graph path in the servo loop.
TEST INFRASTRUCTURE - NOT PRODUCT CODE (same rules as oracle/cns_oracle.py: tests / smoke / bench only).

  align_cur_to_tar   : cns/frontend/classic.py:142-148 followed by CameraIntrinsic.pixel_to_norm_camera_plane
                       (cns/utils/perception.py:66-68) and the float32 cast of graph_gen.py:257-263
  percentile_f32     : what `np.percentile(err_f32, 90)` computes in the numpy of this image (2.3, method
                       "linear"): q = f32(90)/f32(100); virtual index (n-1)*q in float32; float32 lerp with the
                       two-sided form of numpy's `_lerp`.  Third-party arithmetic (numpy is un-pinned by the
                       reference) restated so the device can follow it operation by operation; pinned bitwise
                       against np.percentile itself in tests/test_stop_policy_cpu.py.
  PixelStopOracle    : PixelStopPolicy.calculate_error (cns/benchmark/stop_policy.py:86-126)
  ErrorHoldingOracle : ErrorHoldingStopPolicy.__call__ (stop_policy.py:6-66) with the float32 arithmetic numpy
                       gives it (cur_err is an np.float32, python floats are weak operands)

Pinning: tests/golden/stop_policy.npz holds the outputs of the UNMODIFIED reference classes over a 150-frame
sequence (tests/golden/make_golden_stop.py); the oracle reproduces them bit for bit.
"""
import numpy as np

F32 = np.float32


def align_cur_to_tar(tar_n, cur_pos, matches, fx, fy, cx, cy):
    """-> (cur_xy_norm float32 (tar_n,2), valid bool (tar_n,)).  Duplicate target indices: the last match wins
    (numpy fancy assignment)."""
    aligned = np.zeros((tar_n, 2), dtype=np.float32)
    valid = np.zeros(tar_n, dtype=bool)
    for t, c in matches:                          # sequential = numpy's assignment order
        aligned[t] = cur_pos[c]
        valid[t] = True
    xy = (aligned.astype(np.float64) - np.array([cx, cy], dtype=np.float64)) / np.array([fx, fy], dtype=np.float64)
    return xy.astype(np.float32), valid


def percentile_f32(err, pct=90):
    s = np.sort(np.asarray(err, dtype=np.float32))
    n = len(s)
    q = F32(pct) / F32(100)
    vidx = F32(n - 1) * q                         # float32 product
    if vidx >= n - 1:                             # only n == 1 for pct < 100
        return s[-1]
    lo = int(np.floor(vidx))
    g = F32(vidx - F32(lo))
    a, b = s[lo], s[lo + 1]
    d = F32(b - a)
    if g >= F32(0.5):
        return F32(b - F32(d * F32(F32(1) - g)))
    return F32(a + F32(d * g))


def pairwise_sum_f32(a):
    """What `np.sum` of a contiguous float32 vector computes (numpy's pairwise_sum: leaves of <= 128 elements
    with 8 running sums, halves split at a multiple of 8).  Restated for the device kernel, which follows it;
    pinned bitwise against np.sum in tests/test_stop_policy_cpu.py.  PixelStopOracle itself just calls np.sum."""
    a = np.asarray(a, dtype=np.float32)
    n = len(a)
    if n < 8:
        r = F32(0)
        for x in a:
            r = F32(r + x)
        return r
    if n <= 128:
        r = a[:8].copy()
        full = n - n % 8
        for i in range(8, full, 8):
            r = r + a[i:i + 8]                    # eight independent float32 running sums
        res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
        for x in a[full:]:
            res = F32(res + x)
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return F32(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


class PixelStopOracle(object):
    def __init__(self, decay=0.95):
        self.decay = decay
        self.maximum = 1.0 / (1 - decay)
        self.weight = None
        self.witness = None

    def reset(self):
        if self.weight is not None:
            self.weight[:] = 0
            self.witness[:] = 0

    def calculate_error(self, pos_cur, pos_tar, node_mask):
        pos_cur, pos_tar = np.asarray(pos_cur, np.float32), np.asarray(pos_tar, np.float32)
        m = np.asarray(node_mask, bool)
        if self.weight is None or len(self.weight) != len(m):
            self.weight = np.zeros(len(m), np.float32)
            self.witness = np.zeros(len(m), np.int64)
        first = (self.witness == 0) & m           # pre_is_0 & now_is_1
        self.witness[m] += 1
        w = self.weight
        upd = F32(self.decay) * w                 # float32 product, then + {0,1}
        upd = upd + m.astype(np.float32)
        w[:] = np.where(first, F32(self.maximum), upd)
        d = pos_cur - pos_tar
        delta = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])[m]
        ws = w[m]
        err = ws * delta
        keep = err < percentile_f32(err)
        ws = ws[keep]
        ws = ws / (ws.sum() + F32(1e-7))
        return np.sum(ws * delta[keep])


class ErrorHoldingOracle(object):
    """State machine of stop_policy.py:6-66, fed with the error instead of the data object.  Written with plain
    python operators on the np.float32 error so numpy's scalar promotion (python floats are weak operands)
    produces the reference's arithmetic by construction."""

    def __init__(self, waiting_time, conduct_thresh, on_reset=None):
        self.waiting_time, self.conduct_thresh = waiting_time, conduct_thresh
        self.cur_err = float("inf")
        self.on_reset = on_reset                  # PixelStopPolicy.reset also clears weights / witness counts (:96-100)
        self.reset()

    def reset(self):
        self.min_err, self.prev_t, self.prev_err, self.accum_time = None, None, None, 0
        if self.on_reset is not None:
            self.on_reset()

    def __call__(self, cur_err, timestamp):
        if cur_err is None:                       # no graph this frame (stop_policy.py:26-29)
            self.reset()
            self.cur_err = float("inf")
            return False
        self.cur_err = cur_err
        if cur_err > self.conduct_thresh:         # too far: forget the history (:34-36)
            self.reset()
            return False
        if self.min_err is None or cur_err < 0.9 * self.min_err:      # new best: restart the clock (:38-43)
            self.min_err, self.prev_t, self.prev_err, self.accum_time = cur_err, timestamp, cur_err, 0
            return False
        level = 1.1 * self.min_err + 5e-4
        e0, e1, t0 = self.prev_err, cur_err, self.prev_t
        self.prev_t, self.prev_err = timestamp, cur_err
        # time within [t0, t1] that the linearly interpolated error spent at or below `level` (:52-60)
        if e0 < level <= e1:
            held = (level - e0) / (e1 - e0) * (timestamp - t0)
        elif e1 < level <= e0:
            held = (level - e1) / (e0 - e1) * (timestamp - t0)
        elif e0 <= level and e1 <= level:
            held = timestamp - t0
        else:
            held = 0
        self.accum_time += held
        return self.accum_time > self.waiting_time
