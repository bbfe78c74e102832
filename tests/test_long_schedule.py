"""CPU check of the CTA-wide array kernels' schedule (cudasw4_b200/csrc/kernels_s16_long.cuh, kernels_s32_long.cuh).

The kernels have no spin waits: every cross-warp hand-over relies on the CTA barrier taken every kLongBatch steps. This
test replays the schedule arithmetic (who writes / reads which FIFO slot, border row and ring slot at which step) for
every array width the engine can pick and checks that each read is separated from its write by a barrier, and that no
slot is overwritten before its last reader is past a barrier. Constants are parsed from the kernel header, the shape
rule (W, P, S from the query length) is restated from engine.cu / s16_long_ring_slots."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _constants():
    src = open(os.path.join(ROOT, "cudasw4_b200", "csrc", "kernels_s16_long.cuh")).read()
    get = lambda name: int(re.search(r"constexpr int %s = (\d+);" % name, src).group(1))
    return {k: get(k) for k in ("kLongLag", "kLongFifoRows", "kLongBatch", "kLongMaxWarps")}


def _shape(q, c):
    """engine.cu: P = round_up(q + 32, 16); W = largest power of two <= kLongMaxWarps with kLongLag * W + 64 <= P."""
    P = (q + 32 + 15) // 16 * 16
    W = c["kLongMaxWarps"]
    while W > 1 and c["kLongLag"] * W + 64 > P:
        W //= 2
    S = (c["kLongLag"] * W + 16 + 31) // 32 * 32
    return W, P, S


def _barrier_between(t_write, t_read, batch):
    """A barrier is taken before every step that is a multiple of `batch`: is there one with t_write < b <= t_read?"""
    b = (t_write // batch + 1) * batch
    return b <= t_read


@pytest.mark.parametrize("q", [128, 129, 143, 144, 160, 255, 333, 415, 416, 567, 767, 768, 800, 1000, 5478, 35000])
def test_hand_over_and_ring_are_barrier_separated(q):
    c = _constants()
    lag, fifo, batch = c["kLongLag"], c["kLongFifoRows"], c["kLongBatch"]
    W, P, S = _shape(q, c)
    assert W >= 2 and P % batch == 0 and P >= q + 32 and P >= lag * W + 64 and S >= lag * W + 16 and S % 32 == 0
    assert lag == 32 + batch, "a warp trails its predecessor by its 32 lanes plus one batch"
    rows = sorted(set(list(range(0, min(q, 200))) + list(range(max(0, q - 200), q))))
    for w in range(1, W):  # shared-memory FIFO between warp w-1 (lane 31 writes) and warp w (lane 0 reads)
        for r in rows:
            t_w = r + lag * (w - 1) + 31
            t_r = r + lag * w
            assert _barrier_between(t_w, t_r, batch), (w, r)
            # the slot's next writer: row r + fifo of this period, or the first row of the next period with the same slot
            nxt = [t_w + fifo] if r + fifo < q else []
            r2 = r % fifo
            nxt.append(P + r2 + lag * (w - 1) + 31)
            assert all(_barrier_between(t_r, t, batch) for t in nxt), (w, r)
    for r in rows:  # global border: warp W-1 (period k) -> cp.async prefetch one batch ahead -> warp 0 (period k+1)
        t_w = r + lag * (W - 1) + 31
        t_use = P + r
        t_issue = t_use // batch * batch - batch          # issued at the start of the previous batch ...
        t_done = t_issue + batch                          # ... and waited for at the next barrier
        assert _barrier_between(t_w, t_issue, batch) and t_done <= t_use, r
        # the FIFO slot the prefetch lands in was last read for row r - fifo, or for the previous period's last row with
        # the same slot; the copy is issued right after a barrier, so "read before t_issue" is enough
        prev_reads = [P + r - fifo] if r >= fifo else [rho for rho in range(r % fifo, q, fifo)][-1:]
        assert all(t < t_issue for t in prev_reads), r
    # ring: slot (tau mod S) is filled during the batch before tau's batch (issued right after a barrier, waited for at
    # the next one), read by the lanes until tau + lag (W - 1) + 31, and refilled for tau + S one ring turn later
    last_skew = lag * (W - 1) + 31
    for tau in range(0, 3 * S):
        t_fill_issue = tau // batch * batch - batch
        assert t_fill_issue + batch <= tau
        t_refill_issue = (tau + S) // batch * batch - batch
        assert tau + last_skew < t_refill_issue, (tau, S, W)
    # descriptor table: written by warp 0 at k P, read by warp w at k P + lag w, rewritten at (k + 2) P
    for w in range(1, W):
        assert _barrier_between(0, lag * w, batch) and 2 * P > lag * w + batch


def test_engine_and_kernel_agree_on_the_shape_rule():
    eng = open(os.path.join(ROOT, "cudasw4_b200", "csrc", "engine.cu")).read()
    assert "(qlen + 32 + 15) / 16 * 16" in eng and "kLongLag * longWarps + 64 > p0" in eng
    hdr = open(os.path.join(ROOT, "cudasw4_b200", "csrc", "kernels_s16_long.cuh")).read()
    assert "(kLongLag * warps + 16 + 31) / 32 * 32" in hdr
