"""CPU: the proof obligation of the checkpoint schedule (bsi_template.cuh header).

The trajectory kernel compares every state with the checkpoints that are live at that time and
never stores a trajectory.  Model of one lane with transient mu and period lambda: x_s == x_t for
s < t  <=>  s >= mu and lambda | (t - s).  The schedule must guarantee that every lane with
mu + lambda <= H is matched at some time <= end, and then the match time minus the checkpoint time
is exactly lambda; and a match against final slot s >= 1 must imply mu > (time of slot s-1)."""
import pytest

from boolsi_b200 import _lib


def _simulate(mu, lam, times, slots, end):
    """-> (match time, checkpoint time, slot, time held by the previous slot) or None."""
    held = {}                                   # slot -> checkpoint time
    due = dict(zip(times, slots))
    for t in range(end + 1):
        for slot in sorted(held):               # compare first (slot 0 first), then save
            c = held[slot]
            if c < t and c >= mu and (t - c) % lam == 0:
                return t, c, slot, held.get(slot - 1)
        if t in due:
            held[due[t]] = t
    return None


@pytest.mark.parametrize('H', list(range(1, 41)) + [64, 100, 257])
def test_every_closing_lane_is_matched_with_exact_lambda(H):
    times, slots, end, first_final = _lib.attract_schedule(H)
    assert times == sorted(set(times)) and times[-1] == H and end >= H
    assert set(slots) <= {0, 1, 2} and slots[-1] == max(slots)
    finals = [t for t in times if t >= first_final]
    assert [slots[times.index(t)] for t in finals] == list(range(len(finals)))   # compact, in order
    step = 1 if H <= 64 else 7
    for mu in range(0, H, step):
        for lam in range(1, H - mu + 1, step):
            hit = _simulate(mu, lam, times, slots, end)
            assert hit is not None, (H, mu, lam)
            t, c, slot, prev = hit
            assert t - c == lam and c >= mu
            if slot >= 1:
                assert prev is not None and prev < mu        # the lower bound the mu walk starts from
    # ... and nothing that closes later than the horizon is ever reported with a wrong lambda
    for mu in range(0, H + 3, step):
        for lam in (H + 1 - mu if mu <= H else 1, H + 5):
            if lam < 1:
                continue
            hit = _simulate(mu, lam, times, slots, end)
            if hit is not None:
                assert hit[0] - hit[1] == lam


def test_schedule_is_cheap_for_huge_horizons():
    times, slots, end, _ = _lib.attract_schedule(1 << 30)
    assert len(times) < 72 and end < (1 << 31)
