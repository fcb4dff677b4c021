"""CPU-only: the planner of the finite-difference share path (dkgv_share_fd_plan is a pure host function
of the C-ABI library, no GPU needed) - invariants of the plan and the decision AUTO takes."""
import pytest

import dvt_circuits_b200 as dk


@pytest.mark.parametrize("t,n", [(683, 1024), (43, 64), (2, 3), (5, 12), (130, 200), (64, 64), (200, 150), (3, 65535)])
def test_plan_invariants(t, n):
    p = dk.share_fd_plan(t, n)
    assert p["exists"]
    m, h = p["parts"], p["h"]
    assert 1 <= m <= 16 and h == -(-t // m) and (m - 1) * h < t  # every part holds at least one coefficient
    assert 2 <= h < n
    assert p["hi"] - p["lo"] + 1 == h and p["lo"] <= 1 <= p["hi"]  # the seed window contains id 1 (and 0 when lo <= 0)
    assert p["steps"] == n - p["hi"]
    assert p["modmul_fd"] > 0 and p["modmul_horner"] > 0
    assert p["use"] == (p["modmul_fd"] * 10 < p["modmul_horner"] * 9)
    # the planner's choice is the cheapest among all forced part counts
    for force in range(1, 17):
        q = dk.share_fd_plan(t, n, force)
        if q["exists"]:
            assert q["parts"] == force and q["modmul_fd"] >= p["modmul_fd"]


def test_plan_headline_shape_and_degenerate_shapes():
    p = dk.share_fd_plan(683, 1024)
    assert p["use"] and p["parts"] > 1
    assert p["modmul_horner"] / p["modmul_fd"] > 4  # BASELINE config B: > 4x fewer field products than Horner per share
    assert p["modmul_horner"] == 79049256  # (t-1) * sum over ids of (signed-digit chain + 12), cf. bench.executed_modmul_per_share
    assert not dk.share_fd_plan(1, 5)["exists"]   # constant polynomial: nothing to difference
    assert not dk.share_fd_plan(4, 2)["exists"]   # two recipients: no window of >= 2 seeds leaves a point to extend
    assert not dk.share_fd_plan(0, 8)["exists"]
    one = dk.share_fd_plan(683, 1024, 1)          # unsplit: t seeds, the window is roughly symmetric around 0
    assert one["parts"] == 1 and one["h"] == 683 and one["lo"] < 0 < one["hi"]
