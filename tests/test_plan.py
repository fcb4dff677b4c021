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


def test_consistency_shortcut_recurrences():
    """The scalar side of the consistency shortcut (share_fd.cu: k_fd_tables, k_fd_polycheck, k_fd_interp), restated with Python
    integers: (2) the t-th forward differences sum_j (-1)^j C(t,j) s(x+j) vanish for every window iff the shares lie on a
    polynomial of degree < t; the Newton interpolant of s(1..t), converted by new_c[i] = c[i-1] / j - c[i] (+ D_{j-1} at i = 0),
    gives back the monomial coefficients."""
    import random
    from math import comb
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    rnd = random.Random(17)
    for t, n in [(2, 5), (3, 9), (7, 12), (20, 33)]:
        a = [rnd.randrange(R) for _ in range(t)]
        s = [sum(c * pow(x, k, R) for k, c in enumerate(a)) % R for x in range(1, n + 1)]
        binom = [(-1) ** j * comb(t, j) % R for j in range(t + 1)]
        windows = lambda seq: [sum(binom[j] * seq[w + j] for j in range(t + 1)) % R for w in range(n - t)]  # noqa: E731
        assert not any(windows(s))
        # a deviation c * prod_{i<=t}(x - i) keeps s(1..t) (and hence the interpolant) intact but is caught by the differences
        dev = list(s)
        for x in range(1, n + 1):
            d = 5
            for i in range(1, t + 1):
                d = d * (x - i) % R
            dev[x - 1] = (dev[x - 1] + d) % R
        assert dev[:t] == s[:t] and any(windows(dev))
        one = list(s)
        one[n - 1] ^= 1
        assert any(windows(one))
        # Newton forward differences of s(1..t), then Horner in the Newton basis on coefficient vectors
        cur = s[:t]
        for r in range(1, t):
            cur = [(cur[i] - cur[i - 1]) % R if i >= r else cur[i] for i in range(t)]
        D = cur
        c = [D[t - 1]] + [0] * (t - 1)
        for j in range(t - 1, 0, -1):
            inv_j = pow(j, -1, R)
            c = [(((c[i - 1] if i else 0) * inv_j - c[i]) + (D[j - 1] if i == 0 else 0)) % R for i in range(t)]
        assert c == a


def test_bench_clock_sampler_window():
    """bench.py's ClockSampler: only samples stamped at or after `since` (the start of the timed region) are summarised;
    throttle reasons are collected; no sample -> None clocks (never a crash)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    def sampler(rows, stamps):
        s = bench.ClockSampler(0)
        s.proc, s.rows, s.stamps = FakeProc(), [list(r) for r in rows], list(stamps)
        return s

    na = "Not Active"
    rows = [["0", "345", "1965", "200.0", "0x0", na, na, na, na],       # idle, before the timed region
            ["0", "1965", "1965", "900.0", "0x0", na, na, na, na],
            ["0", "1950", "1965", "950.0", "0x4", na, na, na, "Active"],
            ["0", "1965", "1965", "940.0", "0x0", na, na, na, na]]
    out = sampler(rows, [1.0, 2.0, 2.1, 2.2]).stop(since=1.5)
    assert {k: out[k] for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")} == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"],
                                                                                    "samples": 3}
    out = sampler(rows, [1.0, 2.0, 2.1, 2.2]).stop()
    assert out["samples"] == 4 and out["sm_mhz"] == 1957.5
    out = sampler(rows, [1.0, 2.0, 2.1, 2.2]).stop(since=5.0)
    assert out["samples"] == 0 and out["sm_mhz"] is None and out["reasons"] == []
    out = sampler(rows, [1.0, 2.0]).stop(since=1.5)  # a row whose stamp has not been appended yet is ignored
    assert out["samples"] == 1


def test_bench_has_no_collective_inside_a_single_rank_leg():
    """bench.py at N > 1: a barrier / max-over-ranks / gather inside an `if rank == 0:` block waits for ranks that never come (the full
    run at N = 2 once hung on the per-step barrier of timed_steps inside the rank-0 config-A leg).  Static check over the source."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)
    collective = {"barrier", "dmax", "all_ranks", "gather_stats"}
    dist_calls = {"barrier", "all_reduce", "all_gather", "all_gather_into_tensor", "broadcast_object_list", "broadcast", "all_gather_object"}

    def is_rank0_test(t):
        return (isinstance(t, ast.Compare) and isinstance(t.left, ast.Name) and t.left.id == "rank" and len(t.ops) == 1
                and isinstance(t.ops[0], ast.Eq) and isinstance(t.comparators[0], ast.Constant) and t.comparators[0].value == 0)

    bad, seen = [], 0
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and is_rank0_test(node.test):
            seen += 1
            for stmt in node.body:
                for c in ast.walk(stmt):
                    if not isinstance(c, ast.Call):
                        continue
                    f = c.func
                    if isinstance(f, ast.Name) and f.id in collective:
                        bad.append((f.id, c.lineno))
                    if isinstance(f, ast.Name) and f.id == "timed_steps":
                        kw = {k.arg: k.value for k in c.keywords}
                        if not (isinstance(kw.get("ranks_together"), ast.Constant) and kw["ranks_together"].value is False):
                            bad.append(("timed_steps without ranks_together=False", c.lineno))
                    if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "dist" and f.attr in dist_calls:
                        bad.append(("dist." + f.attr, c.lineno))
                    if isinstance(f, ast.Attribute) and f.attr in ("comm_init", "all_gather_dev", "share_matrix_verify_sharded_dev",
                                                                   "share_matrix_enqueue_sharded_dev", "share_matrix_enqueue_sharded",
                                                                   "bls_verify_batch_sharded_dev", "agg_final_keys_sharded"):
                        bad.append((f.attr, c.lineno))
    assert seen >= 5
    assert not bad, bad


def test_python_sources_have_no_undefined_names():
    """coarse static check (no linter in this image): every name a function of bench.py / the package loads is bound somewhere in that
    function, at module level or in builtins - a typo in a branch only an N > 1 run takes would otherwise wait for the GPU box"""
    import ast
    import builtins
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = ["bench.py", "__graft_entry__.py", "dvt_circuits_b200/binding.py", "dvt_circuits_b200/synthetic.py", "dvt_circuits_b200/pipeline.py",
             "tools/multi_gpu_check.py", "tools/diag_gather.py", "tools/pairing_lanes.py", "tools/profile_repair.py"]
    for rel in files:
        tree = ast.parse(open(os.path.join(root, rel)).read())
        mod = set(dir(builtins)) | {"__file__", "__name__"}
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                mod.add(n.name)
        for n in tree.body:
            if isinstance(n, (ast.Import, ast.ImportFrom)):
                mod |= {(a.asname or a.name).split(".")[0] for a in n.names}
            elif isinstance(n, (ast.Assign, ast.AugAssign, ast.AnnAssign, ast.For, ast.With, ast.If, ast.Try)):
                mod |= {x.id for x in ast.walk(n) if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Store)}
                for x in ast.walk(n):
                    if isinstance(x, (ast.Import, ast.ImportFrom)):
                        mod |= {(a.asname or a.name).split(".")[0] for a in x.names}
        for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
            bound = set(mod)
            # names of the enclosing functions are visible too: collect over the outermost function that contains fn
            for outer in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and any(x is fn for x in ast.walk(n))]:
                for x in ast.walk(outer):
                    if isinstance(x, ast.Name) and isinstance(x.ctx, (ast.Store, ast.Del)):
                        bound.add(x.id)
                    elif isinstance(x, ast.arg):
                        bound.add(x.arg)
                    elif isinstance(x, (ast.Import, ast.ImportFrom)):
                        bound |= {(a.asname or a.name).split(".")[0] for a in x.names}
                    elif isinstance(x, ast.ExceptHandler) and x.name:
                        bound.add(x.name)
            missing = sorted({(x.id, x.lineno) for x in ast.walk(fn) if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Load) and x.id not in bound})
            assert not missing, (rel, fn.name, missing)
