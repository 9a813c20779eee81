"""Host logic of the persistent dataflow LDL^T (csrc/ldlt_schedule.hpp): the ticket order handed
to the device must be a topological order of the tile DAG that covers every tile exactly once --
that invariant is what makes the device-side spin-waits deadlock-free.  CPU only."""
import numpy as np
import pytest

import ipm_zoo_b200 as z


def expected_counts(n):
    nt = (n + 127) // 128
    rows = [min(128, n - 128 * i) for i in range(nt)]
    halves = [2 if r > 64 else 1 for r in rows]
    trsm = sum(halves[i] * i for i in range(nt))  # tile (i, k) for every k < i
    return nt, trsm


@pytest.mark.parametrize("n", [1, 63, 64, 65, 127, 128, 129, 192, 193, 256, 300, 1000, 2048, 3001, 4097, 8192])
def test_schedule_is_valid_topological_order(n):
    r = z.schedule_check(n)
    assert r["valid"]
    nt, trsm = expected_counts(n)
    assert r["diag"] == nt and r["trsm"] == trsm
    # every off-diagonal tile (i, j) with j >= 1 receives >= 1 update task (the updates of a diagonal tile
    # may all be fused into its DIAGU task)
    assert r["upd"] >= (nt - 1) * (nt - 2) // 2
    assert r["work_us"] >= r["makespan_us"] > 0


@pytest.mark.parametrize("workers", [1, 2, 7, 148, 1000])
def test_schedule_valid_for_any_worker_count(workers):
    for n in (129, 700, 2500):
        r = z.schedule_check(n, workers)
        assert r["valid"]
        # one worker: the list schedule is serial, makespan = total modelled work
        if workers == 1:
            assert abs(r["makespan_us"] - r["work_us"]) < 1e-6 * r["work_us"]


def test_schedule_scaling_model():
    """More workers never make the simulated makespan worse by much, and the n=8192 plan keeps the
    simulated machine >= 80 % busy (the look-ahead chain is hidden behind the bulk updates)."""
    a = z.schedule_check(8192, 148)
    assert a["valid"] and a["work_us"] / (a["makespan_us"] * 148) > 0.8
    b = z.schedule_check(8192, 74)
    assert b["makespan_us"] > a["makespan_us"]


def test_random_sizes():
    rng = np.random.default_rng(0)
    for n in rng.integers(1, 6000, size=40):
        assert z.schedule_check(int(n), int(rng.integers(1, 200)))["valid"]


@pytest.mark.parametrize("n,m,expect_tasks", [(8192, 4096, 3 * 2080), (2048, 1024, 136), (4096, 1, 528), (1024, 512, 36),
                                              (1000, 2000, 0), (512, 256, 0), (8192, 8192, 5 * 2080), (3001, 777, 300)])
def test_assembly_task_list_covers_every_tile_once(n, m, expect_tasks):
    """Condensed assembly on the dataflow kernel (launch_assembly_dataflow): UPD-only task list, every lower tile's K
    range [0, ceil(m/128)) applied exactly once and in order, at most 15 panels per task; not applicable (0 tasks,
    the SYRK kernel runs) below 8 tile rows or when m > n."""
    import ipm_zoo_b200 as z
    ok, ntasks = z.assembly_schedule_check(n, m)
    assert ok
    assert ntasks == expect_tasks
