"""The host set-up utilities (ordered nearest neighbours, first-fit colouring without the dense scratch, max-min ordering, shard
plan: SURVEY.md 8f ranks 1-2, 8e) against the oracle / the numpy prototype, bit-exact -- the same checks as tests/test_abi_cpu.py
and tests/test_sharding_cpu.py, marked so that they also run on the GPU box, whose host (core count, OpenMP schedule) differs
from the build container's."""
import pytest

import test_abi_cpu as A
import test_sharding_cpu as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,d", [(7, 10, 2), (3000, 10, 2), (2500, 5, 3), (1500, 20, 2), (900, 4, 1)])
def test_find_ordered_nn_bit_exact(n, m, d):
    A.test_find_ordered_nn_bit_exact(n, m, d)


def test_find_ordered_nn_degenerate_geometries():
    A.test_find_ordered_nn_degenerate_geometries()


@pytest.mark.parametrize("n,m", [(50, 3), (2000, 10), (1500, 20)])
def test_greedy_coloring_bit_exact(n, m):
    A.test_greedy_coloring_bit_exact(n, m)


@pytest.mark.parametrize("n,m,d,order", [(120000, 10, 2, "random"), (20000, 10, 2, "maxmin"), (20000, 31, 2, "random")])
def test_parallel_coloring_equals_the_sequential_first_fit(n, m, d, order, monkeypatch):
    A.test_parallel_coloring_equals_the_sequential_first_fit(n, m, d, order, monkeypatch)


def test_order_maxmin_is_exact_farthest_point():
    A.test_order_maxmin_is_exact_farthest_point()


@pytest.mark.parametrize("P", [2, 8])
def test_native_shard_plan_equals_the_numpy_prototype(P):
    S.test_native_plan_equals_the_numpy_prototype(P)


@pytest.mark.parametrize("P", [4])
def test_plans_are_closed_and_mutually_consistent(P):
    S.test_plans_are_closed_and_mutually_consistent(P)
