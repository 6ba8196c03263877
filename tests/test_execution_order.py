"""CPU: the traversal kernel's scheduler hands tickets out in order and an instance waits for its producers, so the
execution order (ticket -> instance slot) must run every producer before its consumers.  rs_plan_check_execution_order
rebuilds the order exactly as rs_create does (plan.cpp: materialize_tasks, build_execution_order) and checks it against a
host mirror of the dispatcher's dependency resolution (kernels.cu: flag_index) -- for the task-major order and for the
parent-board-major walk of a large final round, on one GPU and on every rank of a board-sharded plan."""
import pytest

import rustsolver_b200 as rb
from tests import util


def _plan(case, rank=0, world=1, flags=0):
    if case == "river":
        o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    elif case == "turn":
        o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    else:  # flop-rooted with all-in run-outs: three rounds, 49 turn boards, 2 352 river boards
        o = util.small_options("4d5dAs", ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
    n, tree = rb.build_game_tree(o)
    return rb.Plan(tree, o.ranges(), o.board_mask, [], rank=rank, world_size=world, flags=flags)


@pytest.mark.parametrize("case", ["river", "turn", "flop"])
@pytest.mark.parametrize("flags", [0, rb.RS_FLAG_NO_CHAIN_SPLIT])
def test_task_major_and_board_major_orders_run_producers_first(case, flags):
    plan = _plan(case, flags=flags)
    for trav in range(2):
        n, moved = plan.check_execution_order(trav, force_board_major=False)
        assert n > 0 and moved == 0  # these games are small: slots run in task-major order
        n2, moved2 = plan.check_execution_order(trav, force_board_major=True)
        assert n2 == n
        if case == "flop":
            assert moved2 > n // 2  # 49 parent boards: most river slots move
        elif case == "turn":
            assert moved2 == 0      # one parent board: the parent-board-major walk IS the task-major order


@pytest.mark.parametrize("world", [2, 8])
def test_every_rank_of_a_sharded_plan(world):
    for rank in range(world):
        plan = _plan("flop", rank=rank, world=world)
        for trav in range(2):
            n, moved = plan.check_execution_order(trav, force_board_major=True)
            assert n > 0 and moved > 0


def test_the_checker_rejects_an_order_that_runs_consumers_first():
    plan = _plan("turn")
    with pytest.raises(rb.EngineError, match="runs before its producer"):
        plan.check_execution_order(0, force_board_major=-1)


@pytest.mark.parametrize("world,rank", [(1, 0), (8, 0), (8, 7)])
def test_config4_at_full_size(world, rank):
    """The bench workload itself: 1.4 - 1.6 M instances per traversal, the river walked parent board by parent board
    (the order rs_create uploads for config 4), unsharded and on the first / last rank of an 8-GPU run."""
    from rustsolver_b200 import configs
    w = configs.config4()
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, rank=rank, world_size=world)
    for trav in range(2):
        tickets, moved = plan.check_execution_order(trav)  # not forced: config 4's vectors exceed the L2-sized threshold
        assert tickets > 150_000 and moved > tickets // 2
