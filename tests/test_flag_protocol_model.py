"""A model check of the error-flag all-gather the multi-GPU per-step path uses
(wg_xflag_publish / wg_xflag_finish, csrc/wg_kernels.cuh), on CPU: every interleaving of N ranks
that each run, per launch k, `for p: slots[p][set][me] = k` (publish, one store at a time) and then
`for q: wait until slots[me][set][q] == k` (finish, one slot at a time).

The state of the whole system is the tuple of program counters (a rank's stores are determined
by its own counter), so all interleavings are explored exhaustively by a search over those tuples.
With ONE set of slots a deadlock is reachable -- a fast rank's publish of launch k+1 overwrites the
slot a slow rank still waits to see at k; that is the time-out the first version produced at N=8
(profiles/r02_multi_gpu_per_step.txt) -- with TWO alternating sets none is."""
import itertools

import pytest


def explore(n_ranks, launches, sets):
    per_launch = 2 * n_ranks                      # n stores, then n waits
    end = launches * per_launch

    def slot_value(pcs, owner, which_set, writer):
        """what slots[owner][which_set][writer] holds given the writer's program counter: the last
        launch k (of that set) whose store to `owner` the writer has executed"""
        pc = pcs[writer]
        k_done = 0
        full, rest = divmod(pc, per_launch)       # launches fully executed, steps into the next
        for k in range(1, launches + 1):
            if (k % sets) != which_set:
                continue
            stored = (k <= full) or (k == full + 1 and rest > owner)   # stores go to owners 0..n-1 in order
            if stored:
                k_done = k
        return k_done

    def can_step(pcs, r):
        pc = pcs[r]
        if pc >= end:
            return False
        k, step = pc // per_launch + 1, pc % per_launch
        if step < n_ranks:
            return True                            # a store never blocks
        q = step - n_ranks
        return slot_value(pcs, r, k % sets, q) == k

    start = tuple([0] * n_ranks)
    seen, stack, deadlocks = {start}, [start], []
    while stack:
        pcs = stack.pop()
        movers = [r for r in range(n_ranks) if can_step(pcs, r)]
        if not movers and any(pc < end for pc in pcs):
            deadlocks.append(pcs)
            continue
        for r in movers:
            nxt = tuple(pc + (1 if i == r else 0) for i, pc in enumerate(pcs))
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
    return len(seen), deadlocks


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_two_alternating_slot_sets_never_deadlock(n_ranks):
    states, deadlocks = explore(n_ranks, launches=4, sets=2)
    assert states > 100 and deadlocks == []


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_a_single_slot_set_can_deadlock(n_ranks):
    # the bug of the first version, reproduced: some interleaving leaves a rank waiting for a value
    # that has already been overwritten
    _, deadlocks = explore(n_ranks, launches=3, sets=1)
    assert deadlocks, "a single set was expected to admit a deadlock"


# ---- the ghost-plane exchange itself (wg_halo_push / wg_halo_wait) --------------------------------
def explore_halo(n_ranks, steps, wait_for_flags=True):
    """Ranks in a line. Step k of a rank: compute (READS its two ghost planes of array k % 2, which
    must hold the neighbours' step k-1 planes), push its faces into the neighbours' ghost planes of
    the array it just wrote ((k + 1) % 2, version k), publish flag k to both, wait for both
    neighbours' flags >= k. Returns the violations (a ghost plane read with the wrong version)."""
    OPS = 7      # compute, push lo, push hi, flag lo, flag hi, wait lo, wait hi
    end = steps * OPS

    def done_op(pc, k, op):
        """has a rank with program counter pc executed operation `op` of step k?"""
        return pc > (k - 1) * OPS + op

    def ghost_version(pcs, r, side, array):
        """version in rank r's ghost plane (side 0 = from the rank below, 1 = from above) of `array`"""
        w = r - 1 if side == 0 else r + 1
        if w < 0 or w >= n_ranks:
            return None
        push_op = 2 if side == 0 else 1          # the rank below pushes its HI face to me, the one above its LO face
        v = 0 if array == 1 else None            # the initial exchange fills array 1's ghosts (read by step 1) with version 0
        for k in range(1, steps + 1):
            if (k + 1) % 2 == array and done_op(pcs[w], k, push_op):
                v = k
        return v

    def flag(pcs, r, side):
        w = r - 1 if side == 0 else r + 1
        flag_op = 4 if side == 0 else 3          # the rank below sets MY flag with its "flag hi" op
        v = 0
        for k in range(1, steps + 1):
            if done_op(pcs[w], k, flag_op):
                v = k
        return v

    def can_step(pcs, r):
        pc = pcs[r]
        if pc >= end:
            return False
        k, op = pc // OPS + 1, pc % OPS
        if op in (5, 6) and wait_for_flags:
            side = op - 5
            w = r - 1 if side == 0 else r + 1
            if 0 <= w < n_ranks:
                return flag(pcs, r, side) >= k
        return True

    start = tuple([0] * n_ranks)
    seen, stack, violations = {start}, [start], []
    while stack:
        pcs = stack.pop()
        for r in range(n_ranks):
            if not can_step(pcs, r):
                continue
            k, op = pcs[r] // OPS + 1, pcs[r] % OPS
            if op == 0:                           # the read: both ghosts of array k % 2 must be version k - 1
                for side in (0, 1):
                    v = ghost_version(pcs, r, side, k % 2)
                    if v is not None and v != k - 1:
                        violations.append((pcs, r, k, side, v))
            nxt = tuple(pc + (1 if i == r else 0) for i, pc in enumerate(pcs))
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
    return len(seen), violations


def test_ghost_planes_are_always_read_at_the_right_version():
    """the argument at wg_halo_push, checked over every interleaving of three ranks and five steps:
    one flag per direction and two alternating arrays suffice"""
    states, violations = explore_halo(3, steps=5)
    assert states > 1000 and violations == []


def test_without_the_wait_the_exchange_is_wrong():
    # sanity of the model: drop wg_halo_wait and stale / overwritten ghost planes are read
    _, violations = explore_halo(3, steps=3, wait_for_flags=False)
    assert violations
